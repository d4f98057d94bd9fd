/* capi.cpp -- a small C view of the C++ host for tests and tools: load a kazen XML scene through
 * the plugin system and hand out the flattened POD tables (so the same tables can be fed to the
 * GPU path and, in tests, to the oracle), or render it. */
#include <kazen/scene.h>
#include <cstring>
#include <iostream>

using namespace kazen;

static thread_local std::string g_err;

extern "C" {

const char *kazen_host_last_error() { return g_err.c_str(); }

/* overrides: NULL or "tag.prop=t:value;tag.prop=t:value" (t in i,f,b,s; "tag.type=s:name" swaps the plugin) */
void *kazen_host_load(const char *xml_path, const char *overrides) {
    try {
        ParseOverrides ov;
        if (overrides) {
            std::string s(overrides);
            size_t b = 0;
            while (b < s.size()) {
                size_t e = s.find(';', b); if (e == std::string::npos) e = s.size();
                const std::string item = s.substr(b, e - b);
                const size_t dot = item.find('.'), eq = item.find('=');
                if (dot != std::string::npos && eq != std::string::npos && dot < eq) ov[item.substr(0, dot)][item.substr(dot + 1, eq - dot - 1)] = item.substr(eq + 1);
                b = e + 1;
            }
        }
        std::string path(xml_path);
        const size_t slash = path.find_last_of('/');
        resolverPrepend(slash == std::string::npos ? "." : path.substr(0, slash));
        Object *o = loadFromXML(path, &ov);
        if (o->getClassType() != Object::EScene) { delete o; g_err = "root object is not a scene"; return nullptr; }
        return static_cast<Scene *>(o);
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
const kz_scene_desc *kazen_host_scene_desc(void *scene) {
    try { return &static_cast<Scene *>(scene)->flatten(); } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}
int kazen_host_accel_builder(void *scene) { return static_cast<Scene *>(scene)->getAccel()->builder(); }
int kazen_host_describe(void *scene, char *buf, size_t n) {
    const std::string s = static_cast<Scene *>(scene)->toString();
    snprintf(buf, n, "%s", s.c_str());
    return (int)s.size();
}
int kazen_host_render(void *scene, const char *output_stem, int gpus, int write_raw) {
    try { Scene *s = static_cast<Scene *>(scene); s->gpus = gpus; renderer::render(s, output_stem, write_raw != 0); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}
void kazen_host_free(void *scene) { delete static_cast<Scene *>(scene); }
int kazen_host_registered(char *buf, size_t n) {
    std::string s;
    for (const std::string &k : ObjectFactory::registeredNames()) s += k + " ";
    snprintf(buf, n, "%s", s.c_str());
    return (int)s.size();
}
/* decode an image file the way imagetexture does; rgb must hold 3*w*h floats when non-NULL (call twice) */
int kazen_host_read_image(const char *path, int *w, int *h, float *rgb) {
    std::vector<float> px; std::string err;
    if (!readImage(path, *w, *h, px, err)) { g_err = err; return -1; }
    if (rgb) memcpy(rgb, px.data(), px.size() * sizeof(float));
    return 0;
}
int kazen_host_write_exr(const char *path, int w, int h, const float *rgb) {
    try { writeEXR(path, w, h, rgb); return 0; } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
void kazen_host_fallback_tables(uint16_t *blue_noise, uint32_t *pmj) {
    std::vector<uint16_t> bn; std::vector<uint32_t> pm;
    fallbackPmjTables(bn, pm);
    memcpy(blue_noise, bn.data(), bn.size() * 2); memcpy(pmj, pm.data(), pm.size() * 4);
}

}  // extern "C"
