/* kazen -- command line front end (main.cpp:20-83 of the reference): kazen <scene.xml> [options].
 * The XML is parsed through the plugin system; the "path_mis" integrator and the "gpu_bvh"
 * accelerator resolve to the B200 implementations behind include/kzgpu.h. */
#include <kazen/scene.h>
#include <cstring>
#include <iostream>

using namespace kazen;

static void usage() {
    std::cerr << "Syntax: kazen <scene.xml> [-o <output stem>] [--gpus N] [--spp N] [--size WxH] [--sampler <name>]\n"
                 "              [--accel sah|lbvh] [--max-depth N] [--raw] [--print] [--spp-range A:B] [--resume <stem>.rgbw]\n";
}

int main(int argc, char **argv) {
    if (argc < 2) { usage(); return -1; }
    std::string sceneName, out, resume; int gpus = 1, sppBegin = 0, sppEnd = -1; bool raw = false, print = false;
    ParseOverrides ov;
    try {
        for (int i = 1; i < argc; ++i) {
            const std::string a = argv[i];
            auto next = [&]() -> std::string { if (i + 1 >= argc) throw Exception("missing value after " + a); return argv[++i]; };
            if (a == "-o") out = next();
            else if (a == "--gpus") gpus = std::stoi(next());
            else if (a == "--spp") ov["sampler"]["sampleCount"] = "i:" + next();
            else if (a == "--sampler") ov["sampler"]["type"] = "s:" + next();
            else if (a == "--max-depth") ov["integrator"]["maxDepth"] = "i:" + next();
            else if (a == "--accel") ov["scene"]["accelBuilder"] = "s:" + next();
            else if (a == "--size") { const std::string v = next(); const size_t x = v.find('x'); if (x == std::string::npos) throw Exception("--size expects WxH");
                                      ov["camera"]["width"] = "i:" + v.substr(0, x); ov["camera"]["height"] = "i:" + v.substr(x + 1); }
            else if (a == "--raw") raw = true;
            else if (a == "--resume") resume = next();
            else if (a == "--spp-range") { const std::string v = next(); const size_t c = v.find(':'); if (c == std::string::npos) throw Exception("--spp-range expects A:B");
                                           sppBegin = std::stoi(v.substr(0, c)); sppEnd = std::stoi(v.substr(c + 1)); }
            else if (a == "--print") print = true;
            else if (!a.empty() && a[0] == '-') throw Exception("unknown option " + a);
            else sceneName = a;
        }
        if (sceneName.size() < 4 || sceneName.substr(sceneName.size() - 4) != ".xml") throw Exception("Fatal error: unknown file \"" + sceneName + "\", expected an extension of type .xml");
        const size_t slash = sceneName.find_last_of('/');
        resolverPrepend(slash == std::string::npos ? "." : sceneName.substr(0, slash));      /* main.cpp:52 */
        std::unique_ptr<Object> root(loadFromXML(sceneName, &ov));
        if (root->getClassType() != Object::EScene) throw Exception("the root element must be a <scene>");
        Scene *scene = static_cast<Scene *>(root.get());
        scene->gpus = gpus; scene->sppBegin = sppBegin; scene->sppEnd = sppEnd; scene->resumeFrame = resume;
        if (print) std::cout << scene->toString() << std::endl;
        if (out.empty()) out = sceneName.substr(0, sceneName.size() - 4);
        renderer::render(scene, out, raw);
    } catch (const std::exception &e) {
        std::cerr << "Fatal error: " << e.what() << std::endl;
        return -1;
    }
    return 0;
}
