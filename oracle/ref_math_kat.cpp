/* oracle/ref_math_kat.cpp -- TEST INFRASTRUCTURE, built only where /root/reference is mounted (oracle/Makefile, target _ref/ref_math_kat).
 *
 * Runs the reference's OWN shading-math function bodies -- include/kazen/ggx_brdf.h, frame.h, dpdf.h compiled in place, and the
 * bodies of coordinateSystem / fresnel / fresnelDielectric / refract / reflect (src/kazen/common.cpp:436-540) and of the Warp
 * functions (src/kazen/warp.cpp:41-130) extracted at build time by line range into oracle/_ref/ -- next to the oracle's
 * restatements (kzo_math.h, kzo_shading.h) on seeded random inputs, compares every output BIT FOR BIT, and prints the reference
 * outputs as JSON (tests/golden/math_kat.json) so the same check runs wherever the reference is absent.
 *
 * The reference's third-party headers are replaced by oracle/ref_shim (a 100-line eager stand-in for the Eigen vector types, empty
 * OpenImageIO / fmt stubs): Eigen's coefficient-wise semantics are simple enough to restate, the function bodies under test are the
 * reference's own.  <math.h> is included first so that the unqualified abs/pow/cos/sin of ggx_brdf.h bind to the float overloads,
 * as they do in the reference build through OpenImageIO's headers. */
#include <math.h>
#include <stdlib.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <kazen/common.h>
#include <kazen/vector.h>
#include <kazen/color.h>
#include <kazen/frame.h>
namespace kazen {
#include "_ref/common_extract.inc"
struct Warp {            /* declarations as in include/kazen/warp.h:24-51 (that header drags in the plugin system) */
    static Point2f squareToUniformDisk(const Point2f &sample);
    static float squareToUniformDiskPdf(const Point2f &p);
    static Vector3f squareToUniformSphere(const Point2f &sample);
    static float squareToUniformSpherePdf(const Vector3f &v);
    static Vector3f squareToUniformHemisphere(const Point2f &sample);
    static float squareToUniformHemispherePdf(const Vector3f &v);
    static Vector3f squareToCosineHemisphere(const Point2f &sample);
    static float squareToCosineHemispherePdf(const Vector3f &v);
    static Vector3f squareToBeckmann(const Point2f &sample, float alpha);
    static float squareToBeckmannPdf(const Vector3f &m, float alpha);
};
#include "_ref/warp_extract.inc"
}
#include <kazen/ggx_brdf.h>
#include <kazen/dpdf.h>

/* ---- stand-in hosts for the reference's method bodies --------------------------------------------------------------------------
 * The real classes hang off the plugin system (Object, PropertyList, OpenImageIO, Embree); the hosts below declare only the
 * members / records / virtual interfaces those bodies touch (include/kazen/{sampler,mesh,light,bsdf,scene}.h), and every
 * function BODY is #included from a line range of the reference's sources extracted at build time (oracle/Makefile). */
#include <kazen/ray.h>
#include <kazen/hash.h>
#include <kazen/pcg32.h>
namespace kazen {
namespace random {
#include "_ref/permute_extract.inc"
}
class Sampler { public: virtual ~Sampler() {} virtual float next1D() = 0; virtual Point2f next2D() = 0; virtual Point2f nextPixel2D() { return Point2f(0.f, 0.f); } };   /* sampler.h:44-107 */
struct ReplaySampler : Sampler { std::vector<float> q; size_t k = 0; float next1D() { return q[k++]; } Point2f next2D() { const float a = q[k++], b = q[k++]; return Point2f(a, b); } };
struct SamplerMembers : Sampler { uint64_t m_seed; uint32_t m_sampleCount, m_sampleIndex, m_dimensionIndex; };               /* sampler.h:100-106 */
struct IndependentBodies : SamplerMembers { pcg32 m_random;
#include "_ref/independent_extract.inc"
};
struct StratifiedBodies : SamplerMembers { pcg32 m_random; int m_resolution; Point2i m_pixel;
#include "_ref/stratified_extract.inc"
};
struct CorrelatedBodies : SamplerMembers { pcg32 m_random; Point2i m_resolution; Point2i m_pixel; uint32_t m_permutationSeed;
#include "_ref/correlated_extract.inc"
};

class Mesh; struct LightQueryRecord; struct BSDFQueryRecord;
struct Light {                                                                                                              /* light.h:46-66 */
    virtual ~Light() {}
    virtual Color3f eval(const LightQueryRecord &lRec) const = 0;
    virtual Color3f sample(LightQueryRecord &lRec, Sampler *sampler, const Mesh *mesh) const = 0;
    virtual float pdf(const LightQueryRecord &lRec, const Mesh *mesh) const = 0;
    virtual bool getPrimaryVisibility() const = 0;
};
struct BSDF {                                                                                                               /* bsdf.h:58-127 */
    virtual ~BSDF() {}
    virtual Color3f eval(const BSDFQueryRecord &bRec) const = 0;
    virtual float pdf(const BSDFQueryRecord &bRec) const = 0;
    virtual Color3f sample(BSDFQueryRecord &bRec, float sample1, const Point2f &sample2) const = 0;
    virtual float regularize(const Point2f &uv) const { return 0.f; }                                                       /* bsdf.h:125 */
    virtual bool isDiffuse() const { return false; }                                                                        /* bsdf.h:121 */
};
class Mesh {                                                                                                                /* mesh.h:60-187 */
public:
    MatrixXf m_V, m_N, m_UV; MatrixXu m_F; DiscretePDF *m_dpdf = nullptr; BSDF *m_bsdf = nullptr; Light *m_light = nullptr;
    const MatrixXf &getVertexPositions() const { return m_V; }
    const MatrixXf &getVertexNormals() const { return m_N; }
    const MatrixXf &getVertexTexCoords() const { return m_UV; }
    const MatrixXu &getIndices() const { return m_F; }
    bool isLight() const { return m_light != nullptr; }                                                                     /* mesh.h:133 */
    const Light *getLight() const { return m_light; }
    const BSDF *getBSDF() const { return m_bsdf; }
    float pdf() const { return m_dpdf->getNormalization(); }                                                                /* mesh.h:165-168 */
    void sample(Sampler *sampler, Point3f &p, Normal3f &n) const;
    float surfaceArea(uint32_t index) const;
};
#include "_ref/mesh_extract.inc"
struct LightQueryRecord {                                                                                                   /* light.h:10-40 */
    Point3f ref; Point2f uv; Vector3f wi; Point3f p; Normal3f n; Ray3f shadowRay; EMeasure measure; float pdf;
    LightQueryRecord(const Point3f &ref) : ref(ref) {}
    LightQueryRecord(const Point3f &ref, const Point3f &p, const Normal3f &n) : ref(ref), p(p), n(n) { wi = (p - ref).normalized(); }
};
struct AreaLightBodies : Light { Color3f m_radiance; bool m_lightPrimaryVisibility = false;
#include "_ref/light_extract.inc"
};
struct Intersection {                                                                                                       /* mesh.h:18-57 */
    Point3f p; float t; Point2f uv; Frame shFrame, geoFrame; const Mesh *mesh; Vector3f dpdu, dpdv, dndu, dndv; float accumulatedRoughness = 0.f;
    Intersection() : mesh(nullptr) {}
    Vector3f toLocal(const Vector3f &d) const { return shFrame.toLocal(d); }
    Vector3f toWorld(const Vector3f &d) const { return shFrame.toWorld(d); }
};
inline bool refPostIntersection(Intersection &its, uint32_t f) {                 /* accel.cpp:113-236 */
    bool foundIntersection = true;
#include "_ref/accel_extract.inc"
    return foundIntersection;
}
struct BSDFQueryRecord {                                                                                                    /* bsdf.h:20-53 */
    Intersection its; Vector3f wi, wo; float eta; EMeasure measure; float pdf; Point2f uv;
    BSDFQueryRecord(const Vector3f &wi) : wi(wi), eta(1.f), measure(EUnknownMeasure) {}
    BSDFQueryRecord(const Vector3f &wi, const Vector3f &wo, EMeasure measure) : wi(wi), wo(wo), eta(1.f), measure(measure) {}
};
template <typename T> struct Texture {                                                                 /* texture.h:9-16; the base doubles as the constant texture */
    T value;
    virtual ~Texture() {}
    virtual T eval(const Point2f &) const { return value; }
    virtual T eval(const Vector3f &) const { return value; }
};
struct KissBodies : BSDF {               /* KazenStandardSurface, bsdf.cpp:1175-1371 (schlickWeight ... sample) and :1397-1399 (regularize) */
    Texture<Color3f> *m_baseColor = nullptr, *m_roughness = nullptr, *m_metallic = nullptr;
    float m_anisotropy, m_specular, m_specularTint, m_sheen, m_sheenTint, m_clearcoat, m_clearcoatRoughness;
#include "_ref/kiss_extract.inc"
};
struct NormalMapBodies : BSDF {          /* NormalMap, bsdf.cpp:290-392 (eval / pdf / sample / getFrame) and :412 (regularize) */
    Texture<Color3f> *m_normalMap = nullptr; BSDF *m_nested = nullptr;
#include "_ref/normalmap_extract.inc"
};
struct DiffuseBodies : BSDF {            /* Diffuse, bsdf.cpp:27-79: eval / pdf / sample / isDiffuse */
    Color3f m_albedo;
#include "_ref/diffuse_extract.inc"
};
/* the seven other BSDF plugins, eval / pdf / sample (+ their private helpers) by line range */
struct DielectricBodies : BSDF {         /* bsdf.cpp:109-143 */
    float m_intIOR, m_extIOR;
#include "_ref/dielectric_extract.inc"
};
struct MirrorBodies : BSDF {             /* bsdf.cpp:165-191 */
#include "_ref/mirror_extract.inc"
};
struct LambertianBodies : BSDF {         /* bsdf.cpp:210-257 */
    Texture<Color3f> *m_albedo = nullptr;
#include "_ref/lambertian_extract.inc"
};
struct GGXBodies : BSDF {                /* bsdf.cpp:640-670 over ggx_brdf.h */
    Texture<Color3f> *m_albedo = nullptr; float m_roughness, m_anisotropy;
#include "_ref/ggx_extract.inc"
};
struct RoughConductorBodies : BSDF {     /* bsdf.cpp:716-794: fresnelCond, evalBeckmann, smithBeckmannG1, eval, pdf, sample */
    float m_alpha; Color3f m_eta, m_k;
#include "_ref/roughconductor_extract.inc"
};
struct RoughPlasticBodies : BSDF {       /* bsdf.cpp:843-920 */
    float m_alpha, m_intIOR, m_extIOR, m_ks; Color3f m_kd;
#include "_ref/roughplastic_extract.inc"
};
struct RoughDielectricBodies : BSDF {    /* bsdf.cpp:964-1132 */
    float m_intIOR, m_extIOR, m_eta, m_invEta, m_alpha;
#include "_ref/roughdielectric_extract.inc"
};
/* texture expression nodes, texture.cpp:114-126 (background), :160-171 (colorramp), :207-231 (blend) */
struct BackgroundTexBodies : Texture<Color3f> { float m_intensity; Texture<Color3f> *m_nested = nullptr;
#include "_ref/tex_background.inc"
};
struct ColorRampTexBodies : Texture<Color3f> { float m_min = 0.0f, m_max = 1.0f; Texture<Color3f> *m_nested = nullptr;
#include "_ref/tex_colorramp.inc"
};
struct BlendTexBodies : Texture<Color3f> { std::string m_blendmode = "mix"; Texture<Color3f> *m_mask = nullptr, *m_input1 = nullptr, *m_input2 = nullptr;
#include "_ref/tex_blend.inc"
};
}
/* Scene::getBackgroundColor (scene.cpp:54-79: null background, NaN guard, m_background->eval(dir)): the out-of-class definition is compiled
 * as a member of a stand-in class that only holds the background texture */
namespace kazen {
struct SceneBackgroundBodies { Texture<Color3f> *m_background = nullptr; const Color3f getBackgroundColor(const Vector3f &dir) const; };
#define Scene SceneBackgroundBodies
#include "_ref/scene_background.inc"
#undef Scene
}
/* PMJ02BN (sampler.cpp:273-390) over the reference's own table headers.  The tables themselves (bluenoise.cpp / pmj02table.cpp) are
 * missing from the public tree, so they are DEFINED here with synthetic contents (a scrambled (0,2)-sequence and hashed blue-noise
 * values): what is pinned is the class body -- pixel-sample bucketing, index permutation, Cranley-Patterson rotation, clamping --
 * not pbrt's numbers.  `const` is dropped from the extern declarations so the arrays can be filled at run time. */
#include <cassert>
#include <memory>
#define const
#include <kazen/pmj02table.h>
#include <kazen/bluenoise.h>
#undef const
namespace kazen {
uint32_t pmj02bnSamples[nPMJ02bnSets][nPMJ02bnSamples][2];
uint16_t BlueNoiseTextures[NumBlueNoiseTextures][BlueNoiseResolution][BlueNoiseResolution];
struct PMJ02BNBodies : SamplerMembers { Point2i m_pixel; int m_pixelTileSize; std::shared_ptr<std::vector<Point2f>> m_pixelSamples;
    void construct() {                   /* the constructor after the PropertyList reads, sampler.cpp:284-309 */
#undef assert
#define assert(x) ((void)0)              /* the synthetic table need not be a perfect (0,2) net for non power-of-4 counts */
#include "_ref/pmj02bn_ctor.inc"
    }
#include "_ref/pmj02bn_extract.inc"
#undef assert
};
}
#include <cassert>
#include <kazen/transform.h>
namespace kazen {
struct PerspectiveBodies {               /* PerspectiveCamera::sampleRay, camera.cpp:70-91; members of camera.cpp:93-103 */
    Vector2f m_invOutputSize; Transform m_sampleToCamera, m_cameraToWorld; float m_nearClip, m_farClip;
#include "_ref/perspective_extract.inc"
};
struct ThinlensBodies {                  /* ThinlensCamera::sampleRay, camera.cpp:191-223 */
    Vector2f m_invOutputSize; Transform m_sampleToCamera, m_cameraToWorld; float m_nearClip, m_farClip, m_apertureRadius, m_focusDistance;
#include "_ref/thinlens_extract.inc"
};
}
#include <kazen/bbox.h>
#ifndef KAZEN_FILTER_RESOLUTION
#define KAZEN_FILTER_RESOLUTION 32       /* rfilter.h:6 */
#endif
namespace kazen {
class ReconstructionFilter { public: virtual ~ReconstructionFilter() {} float getRadius() const { return m_radius; } virtual float eval(float x) const = 0; float m_radius; };   /* rfilter.h */
struct GaussianBodies : ReconstructionFilter { float m_stddev;
#include "_ref/rfilter_gaussian.inc"
};
struct MitchellBodies : ReconstructionFilter { float m_B, m_C;
#include "_ref/rfilter_mitchell.inc"
};
struct TentBodies : ReconstructionFilter {
#include "_ref/rfilter_tent.inc"
};
struct BoxBodies : ReconstructionFilter {
#include "_ref/rfilter_box.inc"
};
class ImageBlock {                       /* block.h; the constructor (block.cpp:9-31) and put (block.cpp:56-85) are the reference's */
public:
    ImageBlock(const Vector2i &size, const ReconstructionFilter *filter);
    void put(const Point2f &_pos, const Color3f &value);
    void resize(int rows, int cols) { m_rows = rows; m_cols = cols; m_px.assign((size_t)rows * cols, Color4f()); }
    Color4f &coeffRef(int y, int x) { return m_px[(size_t)y * m_cols + x]; }
    int cols() const { return m_cols; } int rows() const { return m_rows; }
    Point2i m_offset; Vector2i m_size; int m_borderSize = 0; float *m_filter = nullptr; float m_filterRadius = 0; float *m_weightsX = nullptr, *m_weightsY = nullptr; float m_lookupFactor = 0;
    std::vector<Color4f> m_px; int m_rows = 0, m_cols = 0;
};
#include "_ref/block_extract.inc"
}
struct RefAccel;                          /* closest hit through the oracle's intersector (Embree's stand-in), defined after kzo.cpp */
namespace kazen {
class Integrator; class Camera;
class Scene {                                                                                                               /* scene.h:15-138 */
public:
    std::vector<Mesh *> m_meshes, m_lights; const RefAccel *m_accel = nullptr; Color3f m_background;
    const Integrator *m_integrator = nullptr; const Camera *m_camera = nullptr;
    const Integrator *getIntegrator() const { return m_integrator; }                  /* scene.h:30 */
    const Camera *getCamera() const { return m_camera; }                              /* scene.h:39 */
#include "_ref/scene_extract.inc"
    bool rayIntersect(const Ray3f &ray, Intersection &its) const;                 /* scene.h:79-81  -> Accel::rayIntersect(ray, its, false) */
    bool rayOccluded(const Ray3f &ray, Intersection &its) const;                  /* scene.h:103-105 -> Accel::rayIntersect(ray, its, true) */
    bool rayIntersect(const Ray3f &ray) const { Intersection its; return rayOccluded(ray, its); }                           /* scene.h:92-95 */
    Color3f getBackgroundColor(const Vector3f &) const { return m_background; }   /* constant background */
};
struct NormalIntegratorBodies {          /* integrator.cpp:18-29 */
#include "_ref/integ_normals.inc"
};
struct AoIntegratorBodies {              /* integrator.cpp:43-62 */
#include "_ref/integ_ao.inc"
};
struct WhittedIntegratorBodies {         /* integrator.cpp:78-129 (recursive) */
#include "_ref/integ_whitted.inc"
};
#define override
struct PathMatsIntegratorBodies {        /* integrator.cpp:141-175 */
#include "_ref/integ_pathmats.inc"
};
#undef override
struct PathMisBodies {                   /* PathMisIntegrator, integrator.cpp:195-344: Li and powerHeuristic */
    int m_maxDepth; float m_rayEpsilon; bool m_regularization; float m_accumulatedRoughness;
#include "_ref/integrator_extract.inc"
};
/* the 8-bit conversion of Bitmap::savePNG (bitmap.cpp:46-54: toSRGB, scale, clamp, truncate), one pixel at a time */
static void refQuantise(Color3f &cur, uint8_t *dst) {
    auto coeffRef = [&](int, int) -> Color3f & { return cur; };
    const int i = 0, j = 0;
#include "_ref/bitmap_quantise.inc"
}
/* renderer::renderSample (renderer.cpp:18-38): pixel sample, aperture sample, camera ray, Li, ImageBlock::put -- over the interfaces it
 * calls (integrator.h:26, camera.h:32), implemented by the hosted bodies above */
class Integrator { public: virtual ~Integrator() {} virtual Color3f Li(const Scene *scene, Sampler *sampler, const Ray3f &ray) const = 0; };
class Camera { public: virtual ~Camera() {} virtual Color3f sampleRay(Ray3f &ray, const Point2f &samplePosition, const Point2f &apertureSample) const = 0; };
struct HostedPathMis : Integrator { PathMisBodies b; Color3f Li(const Scene *s, Sampler *sm, const Ray3f &r) const { return b.Li(s, sm, r); } };
struct HostedPerspective : Camera { PerspectiveBodies b; Color3f sampleRay(Ray3f &r, const Point2f &p, const Point2f &a) const { return b.sampleRay(r, p, a); } };
struct HostedThinlens : Camera { ThinlensBodies b; Color3f sampleRay(Ray3f &r, const Point2f &p, const Point2f &a) const { return b.sampleRay(r, p, a); } };
namespace renderer {
#include "_ref/render_sample.inc"
}
}

#include "kzo.cpp"          /* the oracle itself: SceneData, Sampler, the restated Li (file-static) */

/* Accel::rayIntersect (accel.cpp:63-110) with rtcIntersect1 replaced by the oracle's closest-hit query over the same triangles;
 * the bookkeeping around it follows :99-110, the post-intersection is the reference's own block. */
struct RefAccel { const kzo::Accel *accel; const std::vector<kazen::Mesh *> *meshes; };
static bool refAccelIntersect(const RefAccel &a, const kazen::Ray3f &ray, kazen::Intersection &its, bool shadowRay) {
    const kz_ray r{{ray.o.x(), ray.o.y(), ray.o.z()}, ray.mint, {ray.d.x(), ray.d.y(), ray.d.z()}, ray.maxt};
    const kzo::HitRec h = a.accel->traceBvh(r);
    if (h.geom == KZ_INVALID_ID) return false;
    if (shadowRay) { its.t = h.t; its.mesh = (*a.meshes)[h.geom]; return true; }
    its.t = h.t; its.uv = kazen::Point2f(h.u, h.v); its.mesh = (*a.meshes)[h.geom];
    return kazen::refPostIntersection(its, h.prim);
}
bool kazen::Scene::rayIntersect(const Ray3f &ray, Intersection &its) const { return refAccelIntersect(*m_accel, ray, its, false); }
bool kazen::Scene::rayOccluded(const Ray3f &ray, Intersection &its) const { return refAccelIntersect(*m_accel, ray, its, true); }

static uint64_t g_rng = 0x9E3779B97F4A7C15ull;
static float rnd() { g_rng ^= g_rng << 13; g_rng ^= g_rng >> 7; g_rng ^= g_rng << 17; return (float)((g_rng >> 40) * (1.0 / 16777216.0)); }
static float rnd(float a, float b) { return a + (b - a) * rnd(); }
static kzo::V3 rdir(bool upper) { for (;;) { kzo::V3 v(rnd(-1, 1), rnd(-1, 1), upper ? rnd(0.02f, 1) : rnd(-1, 1)); float n = kzo::norm(v); if (n > 0.05f && n <= 1.f) return v / n; } }
static kazen::Vector3f K(kzo::V3 v) { return kazen::Vector3f(v.x, v.y, v.z); }
static kazen::Color3f KC(kzo::V3 v) { return kazen::Color3f(v.x, v.y, v.z); }

struct Out { std::string json; long cases = 0, bad = 0; };
static Out g;
static bool same(float a, float b) { uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4); return x == y || (a != a && b != b); }
static void rec(const char *fn, std::vector<float> in, std::vector<float> ref, std::vector<float> ours, bool keep) {
    ++g.cases;
    bool ok = ref.size() == ours.size();
    for (size_t i = 0; ok && i < ref.size(); ++i) ok = same(ref[i], ours[i]);
    if (!ok) { if (++g.bad <= 12) { fprintf(stderr, "MISMATCH %s:", fn); for (float v : in) fprintf(stderr, " %.9g", v); fprintf(stderr, " -> ref"); for (float v : ref) fprintf(stderr, " %.9g", v);
                                    fprintf(stderr, " ours"); for (float v : ours) fprintf(stderr, " %.9g", v); fprintf(stderr, "\n"); } }
    if (!keep) return;
    auto hex = [](std::vector<float> &v) { std::string s = "["; for (size_t i = 0; i < v.size(); ++i) { uint32_t u; memcpy(&u, &v[i], 4); char b[16]; snprintf(b, sizeof b, "%s%u", i ? "," : "", u); s += b; } return s + "]"; };
    g.json += std::string(g.json.empty() ? "" : ",\n") + "  {\"fn\":\"" + fn + "\",\"in\":" + hex(in) + ",\"out\":" + hex(ref) + "}";
}
static std::vector<float> f3(kzo::V3 v) { return {v.x, v.y, v.z}; }
static std::vector<float> f3(const kazen::Vector3f &v) { return {v.x(), v.y(), v.z()}; }
static std::vector<float> f3(const kazen::Color3f &v) { return {v.x(), v.y(), v.z()}; }

int main() {
    const int N = 4000, KEEP = 40;
    for (int i = 0; i < N; ++i) {
        const bool keep = i < KEEP;
        const float rough = i % 7 == 0 ? 0.f : rnd(), aniso = i % 3 == 0 ? 0.f : rnd(-0.9f, 0.9f);
        const kzo::V3 V = rdir(true), L = rdir(i % 5 != 0), H = kzo::normalized(V + L), f0(rnd(), rnd(), rnd());
        const kzo::V2 u{rnd(), rnd()};
        const kazen::Vector2f ka = kazen::roughnessToAlpha(rough, aniso); const kzo::V2 oa = kzo::roughnessToAlpha(rough, aniso);
        rec("roughnessToAlpha", {rough, aniso}, {ka.x(), ka.y()}, {oa.x, oa.y}, keep);
        rec("lambda", {V.x, V.y, V.z, oa.x, oa.y}, {kazen::lambda(K(V), ka)}, {kzo::ggxLambda(V, oa)}, keep);
        rec("smithG1", {V.x, V.y, V.z, H.x, H.y, H.z, oa.x, oa.y}, {kazen::evaluateSmithG1(K(V), K(H), ka)}, {kzo::smithG1(V, H, oa)}, keep);
        rec("smithG2", {V.x, V.y, V.z, L.x, L.y, L.z, H.x, H.y, H.z, oa.x, oa.y}, {kazen::evaluateSmithG2(K(V), K(L), K(H), ka)}, {kzo::smithG2(V, L, H, oa)}, keep);
        rec("ggxNDF", {H.x, H.y, H.z, oa.x, oa.y}, {kazen::evaluateGGXNDF(K(H), ka)}, {kzo::ggxNDF(H, oa)}, keep);
        rec("ggxVNDF", {V.x, V.y, V.z, H.x, H.y, H.z, oa.x, oa.y}, {kazen::evaluateGGXSmithVNDF(K(V), K(H), ka)}, {kzo::ggxVNDF(V, H, oa)}, keep);
        rec("sampleVNDF", {V.x, V.y, V.z, oa.x, oa.y, u.x, u.y}, f3(kazen::sampleGGXSmithVNDF(K(V), ka, kazen::Point2f(u.x, u.y))), f3(kzo::sampleGGXVNDF(V, oa, u)), keep);
        rec("schlick", {f0.x, f0.y, f0.z, u.x}, f3(kazen::evaluateSchlickFresnel(KC(f0), u.x)), f3(kzo::schlickFresnel(f0, u.x)), keep);
        { kazen::Color3f F; rec("ggxSmithBRDF", {V.x, V.y, V.z, L.x, L.y, L.z, f0.x, f0.y, f0.z, rough, aniso},
              f3(kazen::evaluateGGXSmithBRDF(K(V), K(L), KC(f0), rough, aniso, F)), f3(kzo::ggxSmithBRDF(V, L, f0, rough, aniso)), keep); }
        const kzo::V3 n = rdir(false), w = rdir(false);
        { kazen::Vector3f kb, kc; kazen::coordinateSystem(K(n), kb, kc); kzo::V3 ob, oc; kzo::coordinateSystem(n, ob, oc);
          rec("coordinateSystem", f3(n), {kb.x(), kb.y(), kb.z(), kc.x(), kc.y(), kc.z()}, {ob.x, ob.y, ob.z, oc.x, oc.y, oc.z}, keep);
          kazen::Frame kf((kazen::Normal3f)K(n)); kzo::Frame of(n);
          rec("frameToLocal", {n.x, n.y, n.z, w.x, w.y, w.z}, f3(kf.toLocal(K(w))), f3(of.toLocal(w)), keep);
          rec("frameToWorld", {n.x, n.y, n.z, w.x, w.y, w.z}, f3(kf.toWorld(K(w))), f3(of.toWorld(w)), keep); }
        rec("reflect", {w.x, w.y, w.z, n.x, n.y, n.z}, f3(kazen::reflect(K(w), K(n))), f3(kzo::reflect(w, n)), keep);
        const float eta = i % 4 == 0 ? 1.f / 1.5046f : rnd(1.05f, 2.4f), c = rnd(-1, 1);
        rec("refract", {w.x, w.y, w.z, n.x, n.y, n.z, eta}, f3(kazen::refract(K(w), K(n), eta)), f3(kzo::refractDir(w, n, eta)), keep);
        rec("fresnel", {c, 1.000277f, eta}, {kazen::fresnel(c, 1.000277f, eta)}, {kzo::fresnelExtInt(c, 1.000277f, eta)}, keep);
        { float kt = 0, ot = 0; const float kr = kazen::fresnelDielectric(c, eta, kt), orr = kzo::fresnelDielectric(c, eta, ot);
          rec("fresnelDielectric", {c, eta}, {kr, kt}, {orr, ot}, keep); }
        { const kazen::Point2f kd = kazen::Warp::squareToUniformDisk(kazen::Point2f(u.x, u.y)); const kzo::V2 od = kzo::squareToUniformDisk(u);
          rec("squareToUniformDisk", {u.x, u.y}, {kd.x(), kd.y()}, {od.x, od.y}, keep); }
        rec("squareToCosineHemisphere", {u.x, u.y}, f3(kazen::Warp::squareToCosineHemisphere(kazen::Point2f(u.x, u.y))), f3(kzo::squareToCosineHemisphere(u)), keep);
        const float alpha = rnd(0.02f, 0.9f);
        rec("squareToBeckmann", {u.x, u.y, alpha}, f3(kazen::Warp::squareToBeckmann(kazen::Point2f(u.x, u.y), alpha)), f3(kzo::squareToBeckmann(u, alpha)), keep);
        rec("squareToBeckmannPdf", {V.x, V.y, V.z, alpha}, {kazen::Warp::squareToBeckmannPdf(K(V), alpha)}, {kzo::squareToBeckmannPdf(V, alpha)}, keep);
    }
    /* kiss (KazenStandardSurface) and Diffuse: eval / pdf / sample with constant textures, incl. the accumulated-roughness bias */
    for (int i = 0; i < N; ++i) {
        const bool keep = i < KEEP;
        kazen::Texture<kazen::Color3f> tb, tr, tm;
        const kzo::V3 base(rnd(), rnd(), rnd());
        const float rough = i % 9 == 0 ? 0.f : rnd(), metal = i % 4 == 0 ? 0.f : (i % 4 == 1 ? 1.f : rnd());
        tb.value = KC(base); tr.value = kazen::Color3f(rough); tm.value = kazen::Color3f(metal);
        kazen::KissBodies kb; kb.m_baseColor = &tb; kb.m_roughness = &tr; kb.m_metallic = &tm;
        kb.m_anisotropy = i % 3 == 0 ? 0.f : rnd(-0.8f, 0.8f); kb.m_specular = rnd(); kb.m_specularTint = rnd();
        kb.m_sheen = i % 2 ? rnd() : 0.f; kb.m_sheenTint = rnd(); kb.m_clearcoat = i % 2 ? 0.f : rnd(); kb.m_clearcoatRoughness = rnd();
        kzo::SceneData sc;
        kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
        t.color[0] = base.x; t.color[1] = base.y; t.color[2] = base.z; sc.textures.push_back(t);
        t.color[0] = t.color[1] = t.color[2] = rough; sc.textures.push_back(t);
        t.color[0] = t.color[1] = t.color[2] = metal; sc.textures.push_back(t);
        kz_bsdf_desc m; memset(&m, 0, sizeof(m)); m.type = KZ_BSDF_KISS; m.base_color = 0; m.roughness = 1; m.metallic = 2;
        m.anisotropy = kb.m_anisotropy; m.specular = kb.m_specular; m.specular_tint = kb.m_specularTint; m.sheen = kb.m_sheen; m.sheen_tint = kb.m_sheenTint;
        m.clearcoat = kb.m_clearcoat; m.clearcoat_roughness = kb.m_clearcoatRoughness;
        const kzo::V3 wi = rdir(i % 11 != 0), wo = rdir(i % 13 != 0);
        const float acc = i % 3 == 1 ? rnd(0.f, 0.6f) : 0.f, s1 = rnd(); const kzo::V2 s2{rnd(), rnd()};
        std::vector<float> in = {base.x, base.y, base.z, rough, metal, m.anisotropy, m.specular, m.specular_tint, m.sheen, m.sheen_tint, m.clearcoat, m.clearcoat_roughness,
                                 wi.x, wi.y, wi.z, wo.x, wo.y, wo.z, acc, s1, s2.x, s2.y};
        kazen::BSDFQueryRecord kr(K(wi), K(wo), kazen::ESolidAngle); kr.its.accumulatedRoughness = acc; kr.uv = kazen::Point2f(0.5f, 0.5f);
        kzo::BSDFQueryRecord orr(wi, wo, kzo::ESolidAngle); orr.its.accumulatedRoughness = acc;
        rec("kissEval", in, f3(kb.eval(kr)), f3(kzo::kissEval(sc, m, orr)), keep);
        rec("kissPdf", in, {kb.pdf(kr)}, {kzo::kissPdf(sc, m, orr)}, keep);
        kazen::BSDFQueryRecord ks(K(wi)); ks.its.accumulatedRoughness = acc; ks.uv = kazen::Point2f(0.5f, 0.5f);
        kzo::BSDFQueryRecord os(wi); os.its.accumulatedRoughness = acc;
        const kazen::Color3f kw = kb.sample(ks, s1, kazen::Point2f(s2.x, s2.y)); const kzo::V3 ow = kzo::kissSample(sc, m, os, s1, s2);
        const bool kz0 = kw.x() == 0.f && kw.y() == 0.f && kw.z() == 0.f;      /* wo is only meaningful for a non-zero weight */
        rec("kissSample", in, {kw.x(), kw.y(), kw.z(), kz0 ? 0.f : ks.wo.x(), kz0 ? 0.f : ks.wo.y(), kz0 ? 0.f : ks.wo.z()},
            {ow.x, ow.y, ow.z, kz0 ? 0.f : os.wo.x, kz0 ? 0.f : os.wo.y, kz0 ? 0.f : os.wo.z}, keep);
        kazen::DiffuseBodies db; db.m_albedo = KC(base);
        kz_bsdf_desc dm; memset(&dm, 0, sizeof(dm)); dm.type = KZ_BSDF_DIFFUSE; dm.albedo[0] = base.x; dm.albedo[1] = base.y; dm.albedo[2] = base.z;
        sc.bsdfs.push_back(dm);
        std::vector<float> din = {base.x, base.y, base.z, wi.x, wi.y, wi.z, wo.x, wo.y, wo.z, s2.x, s2.y};
        rec("diffuseEval", din, f3(db.eval(kr)), f3(kzo::bsdfEval(sc, 0, orr)), keep);
        rec("diffusePdf", din, {db.pdf(kr)}, {kzo::bsdfPdf(sc, 0, orr)}, keep);
        kazen::BSDFQueryRecord kd(K(wi)); kzo::BSDFQueryRecord od(wi);
        const kazen::Color3f dw = db.sample(kd, s1, kazen::Point2f(s2.x, s2.y)); const kzo::V3 odw = kzo::bsdfSample(sc, 0, od, s1, s2);
        const bool dz0 = dw.x() == 0.f && dw.y() == 0.f && dw.z() == 0.f;
        rec("diffuseSample", din, {dw.x(), dw.y(), dw.z(), dz0 ? 0.f : kd.wo.x(), dz0 ? 0.f : kd.wo.y(), dz0 ? 0.f : kd.wo.z()},
            {odw.x, odw.y, odw.z, dz0 ? 0.f : od.wo.x, dz0 ? 0.f : od.wo.y, dz0 ? 0.f : od.wo.z}, keep);
        /* normal map over the kiss material: random shading frame + dpdu, random tangent-space normal (bsdf.cpp:290-392) */
        {
            kazen::Texture<kazen::Color3f> tn; const kzo::V3 nrgb = i % 6 == 0 ? kzo::V3(0.5f, 0.5f, 1.f) : kzo::V3(rnd(0.2f, 0.8f), rnd(0.2f, 0.8f), rnd(0.55f, 1.f));
            tn.value = KC(nrgb);
            kazen::NormalMapBodies nb; nb.m_normalMap = &tn; nb.m_nested = &kb;
            kz_texture_desc t2; memset(&t2, 0, sizeof(t2)); t2.type = KZ_TEX_CONSTANT; t2.child[0] = t2.child[1] = t2.child[2] = -1; t2.color[0] = nrgb.x; t2.color[1] = nrgb.y; t2.color[2] = nrgb.z;
            sc.textures.push_back(t2);
            sc.bsdfs.clear(); sc.bsdfs.push_back(m);
            kz_bsdf_desc nm; memset(&nm, 0, sizeof(nm)); nm.type = KZ_BSDF_NORMALMAP; nm.normal_map = 3; nm.nested = 0; sc.bsdfs.push_back(nm);
            const kzo::V3 shn = rdir(false), dpdu = rdir(false) * rnd(0.3f, 2.f);
            kazen::Intersection kit; kit.shFrame = kazen::Frame((kazen::Normal3f)K(shn)); kit.geoFrame = kit.shFrame; kit.dpdu = K(dpdu); kit.uv = kazen::Point2f(0.5f, 0.5f); kit.accumulatedRoughness = acc;
            kzo::Intersection oit; oit.shFrame = kzo::Frame(shn); oit.geoFrame = oit.shFrame; oit.dpdu = dpdu; oit.uv = kzo::V2{0.5f, 0.5f}; oit.accumulatedRoughness = acc;
            std::vector<float> nin = in; nin.insert(nin.end(), {nrgb.x, nrgb.y, nrgb.z, shn.x, shn.y, shn.z, dpdu.x, dpdu.y, dpdu.z});
            kazen::BSDFQueryRecord nr(K(wi), K(wo), kazen::ESolidAngle); nr.its = kit; nr.uv = kit.uv;
            kzo::BSDFQueryRecord onr(wi, wo, kzo::ESolidAngle); onr.its = oit; onr.uv = oit.uv;
            rec("normalMapEval", nin, f3(nb.eval(nr)), f3(kzo::bsdfEval(sc, 1, onr)), false);
            rec("normalMapPdf", nin, {nb.pdf(nr)}, {kzo::bsdfPdf(sc, 1, onr)}, false);
            kazen::BSDFQueryRecord ns(K(wi)); ns.its = kit; ns.uv = kit.uv; kzo::BSDFQueryRecord ons(wi); ons.its = oit; ons.uv = oit.uv;
            const kazen::Color3f nw = nb.sample(ns, s1, kazen::Point2f(s2.x, s2.y)); const kzo::V3 onw = kzo::bsdfSample(sc, 1, ons, s1, s2);
            const bool nz0 = nw.x() == 0.f && nw.y() == 0.f && nw.z() == 0.f;
            rec("normalMapSample", nin, {nw.x(), nw.y(), nw.z(), nz0 ? 0.f : ns.wo.x(), nz0 ? 0.f : ns.wo.y(), nz0 ? 0.f : ns.wo.z()},
                {onw.x, onw.y, onw.z, nz0 ? 0.f : ons.wo.x, nz0 ? 0.f : ons.wo.y, nz0 ? 0.f : ons.wo.z}, false);
            sc.bsdfs.clear(); sc.textures.pop_back();
        }
        /* Color3f helpers, common.cpp:352-395 */
        const kzo::V3 c(rnd(0.f, 1.4f), rnd(0.f, 0.01f), rnd());
        rec("toSRGB", f3(c), f3(KC(c).toSRGB()), f3(kzo::toSRGB(c)), keep);
        rec("toLinearRGB", f3(c), f3(KC(c).toLinearRGB()), f3(kzo::toLinearRGB(c)), keep);
        rec("luminance", f3(c), {KC(c).getLuminance()}, {kzo::luminance(c)}, keep);
    }
    /* meshes: random small meshes with / without normals and texture coordinates; post-intersection at random barycentrics,
     * Mesh::sample + AreaLight::sample / pdf / eval with the same three random numbers on both sides */
    for (int t = 0; t < 1500; ++t) {
        const int nV = 3 + (int)(rnd() * 9), nF = 1 + (int)(rnd() * 8);
        const bool hasN = t % 3 != 2, hasUV = t % 3 == 0, degenerateUV = hasUV && t % 12 == 0;
        kazen::Mesh km; kzo::SceneData sc; sc.meshes.emplace_back(); kzo::MeshData &om = sc.meshes[0];
        km.m_V.resize(3, nV); if (hasN) km.m_N.resize(3, nV); if (hasUV) km.m_UV.resize(2, nV); km.m_F.resize(3, nF);
        om.nV = (uint32_t)nV; om.nF = (uint32_t)nF;
        for (int i = 0; i < nV; ++i) {
            const kzo::V3 pp(rnd(-2, 2), rnd(-2, 2), rnd(-2, 2)); kzo::V3 nn = rdir(false); if (t % 5 == 0) nn = nn * rnd(0.5f, 1.5f);      /* also unnormalised normals */
            const kzo::V2 uv{degenerateUV ? 0.25f : rnd(), degenerateUV ? 0.5f : rnd()};
            km.m_V(0, i) = pp.x; km.m_V(1, i) = pp.y; km.m_V(2, i) = pp.z; om.P.insert(om.P.end(), {pp.x, pp.y, pp.z});
            if (hasN) { km.m_N(0, i) = nn.x; km.m_N(1, i) = nn.y; km.m_N(2, i) = nn.z; om.N.insert(om.N.end(), {nn.x, nn.y, nn.z}); }
            if (hasUV) { km.m_UV(0, i) = uv.x; km.m_UV(1, i) = uv.y; om.UV.insert(om.UV.end(), {uv.x, uv.y}); }
        }
        for (int f = 0; f < nF; ++f) for (int k = 0; k < 3; ++k) { const uint32_t idx = (uint32_t)((f + k * (1 + f % 2) + (int)(rnd() * nV)) % nV); km.m_F(k, f) = idx; om.F.push_back(idx); }
        for (int q = 0; q < 8; ++q) {
            const uint32_t f = (uint32_t)(rnd() * nF) % (uint32_t)nF; float bu = rnd(), bv = rnd(); if (bu + bv > 1) { bu = 1 - bu; bv = 1 - bv; }
            kazen::Intersection ki; ki.mesh = &km; ki.uv = kazen::Point2f(bu, bv); ki.t = 1.f;
            ki.dpdu = kazen::Vector3f(0.f); ki.dpdv = kazen::Vector3f(0.f);
            kazen::refPostIntersection(ki, f);
            kzo::Intersection oi; kzo::HitRec h{1.f, bu, bv, f, 0u};
            kzo::fillIntersection(sc, h, oi);
            std::vector<float> in = {(float)t, (float)f, bu, bv};
            rec("postIntersection", in, {ki.p.x(), ki.p.y(), ki.p.z(), ki.uv.x(), ki.uv.y(), ki.geoFrame.n.x(), ki.geoFrame.n.y(), ki.geoFrame.n.z(),
                                          ki.shFrame.s.x(), ki.shFrame.s.y(), ki.shFrame.s.z(), ki.shFrame.t.x(), ki.shFrame.t.y(), ki.shFrame.t.z(), ki.shFrame.n.x(), ki.shFrame.n.y(), ki.shFrame.n.z()},
                {oi.p.x, oi.p.y, oi.p.z, oi.uv.x, oi.uv.y, oi.geoFrame.n.x, oi.geoFrame.n.y, oi.geoFrame.n.z,
                 oi.shFrame.s.x, oi.shFrame.s.y, oi.shFrame.s.z, oi.shFrame.t.x, oi.shFrame.t.y, oi.shFrame.t.z, oi.shFrame.n.x, oi.shFrame.n.y, oi.shFrame.n.z}, false);
        }
        /* the mesh as an emitter: Mesh::activate (mesh.cpp:31-43) builds the area CDF */
        kazen::DiscretePDF dp((size_t)nF); dp.reserve((size_t)nF);
        float area = 0.f;
        for (int i = 0; i < nF; ++i) { const float a = km.surfaceArea((uint32_t)i); dp.append(a); area += a; }
        dp.normalize(); km.m_dpdf = &dp;
        kzo::buildLightCdf(om);
        rec("lightCdfNormalization", {(float)t}, {km.pdf()}, {om.normalization}, false);
        if (!(area > 0.f)) continue;
        kazen::AreaLightBodies kl; kl.m_radiance = kazen::Color3f(3.f, 2.f, 1.f);
        kz_light_desc ol; ol.radiance[0] = 3.f; ol.radiance[1] = 2.f; ol.radiance[2] = 1.f; ol.primary_visibility = 0;
        for (int q = 0; q < 6; ++q) {
            const kzo::V3 ref(rnd(-3, 3), rnd(-3, 3), rnd(-3, 3));
            kazen::ReplaySampler ks; ks.q = {rnd(), rnd(), rnd()};
            kzo::SamplerCfg cfg; memset(&cfg.d, 0, sizeof(cfg.d)); cfg.d.type = KZ_SAMPLER_INDEPENDENT; cfg.d.sample_count = 1; cfg.d.seed = (uint64_t)(t * 8 + q + 1);
            kzo::Sampler osm; osm.cfg = &cfg; osm.generateSample(t, q, 0);
            { kzo::Sampler peek = osm; ks.q = {peek.next1D(), peek.next1D(), peek.next1D()}; }
            kazen::LightQueryRecord klr(kazen::Point3f(ref.x, ref.y, ref.z)); kzo::LightQueryRecord olr(ref);
            const kazen::Color3f kw = kl.sample(klr, &ks, &km); const kzo::V3 ow = kzo::lightSample(ol, om, olr, osm);
            rec("areaLightSample", {(float)t, ref.x, ref.y, ref.z, ks.q[0], ks.q[1], ks.q[2]},
                {kw.x(), kw.y(), kw.z(), klr.p.x(), klr.p.y(), klr.p.z(), klr.n.x(), klr.n.y(), klr.n.z(), klr.wi.x(), klr.wi.y(), klr.wi.z(), klr.pdf, klr.shadowRay.maxt},
                {ow.x, ow.y, ow.z, olr.p.x, olr.p.y, olr.p.z, olr.n.x, olr.n.y, olr.n.z, olr.wi.x, olr.wi.y, olr.wi.z, olr.pdf, olr.shadowRay.tmax}, false);
            const kazen::LightQueryRecord k2(kazen::Point3f(ref.x, ref.y, ref.z), klr.p, klr.n); const kzo::LightQueryRecord o2(ref, olr.p, olr.n);
            rec("areaLightPdfEval", {(float)t, ref.x, ref.y, ref.z}, {kl.pdf(k2, &km), kl.eval(k2).x(), kl.eval(k2).y(), kl.eval(k2).z()},
                {kzo::lightPdf(om, o2), kzo::lightEval(ol, o2).x, kzo::lightEval(ol, o2).y, kzo::lightEval(ol, o2).z}, false);
        }
    }
    /* ---- the integrator loop itself: PathMisIntegrator::Li (integrator.cpp:195-338) on random small scenes.  Reference side:
     *      the Li body over the hosted Scene / Mesh / AreaLight / kiss / diffuse / Stratified bodies; oracle side: the restated Li of
     *      kzo.cpp over SceneData.  Both see the same triangles (closest hits come from the oracle's intersector on both sides, as
     *      Embree's stand-in), the same sampler configuration and the same camera rays; the radiance must agree bit for bit. ---- */
    long liPaths = 0, liLit = 0;
    for (int scn = 0; scn < 24; ++scn) {
        struct Spec { std::vector<kzo::V3> P, N; std::vector<kzo::V2> UV; std::vector<uint32_t> F; int kind; /* 0 diffuse, 1 kiss, 2 normal map over kiss */ int light; bool visible; };
        std::vector<Spec> specs;
        auto quad = [&](kzo::V3 a, kzo::V3 b, kzo::V3 c, kzo::V3 d, bool withN, bool withUV, int kind, int light, bool visible) {
            Spec m; m.P = {a, b, c, d}; m.F = {0, 1, 2, 0, 2, 3}; m.kind = kind; m.light = light; m.visible = visible;
            if (withN) { const kzo::V3 n = kzo::normalized(kzo::cross(b - a, c - a)); m.N = {n, n, n, n}; }
            if (withUV) m.UV = {kzo::V2{0, 0}, kzo::V2{1, 0}, kzo::V2{1, 1}, kzo::V2{0, 1}};
            specs.push_back(m);
        };
        quad(kzo::V3(-3, 0, 3), kzo::V3(3, 0, 3), kzo::V3(3, 0, -3), kzo::V3(-3, 0, -3), true, true, scn % 2, -1, false);                 /* floor */
        quad(kzo::V3(-3, 0, -3), kzo::V3(3, 0, -3), kzo::V3(3, 4, -3), kzo::V3(-3, 4, -3), scn % 3 != 0, scn % 3 == 1, scn % 3 == 1 ? 2 : 1, -1, false);      /* back wall; kind 2 = normal map over kiss */
        for (int k = 0; k < 2; ++k) {                                                                                                      /* two triangle clusters */
            Spec m; m.kind = k == 0 ? 1 : 0; m.light = -1; m.visible = false;
            const bool withN = k == 0 || scn % 4 == 0, withUV = k == 0 && scn % 2 == 0;
            for (int t = 0; t < 6; ++t) {
                const kzo::V3 c(rnd(-2, 2), rnd(0.3f, 2.2f), rnd(-2, 1.5f));
                kzo::V3 p[3]; for (int v = 0; v < 3; ++v) p[v] = c + kzo::V3(rnd(-0.7f, 0.7f), rnd(-0.5f, 0.5f), rnd(-0.7f, 0.7f));
                const kzo::V3 gn = kzo::normalized(kzo::cross(p[1] - p[0], p[2] - p[0]));
                for (int v = 0; v < 3; ++v) {
                    m.F.push_back((uint32_t)m.P.size()); m.P.push_back(p[v]);
                    if (withN) m.N.push_back(kzo::normalized(gn + kzo::V3(rnd(-0.35f, 0.35f), rnd(-0.35f, 0.35f), rnd(-0.35f, 0.35f))));   /* smooth-shading normals */
                    if (withUV) m.UV.push_back(kzo::V2{rnd(), rnd()});
                }
            }
            specs.push_back(m);
        }
        quad(kzo::V3(-1, 3.6f, -1), kzo::V3(1, 3.6f, -1), kzo::V3(1, 3.6f, 1), kzo::V3(-1, 3.6f, 1), scn % 2 == 0, false, 0, 0, false);     /* ceiling light, invisible, faces down */
        quad(kzo::V3(2.9f, 0.5f, 1), kzo::V3(2.9f, 0.5f, -1), kzo::V3(2.9f, 1.5f, -1), kzo::V3(2.9f, 1.5f, 1), false, false, 0, 1, scn % 3 == 0);   /* side light, faces -x */
        if (scn % 4 == 1) quad(kzo::V3(-0.6f, 1.0f, 2.0f), kzo::V3(0.6f, 1.0f, 2.0f), kzo::V3(0.6f, 2.0f, 2.0f), kzo::V3(-0.6f, 2.0f, 2.0f), false, false, 0, 2, false);   /* invisible panel in front of the camera */

        const bool regularize = scn % 2 == 1; const int maxDepth = scn % 5 == 0 ? 8 : 5; const float eps = 1e-3f;
        const kzo::V3 bg = scn % 3 == 2 ? kzo::V3(0.f) : kzo::V3(0.2f, 0.3f, 0.4f);
        /* oracle scene */
        kzo_scene os; kzo::SceneData &sc = os.sc;
        sc.integrator.max_depth = maxDepth; sc.integrator.trace_bias = eps; sc.integrator.regularization = regularize; sc.integrator.accumulated_roughness = 0.5f; sc.integrator.type = KZ_INTEGRATOR_PATH_MIS;
        memset(&sc.sampler.d, 0, sizeof(sc.sampler.d)); sc.sampler.d.type = KZ_SAMPLER_STRATIFIED; sc.sampler.d.sample_count = 16; sc.sampler.d.seed = (uint64_t)(scn + 1); sc.sampler.d.res_x = sc.sampler.d.res_y = 4;
        { kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1; t.color[0] = bg.x; t.color[1] = bg.y; t.color[2] = bg.z; sc.textures.push_back(t); }
        sc.background = scn % 3 == 2 ? -1 : 0;
        /* reference scene */
        std::vector<kazen::Mesh> kmeshes(specs.size()); std::vector<kazen::Mesh *> kptr;
        std::vector<kazen::KissBodies> kkiss(specs.size()); std::vector<kazen::DiffuseBodies> kdiff(specs.size()); std::vector<kazen::AreaLightBodies> klights(specs.size());
        std::vector<std::vector<kazen::Texture<kazen::Color3f>>> ktex(specs.size(), std::vector<kazen::Texture<kazen::Color3f>>(4));
        std::vector<kazen::NormalMapBodies> knmap(specs.size());
        std::vector<kazen::DiscretePDF> kdpdf(specs.size());
        kazen::Scene kscene; kscene.m_background = KC(bg);
        for (size_t g = 0; g < specs.size(); ++g) {
            const Spec &m = specs[g]; kazen::Mesh &km = kmeshes[g]; kzo::MeshData om;
            const int nV = (int)m.P.size(), nF = (int)m.F.size() / 3;
            km.m_V.resize(3, nV); if (!m.N.empty()) km.m_N.resize(3, nV); if (!m.UV.empty()) km.m_UV.resize(2, nV); km.m_F.resize(3, nF);
            om.nV = (uint32_t)nV; om.nF = (uint32_t)nF;
            for (int i = 0; i < nV; ++i) {
                km.m_V(0, i) = m.P[i].x; km.m_V(1, i) = m.P[i].y; km.m_V(2, i) = m.P[i].z; om.P.insert(om.P.end(), {m.P[i].x, m.P[i].y, m.P[i].z});
                if (!m.N.empty()) { km.m_N(0, i) = m.N[i].x; km.m_N(1, i) = m.N[i].y; km.m_N(2, i) = m.N[i].z; om.N.insert(om.N.end(), {m.N[i].x, m.N[i].y, m.N[i].z}); }
                if (!m.UV.empty()) { km.m_UV(0, i) = m.UV[i].x; km.m_UV(1, i) = m.UV[i].y; om.UV.insert(om.UV.end(), {m.UV[i].x, m.UV[i].y}); }
            }
            for (int f = 0; f < nF; ++f) for (int k = 0; k < 3; ++k) { km.m_F(k, f) = m.F[3 * f + k]; om.F.push_back(m.F[3 * f + k]); }
            /* material */
            kz_bsdf_desc bd; memset(&bd, 0, sizeof(bd));
            if (m.kind >= 1) {
                const kzo::V3 base(rnd(0.1f, 0.9f), rnd(0.1f, 0.9f), rnd(0.1f, 0.9f)); const float rough = rnd(0.05f, 0.9f), metal = g % 2 ? rnd() : 0.f;
                ktex[g][0].value = KC(base); ktex[g][1].value = kazen::Color3f(rough); ktex[g][2].value = kazen::Color3f(metal);
                kazen::KissBodies &kb = kkiss[g]; kb.m_baseColor = &ktex[g][0]; kb.m_roughness = &ktex[g][1]; kb.m_metallic = &ktex[g][2];
                kb.m_anisotropy = 0.f; kb.m_specular = rnd(); kb.m_specularTint = rnd(); kb.m_sheen = g % 3 ? 0.f : rnd(); kb.m_sheenTint = rnd(); kb.m_clearcoat = g % 2 ? 0.f : rnd(); kb.m_clearcoatRoughness = rnd();
                km.m_bsdf = &kb;
                kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
                bd.type = KZ_BSDF_KISS; bd.base_color = (int)sc.textures.size(); t.color[0] = base.x; t.color[1] = base.y; t.color[2] = base.z; sc.textures.push_back(t);
                bd.roughness = (int)sc.textures.size(); t.color[0] = t.color[1] = t.color[2] = rough; sc.textures.push_back(t);
                bd.metallic = (int)sc.textures.size(); t.color[0] = t.color[1] = t.color[2] = metal; sc.textures.push_back(t);
                bd.anisotropy = 0.f; bd.specular = kb.m_specular; bd.specular_tint = kb.m_specularTint; bd.sheen = kb.m_sheen; bd.sheen_tint = kb.m_sheenTint; bd.clearcoat = kb.m_clearcoat; bd.clearcoat_roughness = kb.m_clearcoatRoughness;
                if (m.kind == 2) {            /* the kiss material becomes the nested BSDF of a normal map */
                    const kzo::V3 nrgb(rnd(0.3f, 0.7f), rnd(0.3f, 0.7f), rnd(0.7f, 1.f));
                    ktex[g][3].value = KC(nrgb); knmap[g].m_normalMap = &ktex[g][3]; knmap[g].m_nested = &kb; km.m_bsdf = &knmap[g];
                    sc.bsdfs.push_back(bd);
                    t.color[0] = nrgb.x; t.color[1] = nrgb.y; t.color[2] = nrgb.z;
                    kz_bsdf_desc nd; memset(&nd, 0, sizeof(nd)); nd.type = KZ_BSDF_NORMALMAP; nd.nested = (int)sc.bsdfs.size() - 1; nd.normal_map = (int)sc.textures.size(); sc.textures.push_back(t);
                    bd = nd;
                }
            } else {
                const kzo::V3 alb(rnd(0.2f, 0.8f), rnd(0.2f, 0.8f), rnd(0.2f, 0.8f));
                kdiff[g].m_albedo = KC(alb); km.m_bsdf = &kdiff[g];
                bd.type = KZ_BSDF_DIFFUSE; bd.albedo[0] = alb.x; bd.albedo[1] = alb.y; bd.albedo[2] = alb.z;
            }
            om.bsdf = (int)sc.bsdfs.size(); sc.bsdfs.push_back(bd);
            /* emitter: Mesh::activate (mesh.cpp:31-43) */
            om.light = -1;
            if (m.light >= 0) {
                const kzo::V3 rad(rnd(4, 12), rnd(4, 12), rnd(4, 12));
                klights[g].m_radiance = KC(rad); klights[g].m_lightPrimaryVisibility = m.visible; km.m_light = &klights[g];
                kdpdf[g] = kazen::DiscretePDF((size_t)nF); kdpdf[g].reserve((size_t)nF);
                for (int i = 0; i < nF; ++i) kdpdf[g].append(km.surfaceArea((uint32_t)i));
                kdpdf[g].normalize(); km.m_dpdf = &kdpdf[g];
                kz_light_desc ld; ld.radiance[0] = rad.x; ld.radiance[1] = rad.y; ld.radiance[2] = rad.z; ld.primary_visibility = m.visible ? 1 : 0;
                om.light = (int)sc.lights.size(); sc.lights.push_back(ld);
                kzo::buildLightCdf(om); sc.lightMeshes.push_back((int)g);
            }
            for (uint32_t f = 0; f < om.nF; ++f) { kzo::Tri t; t.p0 = om.pos(om.F[3 * f]); t.p1 = om.pos(om.F[3 * f + 1]); t.p2 = om.pos(om.F[3 * f + 2]); t.geom = (uint32_t)g; t.prim = f; sc.accel.tris.push_back(t); }
            sc.meshes.push_back(std::move(om));
        }
        sc.accel.build();
        for (kazen::Mesh &km : kmeshes) { kptr.push_back(&km); if (km.isLight()) kscene.m_lights.push_back(&km); }
        kscene.m_meshes = kptr;
        RefAccel ra{&sc.accel, &kscene.m_meshes}; kscene.m_accel = &ra;
        kazen::PathMisBodies ki; ki.m_maxDepth = maxDepth; ki.m_rayEpsilon = eps; ki.m_regularization = regularize; ki.m_accumulatedRoughness = 0.5f;
        for (int pth = 0; pth < 1500; ++pth) {
            const int px = (int)(rnd() * 640), py = (int)(rnd() * 480), sidx = (int)(rnd() * 16) % 16;
            const kzo::V3 o(rnd(-0.5f, 0.5f), rnd(0.8f, 1.6f), 4.f), target(rnd(-2.5f, 2.5f), rnd(0.f, 3.8f), rnd(-2.5f, 1.f)), d = kzo::normalized(target - o);
            kazen::StratifiedBodies ks; ks.m_seed = sc.sampler.d.seed; ks.m_sampleCount = 16; ks.m_resolution = 4;
            ks.generateSample(kazen::Point2i(px, py), sidx); ks.nextPixel2D(); ks.next2D();                                             /* renderer.cpp:25-28 */
            const kazen::Color3f kL = ki.Li(&kscene, &ks, kazen::Ray3f(kazen::Point3f(o.x, o.y, o.z), K(d), 1e-4f, 1e4f));
            kzo::Sampler osm; osm.cfg = &sc.sampler; osm.generateSample(px, py, sidx); osm.nextPixel2D(); osm.next2D();
            PathCounters pc; const kz_ray kr{{o.x, o.y, o.z}, 1e-4f, {d.x, d.y, d.z}, 1e4f};
            const kzo::V3 oL = Li(&os, osm, kr, pc);
            rec("pathMisLi", {(float)scn, (float)px, (float)py, (float)sidx, o.x, o.y, d.x, d.y, d.z}, f3(kL), f3(oL), false);
            ++liPaths; if (oL.x > 0.f || oL.y > 0.f || oL.z > 0.f) ++liLit;
        }
        /* renderer::renderSample on the same scene: a camera (both models), the Stratified sampler, the hosted Li and a whole-frame ImageBlock
         * against the oracle's render loop body (kzo_render: generateSample, nextPixel2D, next2D, cameraRay, Li, filmPut) */
        {
            const int W = 64, H = 48;
            kz_camera_desc &c = sc.camera; memset(&c, 0, sizeof(c));
            c.type = scn % 2 ? KZ_CAM_THINLENS : KZ_CAM_PERSPECTIVE; c.width = W; c.height = H; c.near_clip = 1e-4f; c.far_clip = 1e4f; c.aperture_radius = 0.05f; c.focus_distance = 4.5f;
            Eigen::Matrix4f s2c, c2w;
            const float m1[16] = {1.3f, 0, 0, -0.65f, 0, -0.98f, 0, 0.49f, 0, 0, 0, 1, 0, 0, 0, 1};               /* (u, v, 0) -> ((2u-1) tx, (1-2v) ty, 1) */
            const float m2[16] = {-1, 0, 0, rnd(-0.3f, 0.3f), 0, 1, 0, rnd(1.0f, 1.5f), 0, 0, -1, 4.f, 0, 0, 0, 1};  /* camera at (x, y, 4) looking down -z */
            for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { s2c(i, j) = m1[4 * i + j]; c.sample_to_camera[4 * i + j] = m1[4 * i + j]; c2w(i, j) = m2[4 * i + j]; c.camera_to_world[4 * i + j] = m2[4 * i + j]; }
            kazen::HostedPerspective kcp; kazen::HostedThinlens kct;
            kcp.b.m_invOutputSize = kct.b.m_invOutputSize = kazen::Vector2f(1.0f / W, 1.0f / H);
            kcp.b.m_sampleToCamera = kct.b.m_sampleToCamera = kazen::Transform(s2c, s2c); kcp.b.m_cameraToWorld = kct.b.m_cameraToWorld = kazen::Transform(c2w, c2w);
            kcp.b.m_nearClip = kct.b.m_nearClip = c.near_clip; kcp.b.m_farClip = kct.b.m_farClip = c.far_clip; kct.b.m_apertureRadius = c.aperture_radius; kct.b.m_focusDistance = c.focus_distance;
            kazen::HostedPathMis kint; kint.b = ki;
            kscene.m_integrator = &kint; kscene.m_camera = scn % 2 ? (const kazen::Camera *)&kct : (const kazen::Camera *)&kcp;
            kazen::GaussianBodies fg; fg.m_radius = 2.f; fg.m_stddev = 0.5f;
            kazen::ImageBlock blk(kazen::Vector2i(W, H), &fg);
            sc.filter.radius = fg.getRadius(); for (int i = 0; i <= 32; ++i) sc.filter.table[i] = blk.m_filter[i];
            const int border = (int)std::ceil(sc.filter.radius - 0.5f);
            std::vector<float> frame((size_t)(W + 2 * border) * (H + 2 * border) * 4, 0.f);
            for (int k = 0; k < 400; ++k) {
                const int px = (int)(rnd() * W) % W, py = (int)(rnd() * H) % H, sidx = (int)(rnd() * 16) % 16;
                kazen::StratifiedBodies ks; ks.m_seed = sc.sampler.d.seed; ks.m_sampleCount = 16; ks.m_resolution = 4;
                ks.generateSample(kazen::Point2i(px, py), sidx);
                kazen::renderer::renderSample(&kscene, &ks, blk, kazen::Point2i(px, py));
                kzo::Sampler osm; osm.cfg = &sc.sampler; osm.generateSample(px, py, sidx);
                const kzo::V2 pix = osm.nextPixel2D(); const kzo::V2 pixelSample{float(px) + pix.x, float(py) + pix.y}; const kzo::V2 apertureSample = osm.next2D();
                const kz_ray ray = cameraRay(sc.camera, pixelSample, apertureSample);
                PathCounters pc; const kzo::V3 value = kzo::V3(1.0f) * Li(&os, osm, ray, pc);
                filmPut(sc, border, frame.data(), pixelSample, value);
            }
            std::vector<float> ref; for (const kazen::Color4f &cc : blk.m_px) { ref.push_back(cc.x()); ref.push_back(cc.y()); ref.push_back(cc.z()); ref.push_back(cc.w()); }
            double lit = 0; for (size_t i = 0; i < ref.size(); i += 4) lit += ref[i] + ref[i + 1] + ref[i + 2];
            rec("renderSample", {(float)scn, (float)(scn % 2)}, ref, frame, false);
            if (!(lit > 0)) fprintf(stderr, "renderSample: scene %d rendered black\n", scn);
        }
        /* the other four integrators on the same scene (not on the normal-mapped ones: they build the BSDF record without `its`) */
        if (scn % 3 != 1) {
            kazen::NormalIntegratorBodies kin; kazen::AoIntegratorBodies kia; kazen::WhittedIntegratorBodies kiw; kazen::PathMatsIntegratorBodies kip;
            for (int pth = 0; pth < 400; ++pth) {
                const int type = 1 + pth % 4;            /* kz_integrator_type: normals, ao, whitted, path_mats */
                const int px = (int)(rnd() * 640), py = (int)(rnd() * 480), sidx = (int)(rnd() * 16) % 16;
                const kzo::V3 o(rnd(-0.5f, 0.5f), rnd(0.8f, 1.6f), 4.f), target(rnd(-2.5f, 2.5f), rnd(0.f, 3.8f), rnd(-2.5f, 1.f)), d = kzo::normalized(target - o);
                kazen::StratifiedBodies ks; ks.m_seed = sc.sampler.d.seed; ks.m_sampleCount = 16; ks.m_resolution = 4;
                ks.generateSample(kazen::Point2i(px, py), sidx); ks.nextPixel2D(); ks.next2D();
                const kazen::Ray3f kray(kazen::Point3f(o.x, o.y, o.z), K(d), 1e-4f, 1e4f);
                const kazen::Color3f kL = type == 1 ? kin.Li(&kscene, &ks, kray) : (type == 2 ? kia.Li(&kscene, &ks, kray) : (type == 3 ? kiw.Li(&kscene, &ks, kray) : kip.Li(&kscene, &ks, kray)));
                kzo::Sampler osm; osm.cfg = &sc.sampler; osm.generateSample(px, py, sidx); osm.nextPixel2D(); osm.next2D();
                PathCounters pc; const kz_ray kr{{o.x, o.y, o.z}, 1e-4f, {d.x, d.y, d.z}, 1e4f};
                sc.integrator.type = type;
                const kzo::V3 oL = LiAlt(&os, osm, kr, pc);
                sc.integrator.type = KZ_INTEGRATOR_PATH_MIS;
                const char *names[5] = {"", "normalsLi", "aoLi", "whittedLi", "pathMatsLi"};
                rec(names[type], {(float)scn, (float)px, (float)py, (float)sidx, o.x, o.y, d.x, d.y, d.z}, f3(kL), f3(oL), false);
            }
        }
    }
    fprintf(stderr, "pathMisLi: %ld paths, %ld with non-zero radiance\n", liPaths, liLit);
    /* film: filter evaluation + ImageBlock's tabulation (block.cpp:9-31) + put (block.cpp:56-85) against the oracle's filmPut fed with
     * the reference's own table; the tables go to the golden file so that pykazen.filter_table and the C++ host's rfilter plugins are
     * checked against them too */
    for (int fk = 0; fk < 6; ++fk) {
        kazen::GaussianBodies fg; kazen::MitchellBodies fm; kazen::TentBodies ft; kazen::BoxBodies fb; const kazen::ReconstructionFilter *rf = nullptr;
        float p1 = 0, p2 = 0;
        if (fk == 0) { fg.m_radius = 2.f; fg.m_stddev = 0.5f; rf = &fg; p1 = 0.5f; }
        else if (fk == 1) { fg.m_radius = 3.f; fg.m_stddev = 1.0f; rf = &fg; p1 = 1.0f; }
        else if (fk == 2) { fm.m_radius = 2.f; fm.m_B = 1.0f / 3.0f; fm.m_C = 1.0f / 3.0f; rf = &fm; p1 = fm.m_B; p2 = fm.m_C; }
        else if (fk == 3) { fm.m_radius = 3.f; fm.m_B = 0.2f; fm.m_C = 0.4f; rf = &fm; p1 = fm.m_B; p2 = fm.m_C; }
        else if (fk == 4) { ft.m_radius = 1.f; rf = &ft; }
        else { fb.m_radius = 0.5f; rf = &fb; }
        const int W = 37, H = 29;
        kazen::ImageBlock blk(kazen::Vector2i(W, H), rf);
        kzo::SceneData sc; memset(&sc.camera, 0, sizeof(sc.camera)); sc.camera.width = W; sc.camera.height = H;
        sc.filter.radius = rf->getRadius(); for (int i = 0; i <= 32; ++i) sc.filter.table[i] = blk.m_filter[i];
        rec(fk < 2 ? "filterTableGaussian" : (fk < 4 ? "filterTableMitchell" : (fk == 4 ? "filterTableTent" : "filterTableBox")), {rf->getRadius(), p1, p2},
            std::vector<float>(blk.m_filter, blk.m_filter + 33), std::vector<float>(sc.filter.table, sc.filter.table + 33), true);
        const int border = (int)std::ceil(sc.filter.radius - 0.5f);
        std::vector<float> frame((size_t)(W + 2 * border) * (H + 2 * border) * 4, 0.f);
        for (int k = 0; k < 600; ++k) {
            const kzo::V2 pos{rnd(-0.2f, W + 0.2f), rnd(-0.2f, H + 0.2f)};
            kzo::V3 val(rnd(0, 5), rnd(0, 5), rnd(0, 5));
            if (k % 37 == 5) val.x = -0.1f; if (k % 41 == 7) val.y = NAN; if (k % 43 == 9) val.z = INFINITY;
            blk.put(kazen::Point2f(pos.x, pos.y), KC(val));
            filmPut(sc, border, frame.data(), pos, val);
        }
        std::vector<float> ref; for (const kazen::Color4f &c : blk.m_px) { ref.push_back(c.x()); ref.push_back(c.y()); ref.push_back(c.z()); ref.push_back(c.w()); }
        rec("imageBlockPut", {(float)fk}, ref, frame, false);
    }
    /* resolve: ImageBlock::toBitmap's divideByFilterWeight (block.cpp:39-45, color.h:93-98) + savePNG's 8-bit conversion (bitmap.cpp:46-54) against kzo_resolve */
    {
        const int N = 4000;
        kzo_scene os; memset(&os.sc.camera, 0, sizeof(os.sc.camera)); os.sc.camera.width = N; os.sc.camera.height = 1; os.border = 0;
        std::vector<float> frame((size_t)N * 4), lin((size_t)N * 3), refLin, refQ, ourQ; std::vector<uint8_t> q8((size_t)N * 3);
        for (int k = 0; k < N; ++k) {
            const float w = k % 11 == 0 ? 0.f : rnd(0.05f, 40.f), scale = k % 7 == 0 ? 30.f : 1.2f;
            float *p = &frame[(size_t)4 * k];
            p[0] = rnd(-0.05f, scale) * w; p[1] = rnd(0.f, scale) * w; p[2] = (k % 13 == 0 ? 1e-4f : rnd(0.f, scale)) * w; p[3] = w;
            if (k % 97 == 3) p[1] = INFINITY;
            kazen::Color4f c4(p[0], p[1], p[2], p[3]);
            kazen::Color3f c = c4.divideByFilterWeight();
            refLin.push_back(c.x()); refLin.push_back(c.y()); refLin.push_back(c.z());
            uint8_t d[3]; kazen::refQuantise(c, d);
            refQ.push_back(d[0]); refQ.push_back(d[1]); refQ.push_back(d[2]);
        }
        kzo_resolve(&os, frame.data(), lin.data(), q8.data());
        for (uint8_t v : q8) ourQ.push_back(v);
        rec("resolveLinear", {(float)N}, refLin, lin, false);
        rec("resolveSRGB8", {(float)N}, refQ, ourQ, false);
    }
    /* cameras: sampleRay of both camera models with random (invertible-looking) matrices; the matrices themselves come from
     * Camera::activate, which uses Eigen's 4x4 inverse and is not restated here */
    for (int t = 0; t < 4000; ++t) {
        kz_camera_desc c; memset(&c, 0, sizeof(c));
        c.type = t % 2 ? KZ_CAM_THINLENS : KZ_CAM_PERSPECTIVE; c.width = 64 + (int)(rnd() * 4000); c.height = 64 + (int)(rnd() * 2200);
        c.near_clip = rnd(1e-4f, 0.5f); c.far_clip = rnd(50.f, 1e4f); c.aperture_radius = rnd(0.f, 0.3f); c.focus_distance = rnd(0.5f, 20.f);
        Eigen::Matrix4f s2c, c2w;
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
            const float a = (i == j ? rnd(0.5f, 2.f) : rnd(-0.4f, 0.4f)) * (i == 3 && j < 3 ? 0.1f : 1.f), b = i == 3 ? (j == 3 ? 1.f : 0.f) : rnd(-1.f, 1.f) * (j == 3 ? 5.f : 1.f);
            s2c(i, j) = a; c.sample_to_camera[4 * i + j] = a; c2w(i, j) = b; c.camera_to_world[4 * i + j] = b;
        }
        const kzo::V2 sp{rnd() * c.width, rnd() * c.height}, ap{rnd(), rnd()};
        kazen::Ray3f kr;
        if (t % 2) { kazen::ThinlensBodies cam; cam.m_invOutputSize = kazen::Vector2f(1.0f / c.width, 1.0f / c.height); cam.m_sampleToCamera = kazen::Transform(s2c, s2c); cam.m_cameraToWorld = kazen::Transform(c2w, c2w);
                     cam.m_nearClip = c.near_clip; cam.m_farClip = c.far_clip; cam.m_apertureRadius = c.aperture_radius; cam.m_focusDistance = c.focus_distance;
                     cam.sampleRay(kr, kazen::Point2f(sp.x, sp.y), kazen::Point2f(ap.x, ap.y)); }
        else { kazen::PerspectiveBodies cam; cam.m_invOutputSize = kazen::Vector2f(1.0f / c.width, 1.0f / c.height); cam.m_sampleToCamera = kazen::Transform(s2c, s2c); cam.m_cameraToWorld = kazen::Transform(c2w, c2w);
               cam.m_nearClip = c.near_clip; cam.m_farClip = c.far_clip; cam.sampleRay(kr, kazen::Point2f(sp.x, sp.y), kazen::Point2f(ap.x, ap.y)); }
        const kz_ray orr = cameraRay(c, sp, ap);
        rec(t % 2 ? "thinlensSampleRay" : "perspectiveSampleRay", {sp.x, sp.y, ap.x, ap.y, (float)c.width, (float)c.height},
            {kr.o.x(), kr.o.y(), kr.o.z(), kr.d.x(), kr.d.y(), kr.d.z(), kr.mint, kr.maxt}, {orr.o[0], orr.o[1], orr.o[2], orr.d[0], orr.d[1], orr.d[2], orr.tmin, orr.tmax}, false);
    }
    /* samplers: the draw pattern of one path (pixel 2D, aperture 2D, then per vertex 1D x5 + 2D) for random pixels / sample indices */
    for (int t = 0; t < 3000; ++t) {
        const int type = t % 3;                          /* 0 independent, 1 stratified, 2 correlated */
        const uint32_t requested = (uint32_t[]){1, 4, 9, 16, 24, 64, 100, 256, 1024}[(t / 3) % 9];
        const uint64_t seed = t % 5 == 0 ? 1ull : (uint64_t)(rnd() * 4e9f) + 1ull;
        const int px = (int)(rnd() * 4096), py = (int)(rnd() * 2160);
        kzo::SamplerCfg cfg; memset(&cfg.d, 0, sizeof(cfg.d));
        cfg.d.seed = seed;
        kazen::IndependentBodies ki; kazen::StratifiedBodies ks; kazen::CorrelatedBodies kc;
        uint32_t count = requested;
        if (type == 1) {                                 /* Stratified ctor, sampler.cpp:83-93 */
            int res = 4; while ((uint32_t)(res * res) < requested) res++;
            count = (uint32_t)(res * res); ks.m_resolution = res; ks.m_seed = seed; ks.m_sampleCount = count;
            cfg.d.type = KZ_SAMPLER_STRATIFIED; cfg.d.res_x = cfg.d.res_y = res;
        } else if (type == 2) {                          /* Correlated ctor, sampler.cpp:178-189 */
            int ry = (int)sqrt((double)requested), rx = (int)((requested + ry - 1) / ry);
            count = (uint32_t)(rx * ry); kc.m_resolution = kazen::Point2i(rx, ry); kc.m_seed = seed; kc.m_sampleCount = count;
            cfg.d.type = KZ_SAMPLER_CORRELATED; cfg.d.res_x = rx; cfg.d.res_y = ry;
        } else { ki.m_seed = seed; ki.m_sampleCount = count; cfg.d.type = KZ_SAMPLER_INDEPENDENT; }
        cfg.d.sample_count = count;
        const int sidx = (int)(rnd() * count) % (int)count;
        kzo::Sampler os; os.cfg = &cfg; os.generateSample(px, py, sidx);
        std::vector<float> ref, ours;
        auto draw = [&](int kind) {                      /* 0: next1D, 1: next2D, 2: nextPixel2D */
            if (kind == 0) { ref.push_back(type == 0 ? ki.next1D() : (type == 1 ? ks.next1D() : kc.next1D())); ours.push_back(os.next1D()); return; }
            kazen::Point2f r = type == 0 ? (kind == 2 ? ki.nextPixel2D() : ki.next2D()) : (type == 1 ? (kind == 2 ? ks.nextPixel2D() : ks.next2D()) : (kind == 2 ? kc.nextPixel2D() : kc.next2D()));
            const kzo::V2 o = kind == 2 ? os.nextPixel2D() : os.next2D();
            ref.push_back(r.x()); ref.push_back(r.y()); ours.push_back(o.x); ours.push_back(o.y);
        };
        if (type == 0) ki.generateSample(kazen::Point2i(px, py), sidx); else if (type == 1) ks.generateSample(kazen::Point2i(px, py), sidx); else kc.generateSample(kazen::Point2i(px, py), sidx);
        draw(2); draw(1);
        for (int v = 0; v < 4; ++v) { for (int k = 0; k < 5; ++k) draw(0); draw(1); }
        const char *names[3] = {"samplerIndependent", "samplerStratified", "samplerCorrelated"};
        rec(names[type], {(float)px, (float)py, (float)sidx, (float)requested, (float)(seed & 0xFFFFFF), (float)(seed >> 24)}, ref, ours, t < 60);
    }
    /* ---- dielectric / mirror / lambertian / ggx / roughconductor / roughplastic / roughdielectric, eval / pdf / sample ---- */
    for (int i = 0; i < 6000; ++i) {
        const bool keep = i < 42;
        const int type = KZ_BSDF_DIELECTRIC + i % 7;
        const kzo::V3 base(rnd(), rnd(), rnd());
        const float rough = i % 11 == 0 ? 0.02f : rnd(0.03f, 0.95f), intIOR = i % 5 == 0 ? 1.5046f : rnd(1.05f, 2.4f), extIOR = i % 5 == 0 ? 1.000277f : rnd(1.f, 1.3f);
        const float aniso = i % 3 == 0 ? 0.f : rnd(-0.8f, 0.8f);
        const bool twoSided = type == KZ_BSDF_DIELECTRIC || type == KZ_BSDF_ROUGHDIELECTRIC;
        const kzo::V3 wi = rdir(twoSided ? (i % 2 == 0) && false : (i % 13 != 0)), wo = rdir(twoSided ? false : (i % 17 != 0));
        const float s1 = rnd(); const kzo::V2 s2{rnd(), rnd()};
        kzo::SceneData sc;
        kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
        t.color[0] = base.x; t.color[1] = base.y; t.color[2] = base.z; sc.textures.push_back(t);
        kz_bsdf_desc m; memset(&m, 0, sizeof(m)); m.type = type; m.base_color = 0; m.roughness = m.metallic = m.normal_map = m.nested = -1;
        m.int_ior = intIOR; m.ext_ior = extIOR; m.anisotropy = aniso;
        kazen::Texture<kazen::Color3f> tb; tb.value = KC(base);
        kazen::DielectricBodies kd; kazen::MirrorBodies km; kazen::LambertianBodies kl; kazen::GGXBodies kg; kazen::RoughConductorBodies kc; kazen::RoughPlasticBodies kp; kazen::RoughDielectricBodies kr;
        kazen::BSDF *ref = nullptr;
        const float alpha = std::max(0.001f, rough * rough);                      /* the constructors' max(MIN_ALPHA, sqr(roughness)), bsdf.cpp:698-700,820-822,958-959 */
        switch (type) {
            case KZ_BSDF_DIELECTRIC: kd.m_intIOR = intIOR; kd.m_extIOR = extIOR; ref = &kd; break;
            case KZ_BSDF_MIRROR: ref = &km; break;
            case KZ_BSDF_LAMBERTIAN: kl.m_albedo = &tb; ref = &kl; break;
            case KZ_BSDF_GGX: kg.m_albedo = &tb; kg.m_roughness = rough; kg.m_anisotropy = aniso; m.alpha = rough; ref = &kg; break;
            case KZ_BSDF_ROUGHCONDUCTOR: {
                const float e[3][3] = {{0.1431189557f, 0.3749570432f, 1.4424785571f}, {0.2004376970f, 0.9240334304f, 1.1022119527f}, {4.3696828663f, 2.9167024892f, 1.6547005413f}};
                const float k[3][3] = {{3.9831604247f, 2.3857207478f, 1.6032152899f}, {3.9129485033f, 2.4528477015f, 2.1421879552f}, {5.2064337956f, 4.2313645277f, 3.7549467933f}};
                const int mat = (i / 7) % 3;                                      /* Au, Cu, Cr: bsdf.cpp:703-714 */
                kc.m_alpha = alpha; kc.m_eta = kazen::Color3f(e[mat][0], e[mat][1], e[mat][2]); kc.m_k = kazen::Color3f(k[mat][0], k[mat][1], k[mat][2]);
                m.alpha = alpha; for (int c = 0; c < 3; ++c) { m.eta[c] = e[mat][c]; m.k[c] = k[mat][c]; }
                ref = &kc; break; }
            case KZ_BSDF_ROUGHPLASTIC:
                kp.m_alpha = alpha; kp.m_intIOR = intIOR; kp.m_extIOR = extIOR; kp.m_kd = KC(base); kp.m_ks = 1 - kp.m_kd.maxCoeff();      /* bsdf.cpp:840 */
                m.alpha = alpha; m.albedo[0] = base.x; m.albedo[1] = base.y; m.albedo[2] = base.z; ref = &kp; break;
            default:
                kr.m_intIOR = intIOR; kr.m_extIOR = extIOR; kr.m_alpha = alpha; kr.m_eta = intIOR / extIOR; kr.m_invEta = extIOR / intIOR;    /* bsdf.cpp:961-962 */
                m.alpha = alpha; ref = &kr; break;
        }
        sc.bsdfs.push_back(m);
        std::vector<float> in = {(float)type, base.x, base.y, base.z, rough, aniso, intIOR, extIOR, (float)((i / 7) % 3), wi.x, wi.y, wi.z, wo.x, wo.y, wo.z, s1, s2.x, s2.y};
        kazen::BSDFQueryRecord ke(K(wi), K(wo), kazen::ESolidAngle); ke.uv = kazen::Point2f(0.5f, 0.5f);
        kzo::BSDFQueryRecord oe(wi, wo, kzo::ESolidAngle); oe.uv = kzo::V2{0.5f, 0.5f};
        rec("extraEval", in, f3(ref->eval(ke)), f3(kzo::bsdfEval(sc, 0, oe)), keep);
        rec("extraPdf", in, {ref->pdf(ke)}, {kzo::bsdfPdf(sc, 0, oe)}, keep);
        kazen::BSDFQueryRecord ks(K(wi)); ks.uv = kazen::Point2f(0.5f, 0.5f);
        kzo::BSDFQueryRecord os(wi); os.uv = kzo::V2{0.5f, 0.5f};
        const kazen::Color3f kw = ref->sample(ks, s1, kazen::Point2f(s2.x, s2.y)); const kzo::V3 ow = kzo::bsdfSample(sc, 0, os, s1, s2);
        const bool z0 = kw.x() == 0.f && kw.y() == 0.f && kw.z() == 0.f;          /* wo / eta / measure only matter for a non-zero weight */
        rec("extraSample", in, {kw.x(), kw.y(), kw.z(), z0 ? 0.f : ks.wo.x(), z0 ? 0.f : ks.wo.y(), z0 ? 0.f : ks.wo.z(), z0 ? 0.f : ks.eta, z0 ? 0.f : (float)(ks.measure == kazen::EDiscrete)},
            {ow.x, ow.y, ow.z, z0 ? 0.f : os.wo.x, z0 ? 0.f : os.wo.y, z0 ? 0.f : os.wo.z, z0 ? 0.f : os.eta, z0 ? 0.f : (float)(os.measure == kzo::EDiscrete)}, keep);
    }
    /* ---- texture expression nodes (texture.cpp:104-270) over constant leaves: background (uv and direction), colorramp, blend mix / multiply / other,
     *      missing children, nesting two deep ---- */
    for (int i = 0; i < 3000; ++i) {
        const bool keep = i < 30;
        kzo::SceneData sc;
        auto constant = [&](kzo::V3 c) { kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
                                         t.color[0] = c.x; t.color[1] = c.y; t.color[2] = c.z; sc.textures.push_back(t); return (int)sc.textures.size() - 1; };
        const kzo::V3 c0(rnd(-0.3f, 1.4f), rnd(-0.3f, 1.4f), rnd(-0.3f, 1.4f)), c1(rnd(), rnd(), rnd()), c2(rnd(0.f, 2.f), rnd(0.f, 2.f), rnd(0.f, 2.f));
        kazen::Texture<kazen::Color3f> k0, k1, k2; k0.value = KC(c0); k1.value = KC(c1); k2.value = KC(c2);
        const int n0 = constant(c0), n1 = constant(c1), n2 = constant(c2);
        const float lo = rnd(-0.5f, 0.5f), hi = rnd(0.2f, 1.5f), intensity = rnd(0.f, 5.f);
        /* colorramp over c0 */
        kazen::ColorRampTexBodies kr; kr.m_min = lo; kr.m_max = hi; kr.m_nested = i % 9 == 0 ? nullptr : &k0;
        kz_texture_desc tr; memset(&tr, 0, sizeof(tr)); tr.type = KZ_TEX_COLORRAMP; tr.a = lo; tr.b = hi; tr.child[0] = i % 9 == 0 ? -1 : n0; tr.child[1] = tr.child[2] = -1;
        sc.textures.push_back(tr); const int nr = (int)sc.textures.size() - 1;
        /* blend(mask = ramp, input1 = c1, input2 = c2) */
        const int mode = i % 4 == 3 ? KZ_BLEND_OTHER : (i % 2 ? KZ_BLEND_MULTIPLY : KZ_BLEND_MIX);
        kazen::BlendTexBodies kb; kb.m_blendmode = mode == KZ_BLEND_MIX ? "mix" : (mode == KZ_BLEND_MULTIPLY ? "multiply" : "screen");
        kb.m_mask = i % 5 == 0 ? nullptr : &kr; kb.m_input1 = i % 7 == 0 ? nullptr : &k1; kb.m_input2 = i % 11 == 0 ? nullptr : &k2;
        kz_texture_desc tb; memset(&tb, 0, sizeof(tb)); tb.type = KZ_TEX_BLEND; tb.mode = mode;
        tb.child[0] = i % 5 == 0 ? -1 : nr; tb.child[1] = i % 7 == 0 ? -1 : n1; tb.child[2] = i % 11 == 0 ? -1 : n2;
        sc.textures.push_back(tb); const int nb = (int)sc.textures.size() - 1;
        /* background(intensity) over the blend */
        kazen::BackgroundTexBodies kg; kg.m_intensity = intensity; kg.m_nested = i % 13 == 0 ? nullptr : &kb;
        kz_texture_desc tg; memset(&tg, 0, sizeof(tg)); tg.type = KZ_TEX_BACKGROUND; tg.a = intensity; tg.child[0] = i % 13 == 0 ? -1 : nb; tg.child[1] = tg.child[2] = -1;
        sc.textures.push_back(tg); const int ng = (int)sc.textures.size() - 1;
        const kzo::V2 uv{rnd(), rnd()}; const kzo::V3 dir = rdir(false);
        std::vector<float> in = {c0.x, c0.y, c0.z, c1.x, c1.y, c1.z, c2.x, c2.y, c2.z, lo, hi, intensity, (float)mode, (float)(i % 5 == 0), (float)(i % 7 == 0), (float)(i % 9 == 0), (float)(i % 11 == 0), (float)(i % 13 == 0)};
        rec("texColorRamp", in, f3(kr.eval(kazen::Point2f(uv.x, uv.y))), f3(kzo::evalTextureUV(sc, nr, uv)), keep);
        rec("texBlend", in, f3(kb.eval(kazen::Point2f(uv.x, uv.y))), f3(kzo::evalTextureUV(sc, nb, uv)), keep);
        rec("texBackgroundUV", in, f3(kg.eval(kazen::Point2f(uv.x, uv.y))), f3(kzo::evalTextureUV(sc, ng, uv)), keep);
        /* eval(dir): only the background node and constants forward a direction (texture.cpp:121-126, texture.h:12); a constant nested directly */
        kazen::BackgroundTexBodies kd; kd.m_intensity = intensity; kd.m_nested = &k2;
        kz_texture_desc td = tg; td.child[0] = n2; sc.textures.push_back(td);
        rec("texBackgroundDir", in, f3(kd.eval(K(dir))), f3(kzo::evalTextureDir(sc, (int)sc.textures.size() - 1, dir)), keep);
        /* Scene::getBackgroundColor over that node: no background at all, a direction with a NaN component, an ordinary direction */
        kazen::SceneBackgroundBodies ks; ks.m_background = i % 17 == 0 ? nullptr : &kd;
        sc.background = i % 17 == 0 ? -1 : (int)sc.textures.size() - 1;
        kzo::V3 bd = dir;
        if (i % 6 == 1) bd.x = std::nanf(""); else if (i % 6 == 3) bd.y = std::nanf(""); else if (i % 6 == 5) bd.z = std::nanf("");
        in.push_back((float)(i % 17 == 0)); in.push_back((float)(i % 6));
        rec("sceneBackground", in, f3(ks.getBackgroundColor(K(bd))), f3(kzo::backgroundColor(sc, bd)), keep);
    }
    /* ---- PMJ02BN (sampler.cpp:273-390): constructor bucketing + generateSample / nextPixel2D / next1D / next2D over synthetic tables ---- */
    {
        /* set 0: the 2-D Sobol' (0,2)-sequence (van der Corput x Sobol' dimension 2); sets 1..4: digit-scrambled copies; blue noise: hashed */
        for (uint32_t i = 0; i < 65536; ++i) {
            uint32_t x = i; x = (x << 16) | (x >> 16); x = ((x & 0x00ff00ffu) << 8) | ((x & 0xff00ff00u) >> 8); x = ((x & 0x0f0f0f0fu) << 4) | ((x & 0xf0f0f0f0u) >> 4);
            x = ((x & 0x33333333u) << 2) | ((x & 0xccccccccu) >> 2); x = ((x & 0x55555555u) << 1) | ((x & 0xaaaaaaaau) >> 1);
            uint32_t y = 0; for (uint32_t v = 1u << 31, k = i; k; k >>= 1, v ^= v >> 1) if (k & 1u) y ^= v;
            for (int set = 0; set < 5; ++set) { kazen::pmj02bnSamples[set][i][0] = x ^ (set * 0x9E3779B9u); kazen::pmj02bnSamples[set][i][1] = y ^ (set * 0x85EBCA6Bu); }
        }
        uint32_t h = 12345u;
        for (int a = 0; a < 48; ++a) for (int b = 0; b < 128; ++b) for (int c = 0; c < 128; ++c) { h = h * 1664525u + 1013904223u; kazen::BlueNoiseTextures[a][b][c] = (uint16_t)(h >> 16); }
        const uint32_t counts[] = {1, 4, 16, 64, 256, 1024, 9, 24, 100};
        for (int t = 0; t < 1800; ++t) {
            const uint32_t count = counts[t % 9];
            const uint64_t seed = t % 5 == 0 ? 1ull : (uint64_t)(rnd() * 4e9f) + 1ull;
            kazen::PMJ02BNBodies kp; kp.m_seed = seed; kp.m_sampleCount = count; kp.construct();
            kzo::SamplerCfg cfg; memset(&cfg.d, 0, sizeof(cfg.d));
            cfg.d.type = KZ_SAMPLER_PMJ02BN; cfg.d.seed = seed; cfg.d.sample_count = count;
            cfg.d.blue_noise = &kazen::BlueNoiseTextures[0][0][0]; cfg.d.pmj02bn = &kazen::pmj02bnSamples[0][0][0];
            kzo::buildPmjPixelSamples(cfg);
            if (t < 9) {     /* the whole per-pixel table the constructor builds */
                std::vector<float> a, b;
                for (const kazen::Point2f &p : *kp.m_pixelSamples) { a.push_back(p.x()); a.push_back(p.y()); }
                for (const kzo::V2 &p : cfg.pmjPix.samples) { b.push_back(p.x); b.push_back(p.y); }
                rec("pmj02bnPixelTable", {(float)count, (float)kp.m_pixelTileSize}, a, b, false);
                rec("pmj02bnTileSize", {(float)count}, {(float)kp.m_pixelTileSize}, {(float)cfg.pmjPix.tileSize}, true);
            }
            for (int q = 0; q < 4; ++q) {
                const int px = (int)(rnd() * 4096), py = (int)(rnd() * 2160), sidx = (int)(rnd() * count) % (int)count;
                kp.generateSample(kazen::Point2i(px, py), sidx);
                kzo::Sampler os; os.cfg = &cfg; os.generateSample(px, py, sidx);
                std::vector<float> ref, ours;
                auto draw = [&](int kind) {
                    if (kind == 0) { ref.push_back(kp.next1D()); ours.push_back(os.next1D()); return; }
                    const kazen::Point2f r = kind == 2 ? kp.nextPixel2D() : kp.next2D(); const kzo::V2 o = kind == 2 ? os.nextPixel2D() : os.next2D();
                    ref.push_back(r.x()); ref.push_back(r.y()); ours.push_back(o.x); ours.push_back(o.y);
                };
                draw(2); draw(1);
                for (int v = 0; v < 5; ++v) { for (int k = 0; k < 5; ++k) draw(0); draw(1); }      /* 5 vertices: 2-D dimensions beyond the 5 table sets are permuted */
                rec("samplerPMJ02BN", {(float)px, (float)py, (float)sidx, (float)count, (float)(seed & 0xFFFFFF), (float)(seed >> 24)}, ref, ours, false);
            }
        }
    }
    /* DiscretePDF (dpdf.h:35-104): append / normalize / sample against the oracle's CDF sampling (kzo_shading.h cdfSample) */
    for (int t = 0; t < 200; ++t) {
        const int m = 1 + (int)(rnd() * 40);
        kazen::DiscretePDF pdf; std::vector<float> w;
        for (int i = 0; i < m; ++i) { const float a = (i % 6 == 5) ? 0.f : rnd(0.f, 3.f); w.push_back(a); pdf.append(a); }
        pdf.normalize();
        std::vector<float> cdf(1, 0.f);
        for (float a : w) cdf.push_back(cdf.back() + a);
        const float sum = cdf.back(); const float norm = 1.0f / sum;
        for (size_t i = 1; i < cdf.size(); ++i) cdf[i] *= norm;
        cdf.back() = 1.0f;
        for (int k = 0; k < 20; ++k) {
            const float v = k == 0 ? 0.f : (k == 1 ? 0.99999994f : rnd());
            std::vector<float> in = w; in.push_back(v);
            rec("dpdfSample", in, {(float)pdf.sample(v)}, {(float)kzo::cdfSample(cdf, v)}, t < 4 && k < 5);
        }
    }
    printf("{\n \"generator\": \"oracle/ref_math_kat.cpp: the reference's own function bodies (ggx_brdf.h, frame.h, dpdf.h, common.cpp:352-395,436-540, warp.cpp:41-130, bsdf.cpp:27-75 Diffuse, bsdf.cpp:1175-1371 KazenStandardSurface, sampler.cpp:43-61,111-143,207-255 Independent / Stratified / Correlated with the real hash.h + pcg32.h, mesh.cpp:47-53,108-133 Mesh::surfaceArea / sample, light.cpp:16-51 AreaLight, accel.cpp:113-236 post-intersection; mesh-based cases are checked here but not kept in the golden list) compiled against oracle/ref_shim; floats as uint32 bit patterns\",\n \"cases_checked\": %ld,\n \"mismatches\": %ld,\n \"kat\": [\n%s\n ]\n}\n", g.cases, g.bad, g.json.c_str());
    fprintf(stderr, "ref_math_kat: %ld cases, %ld mismatches\n", g.cases, g.bad);
    return g.bad ? 1 : 0;
}
