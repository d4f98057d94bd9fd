/* ORACLE (test infrastructure) -- textures, BSDFs, lights, post-intersection, restated from
 *   texture.cpp:10-270, bsdf.cpp:20-92 (diffuse), :281-417 (normalmap), :1157-1418 (kiss),
 *   ggx_brdf.h:15-170, light.cpp:7-66, mesh.cpp:24-53,108-133, dpdf.h:35-104,
 *   accel.cpp:113-236 (post-intersection).
 * Texture filtering arithmetic lives in OpenImageIO (not in /root/reference): PARITY UNPINNED;
 * restated as: finest mip level, periodic wrap, bicubic B-spline interpolation (OIIO's
 * "smart bicubic" default when the footprint is zero), t flipped by the caller. */
#ifndef KZO_SHADING_H
#define KZO_SHADING_H
#include "kzo_math.h"
#include "kzo_sampler.h"
#include "kzo_accel.h"
#include <vector>

namespace kzo {

struct Image { int w, h; std::vector<float> rgb; };

struct MeshData {
    std::vector<float> P, N, UV;
    std::vector<uint32_t> F;
    uint32_t nV = 0, nF = 0;
    int bsdf = -1, light = -1;
    /* DiscretePDF, dpdf.h:35-81 */
    std::vector<float> cdf;
    float normalization = 0.f;
    V3 pos(uint32_t i) const { return V3(P[3 * i], P[3 * i + 1], P[3 * i + 2]); }
    V3 nrm(uint32_t i) const { return V3(N[3 * i], N[3 * i + 1], N[3 * i + 2]); }
    V2 uv(uint32_t i) const { return V2{UV[2 * i], UV[2 * i + 1]}; }
};

/* mesh.h:18-56 */
struct Intersection {
    V3 p; float t = 0.f; V2 uv{0, 0};
    Frame shFrame, geoFrame;
    int mesh = -1;
    V3 dpdu, dpdv;
    float accumulatedRoughness = 0.f;
    V3 toLocal(const V3 &d) const { return shFrame.toLocal(d); }
    V3 toWorld(const V3 &d) const { return shFrame.toWorld(d); }
};

enum Measure { EUnknownMeasure = 0, ESolidAngle = 1, EDiscrete = 2 };

/* bsdf.h:17-53 */
struct BSDFQueryRecord {
    Intersection its;
    V3 wi, wo;
    float eta = 1.f;
    int measure = EUnknownMeasure;
    V2 uv{0, 0};
    explicit BSDFQueryRecord(const V3 &wi_) : wi(wi_) {}
    BSDFQueryRecord(const V3 &wi_, const V3 &wo_, int m) : wi(wi_), wo(wo_), measure(m) {}
};

struct SceneData {
    std::vector<MeshData> meshes;
    std::vector<kz_bsdf_desc> bsdfs;
    std::vector<kz_texture_desc> textures;
    std::vector<Image> images;
    std::vector<kz_light_desc> lights;
    std::vector<int> lightMeshes;     /* scene.cpp:42-46 */
    int background = -1;
    kz_camera_desc camera;
    SamplerCfg sampler;
    std::vector<uint16_t> blueNoise;
    std::vector<uint32_t> pmj;
    kz_integrator_desc integrator;
    kz_filter_desc filter;
    Accel accel;
};

/* ---- textures ------------------------------------------------------------------------- */
inline int wrapi(int i, int n) { int r = i % n; return r < 0 ? r + n : r; }

/* B-spline weights for fractional offset f */
inline void bsplineWeights(float f, float w[4]) {
    float one_f = 1.0f - f;
    w[0] = (one_f * one_f * one_f) / 6.0f;
    w[1] = 2.0f / 3.0f - 0.5f * f * f * (2.0f - f);
    w[2] = 2.0f / 3.0f - 0.5f * one_f * one_f * (2.0f - one_f);
    w[3] = (f * f * f) / 6.0f;
}
inline V3 imageBicubic(const Image &im, float s, float t) {
    float x = s * im.w - 0.5f, y = t * im.h - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    int ix = (int)fx, iy = (int)fy;
    float wx[4], wy[4];
    bsplineWeights(x - fx, wx);
    bsplineWeights(y - fy, wy);
    V3 acc(0.f);
    for (int j = 0; j < 4; ++j) {
        int yy = wrapi(iy - 1 + j, im.h);
        V3 row(0.f);
        for (int i = 0; i < 4; ++i) {
            int xx = wrapi(ix - 1 + i, im.w);
            const float *p = &im.rgb[3 * ((size_t)yy * im.w + xx)];
            row += V3(p[0], p[1], p[2]) * wx[i];
        }
        acc += row * wy[j];
    }
    return acc;
}

inline V3 evalTextureUV(const SceneData &sc, int node, V2 uv);
inline V3 evalTextureDir(const SceneData &sc, int node, V3 dir);

inline V3 evalTextureUV(const SceneData &sc, int node, V2 uv) {
    const kz_texture_desc &t = sc.textures[node];
    switch (t.type) {
        case KZ_TEX_CONSTANT: return V3(t.color[0], t.color[1], t.color[2]);              /* texture.cpp:15-17 */
        case KZ_TEX_IMAGE: {                                                                /* texture.cpp:46-64 */
            V3 c = imageBicubic(sc.images[t.image], uv.x * t.scale, (1.0f - uv.y) * t.scale);
            return t.srgb ? toLinearRGB(c) : c;
        }
        case KZ_TEX_BACKGROUND:                                                             /* texture.cpp:114-119 */
            return t.child[0] >= 0 ? t.a * evalTextureUV(sc, t.child[0], uv) : V3(0.f);
        case KZ_TEX_COLORRAMP: {                                                            /* texture.cpp:155-166 */
            if (t.child[0] < 0) return V3(0.f);
            V3 c = evalTextureUV(sc, t.child[0], uv);
            auto ramp = [&](float in) { in = clampf(in, 0.0f, 1.0f); return t.a + (t.b - t.a) * in; };
            return V3(ramp(c.x), ramp(c.y), ramp(c.z));
        }
        case KZ_TEX_BLEND: {                                                                /* texture.cpp:213-230 */
            V3 mask(0.5f), in1(0.f), in2(1.f);
            if (t.child[0] >= 0) mask = evalTextureUV(sc, t.child[0], uv);
            if (t.child[1] >= 0) in1 = evalTextureUV(sc, t.child[1], uv);
            if (t.child[2] >= 0) in2 = evalTextureUV(sc, t.child[2], uv);
            if (t.mode == KZ_BLEND_MIX)
                return V3(lerpf(mask.x, in1.x, in2.x), lerpf(mask.x, in1.y, in2.y), lerpf(mask.x, in1.z, in2.z));
            if (t.mode == KZ_BLEND_MULTIPLY) return V3(in1.x * in2.x, in1.y * in2.y, in1.z * in2.z);
            return V3(0.f);
        }
    }
    return V3(0.f);
}

inline V3 evalTextureDir(const SceneData &sc, int node, V3 dir) {
    const kz_texture_desc &t = sc.textures[node];
    switch (t.type) {
        case KZ_TEX_CONSTANT: return V3(t.color[0], t.color[1], t.color[2]);              /* texture.cpp:18-20 */
        case KZ_TEX_IMAGE: {                                                                /* texture.cpp:66-81 */
            /* OIIO environment(): lat-long, y up; no colour-space conversion in the reference */
            float s = std::atan2(-dir.x, dir.z) / (2.0f * kPi) + 0.5f;
            float tt = 0.5f - std::atan2(dir.y, std::hypot(dir.z, -dir.x)) / kPi;
            if (std::isnan(s)) s = 0.0f;
            if (std::isnan(tt)) tt = 0.0f;
            return imageBicubic(sc.images[t.image], s, tt);
        }
        case KZ_TEX_BACKGROUND:                                                             /* texture.cpp:120-125 */
            return t.child[0] >= 0 ? t.a * evalTextureDir(sc, t.child[0], dir) : V3(0.f);
        default: return V3(0.f);   /* Texture::eval(dir) base class, texture.h:12 */
    }
}

/* scene.cpp:54-79 */
inline V3 backgroundColor(const SceneData &sc, V3 dir) {
    if (sc.background < 0) return V3(0.f);
    if (std::isnan(dir.x) || std::isnan(dir.y) || std::isnan(dir.z)) return V3(0.f);
    return evalTextureDir(sc, sc.background, dir);
}

/* ---- GGX helpers, ggx_brdf.h:15-120 ----------------------------------------------------- */
inline V3 schlickFresnel(V3 f0, float cosTheta) {          /* lerp(f0, 1, pow(1-c,5)) = t*c1 + (1-t)*c2 */
    float w = std::pow(1.0f - cosTheta, 5.0f);
    return f0 * 1.0f + (V3(1.f) - f0) * w;
}
inline V2 roughnessToAlpha(float roughness, float anisotropy) {
    float alpha = std::max(0.001f, sqr(roughness));
    return V2{alpha * (1.0f + anisotropy), alpha * (1.0f - anisotropy)};
}
inline float ggxLambda(V3 v, V2 a) {
    float squared = (sqr(a.x) * sqr(v.x) + sqr(a.y) * sqr(v.y)) / sqr(v.z);
    return (-1.0f + std::sqrt(1.0f + squared)) * 0.5f;
}
inline float smithG1(V3 V, V3 H, V2 a) {
    if (dot(V, H) <= 0.0f) return 0.0f;
    return 1.0f / (1.0f + ggxLambda(V, a));
}
inline float smithG2(V3 V, V3 L, V3 H, V2 a) {
    if (dot(V, H) <= 0.0f || dot(L, H) < 0.0f) return 0.0f;
    return 1.0f / (1.0f + ggxLambda(V, a) + ggxLambda(L, a));
}
inline float ggxNDF(V3 H, V2 a) {
    float ellipse = sqr(H.x) / sqr(a.x) + sqr(H.y) / sqr(a.y) + sqr(H.z);
    return 1.0f / (kPi * a.x * a.y * sqr(ellipse));
}
inline float ggxVNDF(V3 V, V3 H, V2 a) {
    float VDotH = dot(V, H);
    if (VDotH <= 0.0f) return 0.0f;
    float D = ggxNDF(H, a);
    float G1 = smithG1(V, H, a);
    return D * G1 * VDotH / V.z;
}
inline V3 sampleGGXVNDF(V3 V, V2 a, V2 rnd) {
    V3 Vh = normalized(V3(a.x * V.x, a.y * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    V3 T1 = lensq > 0.0f ? V3(-Vh.y, Vh.x, 0.0f) / std::sqrt(lensq) : V3(1.0f, 0.0f, 0.0f);
    V3 T2 = normalized(cross(Vh, T1));
    float r = std::sqrt(rnd.x);
    float phi = 2.0f * kPi * rnd.y;
    float t1 = r * std::cos(phi);
    float t2 = r * std::sin(phi);
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * std::sqrt(1.0f - t1 * t1) + s * t2;
    V3 Nh = t1 * T1 + t2 * T2 + std::sqrt(std::max(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalized(V3(a.x * Nh.x, a.y * Nh.y, std::max(1e-6f, Nh.z)));
}
/* ggx_brdf.h:151-170 */
inline V3 ggxSmithBRDF(V3 V, V3 L, V3 f0, float roughness, float anisotropy) {
    if (V.z * L.z < 0.0f) return V3(0.0f);
    V2 a = roughnessToAlpha(roughness, anisotropy);
    V3 H = normalized(V + L);
    float D = ggxNDF(H, a);
    float G = smithG2(V, L, H, a);
    V3 F = schlickFresnel(f0, dot(V, H));
    float denom = 4.0f * std::fabs(V.z) * std::fabs(L.z);
    return (D * G) * F / denom;
}

/* ---- BSDFs --------------------------------------------------------------------------------- */
inline V3 bsdfEval(const SceneData &sc, int b, const BSDFQueryRecord &bRec);
inline float bsdfPdf(const SceneData &sc, int b, const BSDFQueryRecord &bRec);
inline V3 bsdfSample(const SceneData &sc, int b, BSDFQueryRecord &bRec, float s1, V2 s2);

inline float schlickWeight(float x) {       /* bsdf.cpp:1175-1179 */
    x = clampf(1.f - x, 0.f, 1.f);
    float x2 = x * x;
    return x2 * x2 * x;
}
inline V3 lerpColor(V3 c1, V3 c2, float t) { return (1.f - t) * c1 + t * c2; }   /* bsdf.cpp:1181-1183 */

inline V3 kissEval(const SceneData &sc, const kz_bsdf_desc &m, const BSDFQueryRecord &bRec) {   /* bsdf.cpp:1215-1267 */
    if (bRec.wi.z <= 0 || bRec.wo.z <= 0) return V3(0.0f);
    V3 V = bRec.wi, L = bRec.wo, H = normalized(V + L);
    V3 Cdlin = evalTextureUV(sc, m.base_color, bRec.uv);
    float metallic = evalTextureUV(sc, m.metallic, bRec.uv).x;
    float roughness = std::min(1.f, evalTextureUV(sc, m.roughness, bRec.uv).x + bRec.its.accumulatedRoughness);
    float Cdlum = luminance(Cdlin);
    V3 Ctint = Cdlum > 0.f ? Cdlin / Cdlum : V3(1.f);
    V3 Ctintmix = (0.08f * m.specular) * lerpColor(V3(1.f), Ctint, m.specular_tint);
    V3 Cspec0 = lerpColor(Ctintmix, Cdlin, metallic);
    float FL = schlickWeight(L.z), FV = schlickWeight(V.z), FH = schlickWeight(dot(L, H));
    float cosThetaD = dot(V, H);
    float Lambert = (1.f - 0.5f * FL) * (1.f - 0.5f * FV);
    float RR = 2.f * roughness * cosThetaD * cosThetaD;
    float retro = RR * (FL + FV + FL * FV * (RR - 1.f));
    V3 Csheen = lerpColor(V3(1.f), Ctint, m.sheen_tint);
    V3 Fsheen = (FH * m.sheen) * Csheen;
    V3 specTerm = ggxSmithBRDF(V, L, Cspec0, roughness, m.anisotropy);
    float ccR = lerpf(m.clearcoat_roughness, .01f, .3f);
    V3 coatTerm = (0.25f * m.clearcoat) * ggxSmithBRDF(V, L, V3(0.04f), ccR, m.anisotropy);
    return ((1.f - metallic) * (Cdlin * kInvPi * (Lambert + retro) + Fsheen) + (specTerm + coatTerm)) * bRec.wo.z;
}
inline float kissPdf(const SceneData &sc, const kz_bsdf_desc &m, const BSDFQueryRecord &bRec) {  /* bsdf.cpp:1269-1299 */
    if (bRec.wi.z <= 0 || bRec.wo.z <= 0) return 0.0f;
    float metallic = evalTextureUV(sc, m.metallic, bRec.uv).x;
    float diffuse = (1.f - metallic) * 0.5f;
    float GTR2 = 1.f / (1.f + m.clearcoat);
    V3 H = normalized(bRec.wi + bRec.wo);
    float jacobian = 4.0f * dot(bRec.wi, H);
    float roughness = std::min(1.f, evalTextureUV(sc, m.roughness, bRec.uv).x + bRec.its.accumulatedRoughness);
    V2 alpha = roughnessToAlpha(roughness, m.anisotropy);
    float specPdf = ggxVNDF(bRec.wi, H, alpha) / jacobian;
    V2 coatalpha = roughnessToAlpha(lerpf(m.clearcoat_roughness, .01f, .3f), 0.f);
    float coatPdf = ggxVNDF(bRec.wi, H, coatalpha) / jacobian;
    return diffuse * kInvPi * bRec.wo.z + (1.f - diffuse) * (GTR2 * specPdf + (1.f - GTR2) * coatPdf);
}
inline V3 kissSample(const SceneData &sc, const kz_bsdf_desc &m, BSDFQueryRecord &bRec, float sample1, V2 sample2) { /* bsdf.cpp:1301-1371 */
    if (bRec.wi.z <= 0) return V3(0.0f);
    bRec.measure = ESolidAngle;
    bRec.eta = 1.0f;
    float metallic = evalTextureUV(sc, m.metallic, bRec.uv).x;
    float diffuse = (1.f - metallic) * 0.5f;
    if (sample1 < diffuse) {
        bRec.wo = squareToCosineHemisphere(sample2);
    } else {
        float sample = (sample1 - diffuse) / (1.f - diffuse);
        float GTR2 = 1.f / (1.f + m.clearcoat);
        V3 H;
        bool flip = bRec.wi.z <= 0.f;
        if (sample < GTR2) {
            float roughness = evalTextureUV(sc, m.roughness, bRec.uv).x;      /* NO bias here, :1323 */
            V2 alpha = roughnessToAlpha(roughness, m.anisotropy);
            H = sampleGGXVNDF(flip ? -bRec.wi : bRec.wi, alpha, sample2);
        } else {
            V2 alpha = roughnessToAlpha(lerpf(m.clearcoat_roughness, 0.01f, .3f), 0.f);
            H = sampleGGXVNDF(flip ? -bRec.wi : bRec.wi, alpha, sample2);
        }
        H = flip ? -H : H;
        bRec.wo = normalized(reflect(bRec.wi, H));
    }
    bool invalid = std::isnan(bRec.wo.x) || std::isnan(bRec.wo.y) || std::isnan(bRec.wo.z);
    if (bRec.wo.z <= 0 || kissPdf(sc, m, bRec) <= kEpsilon || invalid) return V3(0.f);
    return kissEval(sc, m, bRec) / kissPdf(sc, m, bRec);
}

/* ---- the remaining BSDF plugins ------------------------------------------------- */
/* common.cpp:447-476 */
inline float fresnelExtInt(float cosThetaI, float extIOR, float intIOR) {
    float etaI = extIOR, etaT = intIOR;
    if (extIOR == intIOR) return 0.0f;
    if (cosThetaI < 0.0f) { std::swap(etaI, etaT); cosThetaI = -cosThetaI; }
    float eta = etaI / etaT, sinThetaTSqr = eta * eta * (1 - cosThetaI * cosThetaI);
    if (sinThetaTSqr > 1.0f) return 1.0f;
    float cosThetaT = std::sqrt(1.0f - sinThetaTSqr);
    float Rs = (etaI * cosThetaI - etaT * cosThetaT) / (etaI * cosThetaI + etaT * cosThetaT);
    float Rp = (etaT * cosThetaI - etaI * cosThetaT) / (etaT * cosThetaI + etaI * cosThetaT);
    return (Rs * Rs + Rp * Rp) / 2.0f;
}
/* common.cpp:493-523 */
inline float fresnelDielectric(float cosThetaI_, float eta, float &cosThetaT_) {
    float scale = (cosThetaI_ > 0.f) ? 1 / eta : eta, cosThetaTSqr = 1 - (1 - cosThetaI_ * cosThetaI_) * (scale * scale);
    if (cosThetaTSqr <= 0.0f) { cosThetaT_ = 0.0f; return 1.0f; }
    float cosThetaI = std::fabs(cosThetaI_), cosThetaT = std::sqrt(cosThetaTSqr);
    float Rs = (cosThetaI - eta * cosThetaT) / (cosThetaI + eta * cosThetaT);
    float Rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    cosThetaT_ = (cosThetaI_ > 0) ? -cosThetaT : cosThetaT;
    return 0.5f * (Rs * Rs + Rp * Rp);
}
/* common.cpp:525-534 */
inline V3 refractDir(const V3 &wi, const V3 &n, float eta) {
    float cosThetaI = dot(wi, n);
    if (cosThetaI < 0) eta = 1.0f / eta;
    float cosThetaT2 = 1 - (1 - cosThetaI * cosThetaI) * (eta * eta);
    if (cosThetaT2 <= 0.0f) return V3(0.0f);
    float sign = cosThetaI >= 0.0f ? 1.0f : -1.0f;
    return n * (-cosThetaI * eta + sign * std::sqrt(cosThetaT2)) + wi * eta;
}
/* frame.h:63-68 */
inline float tanThetaLocal(const V3 &v) { float temp = 1 - v.z * v.z; return temp <= 0.0f ? 0.0f : std::sqrt(temp) / v.z; }
/* warp.cpp:121-130 */
inline V3 squareToBeckmann(V2 sample, float alpha) {
    float phi = 2 * kPi * sample.x;
    float theta = std::atan(alpha * std::sqrt(std::log(1 / (1 - sample.y))));
    return V3(std::sin(theta) * std::cos(phi), std::sin(theta) * std::sin(phi), std::cos(theta));
}
/* warp.cpp:127-130.  `pow(tan(theta),2)` and `pow(cos(theta),3)` have INT exponents, so std::pow promotes to double and the whole
 * quotient is evaluated in double before the return converts it to float (found by oracle/ref_math_kat.cpp, which runs the
 * reference's own body: the all-float restatement differed in the last bit for 58 % of the inputs). */
inline float squareToBeckmannPdf(const V3 &m, float alpha) {
    float theta = std::acos(m.z / norm(m));
    bool ok = std::fabs(norm(m) - 1) < kEpsilon && m.z >= 0;
    const double e = std::exp(-std::pow((double)std::tan(theta), 2) / (double)(alpha * alpha));
    const double den = (double)(kPi * alpha * alpha) * std::pow((double)std::cos(theta), 3);
    return (float)((double)(ok ? 1 : 0) * e / den);
}
/* bsdf.cpp:730-762 (identical copies at :848-878 and :1103-1133) */
inline float evalBeckmann(const V3 &m, float alpha) {
    float temp = tanThetaLocal(m) / alpha, ct = m.z, ct2 = ct * ct;
    return std::exp(-temp * temp) / (kPi * alpha * alpha * ct2 * ct2);
}
inline float smithBeckmannG1(const V3 &v, const V3 &m, float alpha) {
    if (dot(v, m) * v.z <= 0.0f) return 0.0f;
    float tanTheta = std::fabs(tanThetaLocal(v));
    if (tanTheta == 0.0f) return 1.0f;
    float a = 1.0f / (alpha * tanTheta);
    if (a >= 1.6f) return 1.0f;
    float aSqr = a * a;
    return (3.535f * a + 2.181f * aSqr) / (1.0f + 2.276f * a + 2.577f * aSqr);
}
/* bsdf.cpp:718-727 */
inline V3 fresnelCond(float c, V3 eta, V3 k) {
    V3 tmp_f = eta * eta + k * k;
    V3 tmp = tmp_f * (c * c);
    V3 Rparl2 = (tmp - (2.f * eta * c) + V3(1.f)) / (tmp + (2.f * eta * c) + V3(1.f));
    V3 Rperp2 = (tmp_f - (2.f * eta * c) + V3(c * c)) / (tmp_f + (2.f * eta * c) + V3(c * c));
    return (Rparl2 + Rperp2) / 2.0f;
}
/* ggx_brdf.h:151-170 with the Fresnel term (evaluateGGXSmithBRDF) is ggxSmithBRDF above */

inline V3 extraEval(const SceneData &sc, const kz_bsdf_desc &m, const BSDFQueryRecord &bRec) {
    const V3 wi = bRec.wi, wo = bRec.wo;
    switch (m.type) {
        case KZ_BSDF_DIELECTRIC: case KZ_BSDF_MIRROR: return V3(0.f);                      /* discrete: bsdf.cpp:108-116,166-174 */
        case KZ_BSDF_LAMBERTIAN:                                                             /* bsdf.cpp:211-221 */
            if (bRec.measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return V3(0.f);
            return evalTextureUV(sc, m.base_color, bRec.uv) * kInvPi * wo.z;
        case KZ_BSDF_GGX: {                                                                  /* bsdf.cpp:640-647 */
            if (wi.z <= 0 || wo.z <= 0) return V3(0.f);
            return ggxSmithBRDF(wi, wo, evalTextureUV(sc, m.base_color, bRec.uv), m.alpha, m.anisotropy) * wo.z;
        }
        case KZ_BSDF_ROUGHCONDUCTOR: {                                                       /* bsdf.cpp:765-775 */
            if (wi.z <= 0 || wo.z <= 0) return V3(0.f);
            V3 wh = normalized(wi + wo);
            V3 F = fresnelCond(dot(wh, wo), V3(m.eta[0], m.eta[1], m.eta[2]), V3(m.k[0], m.k[1], m.k[2]));
            float D = evalBeckmann(wh, m.alpha);
            float G = smithBeckmannG1(wi, wh, m.alpha) * smithBeckmannG1(wo, wh, m.alpha);
            return D * F * G / (4.f * wi.z);
        }
        case KZ_BSDF_ROUGHPLASTIC: {                                                         /* bsdf.cpp:881-893 */
            if (wi.z <= 0 || wo.z <= 0) return V3(0.f);
            V3 kd(m.albedo[0], m.albedo[1], m.albedo[2]);
            float ks = 1 - maxcoeff(kd);
            V3 wh = normalized(wi + wo);
            float D = evalBeckmann(wh, m.alpha);
            float F = fresnelExtInt(dot(wh, wo), m.ext_ior, m.int_ior);
            float G = smithBeckmannG1(wo, wh, m.alpha) * smithBeckmannG1(wi, wh, m.alpha);
            return kd * kInvPi * wo.z + V3(ks * (D * F * G) / (4.f * wi.z));
        }
        case KZ_BSDF_ROUGHDIELECTRIC: {                                                      /* bsdf.cpp:969-1014 */
            if (wi.z == 0) return V3(0.f);
            const float m_eta = m.int_ior / m.ext_ior, m_invEta = m.ext_ior / m.int_ior;
            float cosThetaI = wi.z, cosThetaO = wo.z;
            bool reflectS = cosThetaI * cosThetaO > 0.f;
            float eta = cosThetaI > 0.f ? m_eta : m_invEta;
            V3 wm = reflectS ? normalized(wi + wo) : normalized(wi + wo * eta);
            wm = wm * (wm.z > 0.f ? 1.f : -1.f);                                             /* math::sign, common.h:266-268 */
            float ct;
            float F = fresnelDielectric(dot(wi, wm), m_eta, ct);
            float D = evalBeckmann(wm, m.alpha);
            float G = smithBeckmannG1(wo, wm, m.alpha) * smithBeckmannG1(wi, wm, m.alpha);
            if (reflectS) return V3((F * G * D) / (4.f * std::fabs(cosThetaI)));
            float denom = dot(wi, wm) + eta * dot(wo, wm);
            float value = ((1 - F) * D * G * eta * eta * dot(wi, wm) * dot(wo, wm)) / (cosThetaI * sqr(denom));
            return V3(std::fabs(value));
        }
    }
    return V3(0.f);
}
inline float extraPdf(const SceneData &sc, const kz_bsdf_desc &m, const BSDFQueryRecord &bRec) {
    (void)sc;
    const V3 wi = bRec.wi, wo = bRec.wo;
    switch (m.type) {
        case KZ_BSDF_DIELECTRIC: case KZ_BSDF_MIRROR: return 0.f;
        case KZ_BSDF_LAMBERTIAN:                                                             /* bsdf.cpp:223-239 */
            if (bRec.measure != ESolidAngle || wi.z <= 0 || wo.z <= 0) return 0.f;
            return kInvPi * wo.z;
        case KZ_BSDF_GGX: {                                                                  /* bsdf.cpp:649-657 */
            if (wi.z <= 0 || wo.z <= 0) return 0.f;
            V3 H = normalized(wi + wo);
            return ggxVNDF(wi, H, roughnessToAlpha(m.alpha, m.anisotropy)) / (4.0f * dot(wi, H));
        }
        case KZ_BSDF_ROUGHCONDUCTOR: {                                                       /* bsdf.cpp:778-785 */
            if (wi.z <= 0 || wo.z <= 0) return 0.f;
            V3 wh = normalized(wi + wo);
            return evalBeckmann(wh, m.alpha) * wh.z * (1.f / (4.f * dot(wh, wo)));
        }
        case KZ_BSDF_ROUGHPLASTIC: {                                                         /* bsdf.cpp:895-903 */
            if (wi.z <= 0 || wo.z <= 0) return 0.f;
            float ks = 1 - std::max(m.albedo[0], std::max(m.albedo[1], m.albedo[2]));
            V3 wh = normalized(wi + wo);
            float Jh = 1.f / (4.f * std::fabs(dot(wh, wo)));
            return ks * evalBeckmann(wh, m.alpha) * wh.z * Jh + (1 - ks) * wo.z * kInvPi;
        }
        case KZ_BSDF_ROUGHDIELECTRIC: {                                                      /* bsdf.cpp:1016-1048 */
            const float m_eta = m.int_ior / m.ext_ior, m_invEta = m.ext_ior / m.int_ior;
            float cosThetaI = wi.z, cosThetaO = wo.z;
            bool reflectS = cosThetaI * cosThetaO > 0.f;
            float eta = cosThetaI > 0.f ? m_eta : m_invEta;
            V3 wm; float dwm_dwo;
            if (reflectS) { wm = normalized(wi + wo); dwm_dwo = 1.0f / (4.0f * dot(wo, wm)); }
            else {
                wm = normalized(wi + wo * eta);
                float sqrtDenom = dot(wi, wm) + eta * dot(wo, wm);
                dwm_dwo = (eta * eta * dot(wo, wm)) / (sqrtDenom * sqrtDenom);
            }
            wm = wm * (wm.z > 0.f ? 1.f : -1.f);
            float ct;
            float F = fresnelDielectric(dot(wi, wm), m_eta, ct);
            float prob = evalBeckmann(wm, m.alpha) * wm.z;
            prob *= reflectS ? F : (1 - F);
            return std::fabs(prob * dwm_dwo);
        }
    }
    return 0.f;
}
inline V3 extraSample(const SceneData &sc, const kz_bsdf_desc &m, BSDFQueryRecord &bRec, float sample1, V2 sample2) {
    switch (m.type) {
        case KZ_BSDF_DIELECTRIC: {                                                           /* bsdf.cpp:118-144 */
            bRec.measure = EDiscrete;
            float fr = fresnelExtInt(bRec.wi.z, m.ext_ior, m.int_ior);
            if (sample1 < fr) { bRec.wo = V3(-bRec.wi.x, -bRec.wi.y, bRec.wi.z); bRec.eta = 1.f; return V3(1.0f); }
            V3 n(0.f, 0.f, 1.f);
            float factor = m.int_ior / m.ext_ior;
            if (bRec.wi.z < 0.f) { factor = m.ext_ior / m.int_ior; n.z = -1.0f; }
            bRec.wo = refractDir(-bRec.wi, n, factor);
            bRec.eta = m.int_ior / m.ext_ior;
            return V3(1.0f);
        }
        case KZ_BSDF_MIRROR:                                                                 /* bsdf.cpp:176-191 */
            if (bRec.wi.z <= 0) return V3(0.f);
            bRec.wo = V3(-bRec.wi.x, -bRec.wi.y, bRec.wi.z);
            bRec.measure = EDiscrete; bRec.eta = 1.0f;
            return V3(1.0f);
        case KZ_BSDF_LAMBERTIAN:                                                             /* bsdf.cpp:241-257 */
            if (bRec.wi.z <= 0) return V3(0.f);
            bRec.measure = ESolidAngle;
            bRec.wo = squareToCosineHemisphere(sample2);
            bRec.eta = 1.0f;
            return evalTextureUV(sc, m.base_color, bRec.uv);
        case KZ_BSDF_GGX: {                                                                  /* bsdf.cpp:659-670 + ggx_brdf.h:175-203; measure/eta stay at their defaults */
            if (bRec.wi.z <= 0) return V3(0.f);
            V3 albedo = evalTextureUV(sc, m.base_color, bRec.uv);
            V2 alpha = roughnessToAlpha(m.alpha, m.anisotropy);
            V3 H = sampleGGXVNDF(bRec.wi, alpha, sample2);                                   /* flip is dead: wi.z > 0 */
            bRec.wo = reflect(bRec.wi, H);
            float pdf = ggxVNDF(bRec.wi, H, alpha) / (4.0f * dot(bRec.wi, H));
            V3 color = ggxSmithBRDF(bRec.wi, bRec.wo, albedo, m.alpha, m.anisotropy);
            if (bRec.wo.z <= 0) return V3(0.f);
            return color * bRec.wo.z / pdf;
        }
        case KZ_BSDF_ROUGHCONDUCTOR: {                                                       /* bsdf.cpp:788-799 */
            if (bRec.wi.z <= 0) return V3(0.f);
            V3 wh = squareToBeckmann(sample2, m.alpha);
            bRec.wo = normalized(reflect(bRec.wi, wh));
            if (bRec.wo.z <= 0) return V3(0.f);
            return extraEval(sc, m, bRec) / extraPdf(sc, m, bRec);
        }
        case KZ_BSDF_ROUGHPLASTIC: {                                                         /* bsdf.cpp:905-918 */
            if (bRec.wi.z <= 0) return V3(0.f);
            float ks = 1 - std::max(m.albedo[0], std::max(m.albedo[1], m.albedo[2]));
            if (sample1 < ks) {
                V3 wh = squareToBeckmann(sample2, m.alpha);
                bRec.wo = normalized((2.f * dot(wh, bRec.wi) * wh) - bRec.wi);
            } else bRec.wo = squareToCosineHemisphere(sample2);
            if (bRec.wo.z <= 0) return V3(0.f);
            return extraEval(sc, m, bRec) / extraPdf(sc, m, bRec);
        }
        case KZ_BSDF_ROUGHDIELECTRIC: {                                                      /* bsdf.cpp:1050-1096 */
            const float m_eta = m.int_ior / m.ext_ior, m_invEta = m.ext_ior / m.int_ior;
            float alpha = m.alpha * (1.2f - 0.2f * std::sqrt(std::fabs(bRec.wi.z)));
            V3 wm = squareToBeckmann(sample2, alpha);
            float pdf = squareToBeckmannPdf(wm, alpha);
            if (pdf == 0.f) return V3(0.f);
            float cosThetaT;
            float F = fresnelDielectric(dot(bRec.wi, wm), m_eta, cosThetaT);
            if (!(sample1 > F)) {
                bRec.wo = reflect(bRec.wi, wm);
                bRec.eta = 1.0f;
                if (bRec.wi.z * bRec.wo.z <= 0) return V3(0.f);
            } else {
                if (cosThetaT == 0) return V3(0.f);
                float e = cosThetaT < 0 ? 1.f / m_eta : m_eta;                               /* RoughDielectric::refract, bsdf.cpp:1135-1139 */
                bRec.wo = wm * (dot(bRec.wi, wm) * e + cosThetaT) - bRec.wi * e;
                bRec.eta = cosThetaT < 0.f ? m_eta : m_invEta;
                if (bRec.wi.z * bRec.wo.z >= 0) return V3(0.f);
            }
            float D = evalBeckmann(wm, alpha);
            float G = smithBeckmannG1(bRec.wo, wm, alpha) * smithBeckmannG1(bRec.wi, wm, alpha);
            return V3(std::fabs(D * G * dot(bRec.wi, wm) / (pdf * bRec.wi.z)));
        }
    }
    return V3(0.f);
}

/* bsdf.cpp:366-374 */
inline Frame normalMapFrame(const Intersection &its, V3 n) {
    Frame r;
    r.n = normalized(its.shFrame.toWorld(n));
    r.s = normalized(its.dpdu - r.n * dot(r.n, its.dpdu));
    r.t = normalized(cross(r.n, r.s));
    return r;
}

inline V3 bsdfEval(const SceneData &sc, int b, const BSDFQueryRecord &bRec) {
    const kz_bsdf_desc &m = sc.bsdfs[b];
    switch (m.type) {
        case KZ_BSDF_DIFFUSE:                                                  /* bsdf.cpp:27-36 */
            if (bRec.measure != ESolidAngle || bRec.wi.z <= 0 || bRec.wo.z <= 0) return V3(0.0f);
            return V3(m.albedo[0], m.albedo[1], m.albedo[2]) * kInvPi * bRec.wo.z;
        case KZ_BSDF_KISS: return kissEval(sc, m, bRec);
        case KZ_BSDF_NORMALMAP: {                                              /* bsdf.cpp:290-312 */
            const Intersection &its = bRec.its;
            V3 rgb = evalTextureUV(sc, m.normal_map, its.uv);
            V3 n(2 * rgb.x - 1, 2 * rgb.y - 1, 2 * rgb.z - 1);
            if (bRec.wi.z > 0 && bRec.wo.z > 0 && dot(n, bRec.wi) <= 0) return bsdfEval(sc, m.nested, bRec);
            Intersection perturbed(its);
            perturbed.shFrame = normalMapFrame(its, normalized(n));
            BSDFQueryRecord pq(perturbed.toLocal(its.toWorld(bRec.wi)), perturbed.toLocal(its.toWorld(bRec.wo)), bRec.measure);
            if (bRec.wo.z * pq.wo.z <= 0) return V3(0.0f);
            pq.uv = bRec.uv; pq.measure = bRec.measure; pq.eta = bRec.eta;
            return bsdfEval(sc, m.nested, pq);
        }
        default: return extraEval(sc, m, bRec);
    }
    return V3(0.f);
}
inline float bsdfPdf(const SceneData &sc, int b, const BSDFQueryRecord &bRec) {
    const kz_bsdf_desc &m = sc.bsdfs[b];
    switch (m.type) {
        case KZ_BSDF_DIFFUSE:                                                  /* bsdf.cpp:39-55 */
            if (bRec.measure != ESolidAngle || bRec.wi.z <= 0 || bRec.wo.z <= 0) return 0.0f;
            return kInvPi * bRec.wo.z;
        case KZ_BSDF_KISS: return kissPdf(sc, m, bRec);
        case KZ_BSDF_NORMALMAP: {                                              /* bsdf.cpp:314-336 */
            const Intersection &its = bRec.its;
            V3 rgb = evalTextureUV(sc, m.normal_map, its.uv);
            V3 n(2 * rgb.x - 1, 2 * rgb.y - 1, 2 * rgb.z - 1);
            if (bRec.wi.z > 0 && bRec.wo.z > 0 && dot(n, bRec.wi) <= 0) return bsdfPdf(sc, m.nested, bRec);
            Intersection perturbed(its);
            perturbed.shFrame = normalMapFrame(its, normalized(n));
            BSDFQueryRecord pq(perturbed.toLocal(its.toWorld(bRec.wi)), perturbed.toLocal(its.toWorld(bRec.wo)), bRec.measure);
            if (bRec.wo.z * pq.wo.z <= 0) return 0.0f;
            pq.uv = bRec.uv; pq.measure = bRec.measure; pq.eta = bRec.eta;
            return bsdfPdf(sc, m.nested, pq);
        }
        default: return extraPdf(sc, m, bRec);
    }
    return 0.f;
}
inline V3 bsdfSample(const SceneData &sc, int b, BSDFQueryRecord &bRec, float s1, V2 s2) {
    const kz_bsdf_desc &m = sc.bsdfs[b];
    switch (m.type) {
        case KZ_BSDF_DIFFUSE:                                                  /* bsdf.cpp:58-75 */
            if (bRec.wi.z <= 0) return V3(0.0f);
            bRec.measure = ESolidAngle;
            bRec.wo = squareToCosineHemisphere(s2);
            bRec.eta = 1.0f;
            return V3(m.albedo[0], m.albedo[1], m.albedo[2]);
        case KZ_BSDF_KISS: return kissSample(sc, m, bRec, s1, s2);
        case KZ_BSDF_NORMALMAP: {                                              /* bsdf.cpp:338-363 */
            const Intersection &its = bRec.its;
            V3 rgb = evalTextureUV(sc, m.normal_map, its.uv);
            V3 n(2 * rgb.x - 1, 2 * rgb.y - 1, 2 * rgb.z - 1);
            if (bRec.wi.z > 0 && dot(n, bRec.wi) <= 0) {
                bRec.eta = 1.0f;
                return bsdfSample(sc, m.nested, bRec, s1, s2);
            }
            Intersection perturbed(its);
            perturbed.shFrame = normalMapFrame(its, normalized(n));
            BSDFQueryRecord pq(perturbed.toLocal(its.toWorld(bRec.wi)));
            pq.uv = its.uv; pq.measure = bRec.measure; pq.eta = bRec.eta;
            V3 result = bsdfSample(sc, m.nested, pq, s1, s2);
            if (!iszero(result)) {
                bRec.wo = its.toLocal(perturbed.toWorld(pq.wo));
                bRec.eta = pq.eta;
                if (bRec.wo.z * pq.wo.z <= 0) return V3(0.0f);
            }
            return result;
        }
        default: return extraSample(sc, m, bRec, s1, s2);
    }
    return V3(0.f);
}
/* bsdf.h:125, bsdf.cpp:412,1397-1399 */
inline float bsdfRegularize(const SceneData &sc, int b, V2 uv) {
    const kz_bsdf_desc &m = sc.bsdfs[b];
    if (m.type == KZ_BSDF_KISS) return evalTextureUV(sc, m.roughness, uv).x;
    if (m.type == KZ_BSDF_NORMALMAP) return bsdfRegularize(sc, m.nested, uv);
    return 0.f;
}

/* ---- lights ------------------------------------------------------------------------------------ */
/* mesh.cpp:47-53 */
inline float surfaceArea(const MeshData &m, uint32_t f) {
    V3 p0 = m.pos(m.F[3 * f]), p1 = m.pos(m.F[3 * f + 1]), p2 = m.pos(m.F[3 * f + 2]);
    return 0.5f * norm(cross(p1 - p0, p2 - p0));
}
/* mesh.cpp:30-44 + dpdf.h:35-81 */
inline void buildLightCdf(MeshData &m) {
    m.cdf.clear();
    m.cdf.push_back(0.0f);
    for (uint32_t i = 0; i < m.nF; ++i) m.cdf.push_back(m.cdf.back() + surfaceArea(m, i));
    float sum = m.cdf.back();
    if (sum > 0) {
        m.normalization = 1.0f / sum;
        for (size_t i = 1; i < m.cdf.size(); ++i) m.cdf[i] *= m.normalization;
        m.cdf.back() = 1.0f;
    } else {
        m.normalization = 0.0f;
    }
}
/* dpdf.h:99-104 */
inline size_t cdfSample(const std::vector<float> &cdf, float v) {
    auto entry = std::lower_bound(cdf.begin(), cdf.end(), v);
    size_t index = (size_t)std::max((ptrdiff_t)0, (ptrdiff_t)(entry - cdf.begin() - 1));
    return std::min(index, cdf.size() - 2);
}
/* mesh.cpp:108-133 */
inline void meshSample(const MeshData &m, Sampler &sampler, V3 &p, V3 &n) {
    size_t index = cdfSample(m.cdf, sampler.next1D());
    float su0 = std::sqrt(sampler.next1D());
    float u = 1 - su0;
    float v = sampler.next1D() * su0;
    uint32_t i0 = m.F[3 * index], i1 = m.F[3 * index + 1], i2 = m.F[3 * index + 2];
    V3 p0 = m.pos(i0), p1 = m.pos(i1), p2 = m.pos(i2);
    p = p0 + u * (p1 - p0) + v * (p2 - p0);
    if (!m.N.empty()) {
        V3 n0 = m.nrm(i0), n1 = m.nrm(i1), n2 = m.nrm(i2);
        n = n0 + u * (n1 - n0) + v * (n2 - n0);      /* n.normalized() result discarded, :129 */
    } else {
        n = normalized(cross(p1 - p0, p2 - p0));
    }
}

/* light.h:11-43 */
struct LightQueryRecord {
    V3 ref, wi, p, n;
    V2 uv{0, 0};
    kz_ray shadowRay;
    float pdf = 0.f;
    explicit LightQueryRecord(const V3 &ref_) : ref(ref_) {}
    LightQueryRecord(const V3 &ref_, const V3 &p_, const V3 &n_) : ref(ref_), p(p_), n(n_) { wi = normalized(p - ref); }
};
/* light.cpp:16-19 */
inline V3 lightEval(const kz_light_desc &l, const LightQueryRecord &lRec) {
    float cosTheta = dot(lRec.n, -lRec.wi);
    return cosTheta > 0.f ? V3(l.radiance[0], l.radiance[1], l.radiance[2]) : V3(0.f);
}
/* light.cpp:36-51 */
inline float lightPdf(const MeshData &mesh, const LightQueryRecord &lRec) {
    float pdf = mesh.normalization;
    float cosTheta = dot(lRec.n, -lRec.wi);
    if (cosTheta > 0.f) {
        float distance2 = sqnorm(lRec.p - lRec.ref);
        return pdf * distance2 / cosTheta;
    }
    return 0.f;
}
/* light.cpp:21-34 */
inline V3 lightSample(const kz_light_desc &l, const MeshData &mesh, LightQueryRecord &lRec, Sampler &sampler) {
    meshSample(mesh, sampler, lRec.p, lRec.n);
    lRec.wi = normalized(lRec.p - lRec.ref);
    float dist = norm(lRec.p - lRec.ref);
    lRec.shadowRay = kz_ray{{lRec.ref.x, lRec.ref.y, lRec.ref.z}, 0.f, {lRec.wi.x, lRec.wi.y, lRec.wi.z}, dist};
    lRec.pdf = lightPdf(mesh, lRec);
    if (lRec.pdf > 0.f && !std::isnan(lRec.pdf) && !std::isinf(lRec.pdf)) return lightEval(l, lRec) / lRec.pdf;
    return V3(0.f);
}

/* ---- post-intersection, accel.cpp:113-236 -------------------------------------------------- */
inline void fillIntersection(const SceneData &sc, const HitRec &h, Intersection &its) {
    its.t = h.t;
    its.uv = V2{h.u, h.v};
    its.mesh = (int)h.geom;
    uint32_t f = h.prim;
    float b0 = 1 - (its.uv.x + its.uv.y), b1 = its.uv.x, b2 = its.uv.y;    /* bary << 1-uv.sum(), uv */
    const MeshData &m = sc.meshes[its.mesh];
    bool hasN = !m.N.empty(), hasUV = !m.UV.empty();
    uint32_t i0 = m.F[3 * f], i1 = m.F[3 * f + 1], i2 = m.F[3 * f + 2];
    V3 p0 = m.pos(i0), p1 = m.pos(i1), p2 = m.pos(i2);
    /* The reference reads N.col(idx) unconditionally (accel.cpp:135); with no normals the mesh
     * cannot be smooth-shaded, so the Hanika offset is defined here as the identity (n_i = 0). */
    V3 n0 = hasN ? m.nrm(i0) : V3(0.f), n1 = hasN ? m.nrm(i1) : V3(0.f), n2 = hasN ? m.nrm(i2) : V3(0.f);
    V3 orignP = b0 * p0 + b1 * p1 + b2 * p2;
    V3 tmpu = orignP - p0, tmpv = orignP - p1, tmpw = orignP - p2;
    float dotu = std::min(0.f, dot(tmpu, n0));
    float dotv = std::min(0.f, dot(tmpv, n1));
    float dotw = std::min(0.f, dot(tmpw, n2));
    tmpu -= dotu * n0; tmpv -= dotv * n1; tmpw -= dotw * n2;
    its.p = orignP + b0 * tmpu + b1 * tmpv + b2 * tmpw;
    V3 dp0 = p1 - p0, dp1 = p2 - p0;
    its.geoFrame = Frame(normalized(cross(dp0, dp1)));
    if (hasUV) {
        V2 a = m.uv(i0), b = m.uv(i1), c = m.uv(i2);
        its.uv = V2{b0 * a.x + b1 * b.x + b2 * c.x, b0 * a.y + b1 * b.y + b2 * c.y};
    }
    if (hasN && hasUV) {
        V2 uv0 = m.uv(i0), uv1 = m.uv(i1), uv2 = m.uv(i2);
        V2 duv0{uv1.x - uv0.x, uv1.y - uv0.y}, duv1{uv2.x - uv0.x, uv2.y - uv0.y};
        V3 shNormal = b0 * n0 + b1 * n1 + b2 * n2;
        float length = norm(cross(dp0, dp1));
        if (length > 0.f) {
            float determinant = duv0.x * duv1.y - duv0.y * duv1.x;
            if (determinant > 0.f) {
                float invDet = 1.0f / determinant;
                its.dpdu = (duv1.y * dp0 - duv0.y * dp1) * invDet;
                its.dpdv = (-duv1.x * dp0 + duv0.x * dp1) * invDet;
                its.shFrame.n = normalized(shNormal);
                its.shFrame.s = normalized(its.dpdu - shNormal * dot(shNormal, its.dpdu));
                its.shFrame.t = normalized(cross(its.shFrame.n, its.shFrame.s));
            } else {
                its.shFrame = Frame(normalized(shNormal));
                its.dpdu = its.shFrame.s;
                its.dpdv = its.shFrame.t;
            }
        } else {
            its.shFrame = Frame(normalized(shNormal));
        }
    } else if (hasN) {
        its.shFrame = Frame(normalized(b0 * n0 + b1 * n1 + b2 * n2));
    } else {
        its.shFrame = its.geoFrame;
    }
}

}  // namespace kzo
#endif
