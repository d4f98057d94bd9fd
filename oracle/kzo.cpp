/* ORACLE (test infrastructure) -- C entry points + path_mis integrator + film + camera.
 * Restates: integrator.cpp:185-355 (PathMisIntegrator), renderer.cpp:20-69 (renderSample /
 * renderBlock), camera.cpp:35-91,156-223, block.cpp:9-85, rfilter.cpp:10-102, scene.h:45-56,
 * bitmap.cpp:46-54.  See kzo.h for what is and is not pinned. */
#include "kzo.h"
#include "kzo_shading.h"
#include <atomic>
#include <cstdio>
#include <string>
#include <thread>

using namespace kzo;

/* `bsdf->sample(bRec, sampler->next1D(), sampler->next2D())` (integrator.cpp:122,168,307): the two draws are indeterminately
 * sequenced call arguments; GCC -- the toolchain the reference is built with on Linux and the only one available here -- evaluates
 * them RIGHT TO LEFT, so next2D() takes the lower dimensions and next1D() the one after them.  KZO_ARG_ORDER_LTR=1 restores the
 * textual order (what clang would do). */
#ifndef KZO_ARG_ORDER_LTR
#define KZO_ARG_ORDER_LTR 0
#endif
#if KZO_ARG_ORDER_LTR
#define KZO_DRAW_BSDF_SAMPLES(sampler, s1, s2) float s1 = (sampler).next1D(); V2 s2 = (sampler).next2D()
#else
#define KZO_DRAW_BSDF_SAMPLES(sampler, s1, s2) V2 s2 = (sampler).next2D(); float s1 = (sampler).next1D()
#endif


struct kzo_scene {
    SceneData sc;
    std::atomic<uint64_t> paths{0}, raysExt{0}, raysShadow{0}, vertices{0};
    int border = 0;
};

static thread_local std::string g_err;
static int fail(int code, const std::string &msg) { g_err = msg; return code; }
extern "C" const char *kzo_last_error(void) { return g_err.c_str(); }

/* ------------------------------------------------------------------ scene */
extern "C" int kzo_scene_create(const kz_scene_desc *d, kzo_scene **out) {
    if (!d || !out) return fail(KZ_ERR_INVALID, "null argument");
    kzo_scene *s = new kzo_scene();
    SceneData &sc = s->sc;
    sc.bsdfs.assign(d->bsdfs, d->bsdfs + d->n_bsdfs);
    sc.textures.assign(d->textures, d->textures + d->n_textures);
    sc.lights.assign(d->lights, d->lights + d->n_lights);
    for (uint32_t i = 0; i < d->n_images; ++i) {
        Image im; im.w = d->images[i].width; im.h = d->images[i].height;
        im.rgb.assign(d->images[i].rgb, d->images[i].rgb + (size_t)3 * im.w * im.h);
        sc.images.push_back(std::move(im));
    }
    sc.background = d->background;
    sc.camera = d->camera;
    sc.integrator = d->integrator;
    sc.filter = d->filter;
    sc.sampler.d = d->sampler;
    if (d->sampler.type == KZ_SAMPLER_PMJ02BN) {
        if (!d->sampler.blue_noise || !d->sampler.pmj02bn) { delete s; return fail(KZ_ERR_INVALID, "pmj02bn needs tables"); }
        sc.blueNoise.assign(d->sampler.blue_noise, d->sampler.blue_noise + 48 * 128 * 128);
        sc.pmj.assign(d->sampler.pmj02bn, d->sampler.pmj02bn + 5 * 65536 * 2);
        sc.sampler.d.blue_noise = sc.blueNoise.data();
        sc.sampler.d.pmj02bn = sc.pmj.data();
        buildPmjPixelSamples(sc.sampler);
    }
    for (uint32_t g = 0; g < d->n_meshes; ++g) {
        const kz_mesh_desc &md = d->meshes[g];
        MeshData m;
        m.nV = md.n_vertices; m.nF = md.n_triangles; m.bsdf = md.bsdf; m.light = md.light;
        m.P.assign(md.positions, md.positions + (size_t)3 * m.nV);
        if (md.normals) m.N.assign(md.normals, md.normals + (size_t)3 * m.nV);
        if (md.uvs) m.UV.assign(md.uvs, md.uvs + (size_t)2 * m.nV);
        m.F.assign(md.indices, md.indices + (size_t)3 * m.nF);
        if (m.light >= 0) { buildLightCdf(m); sc.lightMeshes.push_back((int)g); }
        for (uint32_t f = 0; f < m.nF; ++f) {
            Tri t; t.p0 = m.pos(m.F[3 * f]); t.p1 = m.pos(m.F[3 * f + 1]); t.p2 = m.pos(m.F[3 * f + 2]);
            t.geom = g; t.prim = f;
            sc.accel.tris.push_back(t);
        }
        sc.meshes.push_back(std::move(m));
    }
    sc.accel.build();
    s->border = (int)std::ceil(sc.filter.radius - 0.5f);     /* block.cpp:14 */
    *out = s;
    return KZ_OK;
}
extern "C" void kzo_scene_destroy(kzo_scene *s) { delete s; }

template <typename F> static void parallelFor(size_t n, int threads, F f) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (threads == 1 || n < 64) { f(0, n); return; }
    std::atomic<size_t> next{0};
    size_t chunk = std::max<size_t>(64, n / ((size_t)threads * 64));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() {
            for (;;) {
                size_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min(n, b + chunk));
            }
        });
    for (auto &th : pool) th.join();
}

/* ------------------------------------------------------------------ trace */
static inline kz_hit toHit(const HitRec &h) {
    kz_hit o; o.t = h.t; o.u = h.u; o.v = h.v; o.prim_id = h.prim; o.geom_id = h.geom; return o;
}
extern "C" int kzo_trace(kzo_scene *s, const kz_ray *rays, size_t n, int shadow, int brute, int threads, kz_hit *hits) {
    (void)shadow;   /* accel.cpp:98-104: the shadow variant is the same closest-hit query */
    if (!s || (!rays && n) || (!hits && n)) return fail(KZ_ERR_INVALID, "null argument");
    parallelFor(n, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i)
            hits[i] = toHit(brute ? s->sc.accel.traceBrute(rays[i]) : s->sc.accel.traceBvh(rays[i]));
    });
    return KZ_OK;
}

/* integrator.cpp:259-278: closest-hit with stepping through invisible lights */
static bool occludedWalk(kzo_scene *s, kz_ray tempRay, float eps, int *segments) {
    const SceneData &sc = s->sc;
    int seg = 0;
    bool occluded = false;
    while (true) {
        ++seg;
        HitRec h = sc.accel.traceBvh(tempRay);
        if (h.geom != KZ_INVALID_ID) {
            const MeshData &m = sc.meshes[h.geom];
            if (m.light < 0) { occluded = true; break; }
            if (sc.lights[m.light].primary_visibility) { occluded = true; break; }
            V3 o(tempRay.o[0], tempRay.o[1], tempRay.o[2]), d(tempRay.d[0], tempRay.d[1], tempRay.d[2]);
            V3 no = o + d * (h.t + eps);
            tempRay = kz_ray{{no.x, no.y, no.z}, eps, {d.x, d.y, d.z}, tempRay.tmax - h.t};
        } else break;
        if (seg > 4096) break;   /* safety: the reference would loop forever on t == 0 chains */
    }
    if (segments) *segments = seg;
    return occluded;
}
extern "C" int kzo_occluded(kzo_scene *s, const kz_ray *rays, size_t n, float trace_bias, int threads,
                            uint8_t *occluded, uint8_t *segments) {
    if (!s || (!rays && n) || (!occluded && n)) return fail(KZ_ERR_INVALID, "null argument");
    parallelFor(n, threads, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            int seg = 0;
            occluded[i] = occludedWalk(s, rays[i], trace_bias, &seg) ? 1 : 0;
            if (segments) segments[i] = (uint8_t)std::min(seg, 255);
        }
    });
    return KZ_OK;
}

/* ---------------------------------------------------------------- sampler */
extern "C" int kzo_sample_dump(kzo_scene *s, const int32_t *triples, size_t n, const char *pattern, float *out) {
    if (!s || !pattern) return fail(KZ_ERR_INVALID, "null argument");
    size_t per = 0;
    for (const char *p = pattern; *p; ++p) per += (*p == '1') ? 1 : 2;
    for (size_t i = 0; i < n; ++i) {
        Sampler sm; sm.cfg = &s->sc.sampler;
        sm.generateSample(triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]);
        float *o = out + i * per;
        for (const char *p = pattern; *p; ++p) {
            if (*p == '1') *o++ = sm.next1D();
            else { V2 v = (*p == 'P') ? sm.nextPixel2D() : sm.next2D(); *o++ = v.x; *o++ = v.y; }
        }
    }
    return KZ_OK;
}

/* ----------------------------------------------------------------- camera */
/* camera.cpp:70-91 (perspective), :191-223 (thinlens) */
static kz_ray cameraRay(const kz_camera_desc &c, V2 samplePosition, V2 apertureSample) {
    M44 s2c, c2w;
    std::memcpy(s2c.m, c.sample_to_camera, sizeof(s2c.m));
    std::memcpy(c2w.m, c.camera_to_world, sizeof(c2w.m));
    float invW = 1.0f / (float)c.width, invH = 1.0f / (float)c.height;   /* cwiseInverse, camera.cpp:19 */
    V3 nearP = xformPoint(s2c, V3(samplePosition.x * invW, samplePosition.y * invH, 0.0f));
    V3 o, d;
    if (c.type == KZ_CAM_THINLENS) {
        V2 tmp = squareToUniformDisk(apertureSample);
        tmp.x *= c.aperture_radius; tmp.y *= c.aperture_radius;
        V3 apertureP(tmp.x, tmp.y, 0.0f);
        V3 focusP = nearP * (c.focus_distance / nearP.z);
        d = normalized(focusP - apertureP);
        o = xformPoint(c2w, apertureP);
    } else {
        d = normalized(nearP);
        o = xformPoint(c2w, V3(0, 0, 0));
    }
    float invZ = 1.0f / d.z;
    V3 dw = xformVector(c2w, d);
    return kz_ray{{o.x, o.y, o.z}, c.near_clip * invZ, {dw.x, dw.y, dw.z}, c.far_clip * invZ};
}
extern "C" int kzo_camera_rays(kzo_scene *s, const float *samples4, size_t n, kz_ray *out) {
    if (!s) return fail(KZ_ERR_INVALID, "null argument");
    for (size_t i = 0; i < n; ++i)
        out[i] = cameraRay(s->sc.camera, V2{samples4[4 * i], samples4[4 * i + 1]}, V2{samples4[4 * i + 2], samples4[4 * i + 3]});
    return KZ_OK;
}

/* camera.cpp:35-62: perspective matrix, scale/translate, 4x4 inverse (Eigen float inverse
 * restated as a double-precision Gauss-Jordan rounded to float; differences are <= few ulp and
 * only move camera rays sub-pixel -- ray-level parity uses shared ray batches). */
extern "C" void kzo_camera_matrix(int width, int height, float fov_deg, float near_clip, float far_clip, float out16[16]) {
    float aspect = width / (float)height;
    float recip = 1.0f / (far_clip - near_clip);
    float cot = 1.0f / std::tan((fov_deg / 2.0f) * (kPi / 180.0f));
    double P[16] = {cot, 0, 0, 0, 0, cot, 0, 0, 0, 0, far_clip * recip, -near_clip * far_clip * recip, 0, 0, 1, 0};
    double T[16] = {1, 0, 0, -1.0f, 0, 1, 0, -1.0f / aspect, 0, 0, 1, 0, 0, 0, 0, 1};
    double S[16] = {-0.5f, 0, 0, 0, 0, -0.5f * aspect, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    auto mul = [](const double *A, const double *B, double *C) {
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
            float acc = 0.f;                                  /* Eigen float products */
            for (int k = 0; k < 4; ++k) acc += (float)A[i * 4 + k] * (float)B[k * 4 + j];
            C[i * 4 + j] = acc;
        }
    };
    double TP[16], M[16];
    mul(T, P, TP); mul(S, TP, M);
    double a[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { a[i][j] = M[i * 4 + j]; a[i][j + 4] = (i == j); }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
        for (int j = 0; j < 8; ++j) std::swap(a[c][j], a[piv][j]);
        double inv = 1.0 / a[c][c];
        for (int j = 0; j < 8; ++j) a[c][j] *= inv;
        for (int r = 0; r < 4; ++r) if (r != c) { double f = a[r][c]; for (int j = 0; j < 8; ++j) a[r][j] -= f * a[c][j]; }
    }
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) out16[i * 4 + j] = (float)a[i][j + 4];
}

/* ----------------------------------------------------------------- filter */
extern "C" void kzo_filter_table(int kind, float p0, float p1, float p2, float *radius_out, float table33[33]) {
    float radius = 0.f;
    auto eval = [&](float x) -> float {
        switch (kind) {
            case 0: {   /* gaussian(radius=p0, stddev=p1) rfilter.cpp:16-27 */
                float alpha = -1.0f / (2.0f * p1 * p1);
                return std::max(0.0f, std::exp(alpha * x * x) - std::exp(alpha * p0 * p0));
            }
            case 1: {   /* mitchell(radius=p0, B=p1, C=p2) rfilter.cpp:44-66 */
                float B = p1, C = p2;
                x = std::fabs(2.0f * x / p0);
                float x2 = x * x, x3 = x2 * x;
                if (x < 1) return 1.0f / 6.0f * ((12 - 9 * B - 6 * C) * x3 + (-18 + 12 * B + 6 * C) * x2 + (6 - 2 * B));
                else if (x < 2) return 1.0f / 6.0f * ((-B - 6 * C) * x3 + (6 * B + 30 * C) * x2 + (-12 * B - 48 * C) * x + (8 * B + 24 * C));
                return 0.0f;
            }
            case 2: return std::max(0.0f, 1.0f - std::fabs(x));   /* tent rfilter.cpp:77-83 */
            default: return 1.0f;                                  /* box rfilter.cpp:91-97 */
        }
    };
    radius = kind == 2 ? 1.0f : (kind == 3 ? 0.5f : p0);
    for (int i = 0; i < 32; ++i) table33[i] = eval((radius * i) / 32);   /* block.cpp:16-19 */
    table33[32] = 0.0f;
    *radius_out = radius;
}

extern "C" int kzo_light_cdf(kzo_scene *s, int mesh, float *cdf_out, float *normalization) {
    if (!s || mesh < 0 || mesh >= (int)s->sc.meshes.size()) return fail(KZ_ERR_INVALID, "bad mesh");
    const MeshData &m = s->sc.meshes[mesh];
    if (m.cdf.empty()) return fail(KZ_ERR_INVALID, "mesh is not a light");
    std::memcpy(cdf_out, m.cdf.data(), m.cdf.size() * sizeof(float));
    *normalization = m.normalization;
    return KZ_OK;
}

/* ----------------------------------------------------- field-by-field probes */
/* Accel::rayIntersect, accel.cpp:63-236: closest hit + the Intersection it fills (layout: include/kzgpu.h, kzgpu_intersection_dump) */
extern "C" int kzo_intersection_dump(kzo_scene *s, const kz_ray *rays, size_t n, float *out24) {
    if (!s || (!rays && n) || (!out24 && n)) return fail(KZ_ERR_INVALID, "null argument");
    parallelFor(n, 0, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; ++i) {
            float *o = out24 + 24 * i;
            for (int k = 0; k < 24; ++k) o[k] = 0.f;
            HitRec h = s->sc.accel.traceBvh(rays[i]);
            o[0] = h.t; o[1] = -1.f;
            if (h.geom == KZ_INVALID_ID) continue;
            Intersection its;
            fillIntersection(s->sc, h, its);
            o[1] = (float)its.mesh;
            o[2] = its.p.x; o[3] = its.p.y; o[4] = its.p.z; o[5] = its.uv.x; o[6] = its.uv.y;
            o[7] = its.geoFrame.n.x; o[8] = its.geoFrame.n.y; o[9] = its.geoFrame.n.z;
            o[10] = its.shFrame.s.x; o[11] = its.shFrame.s.y; o[12] = its.shFrame.s.z;
            o[13] = its.shFrame.t.x; o[14] = its.shFrame.t.y; o[15] = its.shFrame.t.z;
            o[16] = its.shFrame.n.x; o[17] = its.shFrame.n.y; o[18] = its.shFrame.n.z;
            o[19] = its.dpdu.x; o[20] = its.dpdu.y; o[21] = its.dpdu.z;
        }
    });
    return KZ_OK;
}

/* Scene::getRandomLight (scene.h:45-56) + AreaLight::sample (light.cpp:21-34) fed with given random numbers
 * (layout: include/kzgpu.h, kzgpu_light_sample_dump) */
extern "C" int kzo_light_sample_dump(kzo_scene *s, const float *ref3, const float *u5, size_t n, float *out16) {
    if (!s || (n && (!ref3 || !u5 || !out16))) return fail(KZ_ERR_INVALID, "null argument");
    const SceneData &sc = s->sc;
    for (size_t i = 0; i < n; ++i) {
        float *o = out16 + 16 * i;
        for (int k = 0; k < 16; ++k) o[k] = 0.f;
        o[0] = -1.f;
        size_t nl = sc.lightMeshes.size();
        if (nl == 0) continue;
        Sampler sm; sm.cfg = &sc.sampler; sm.fixed = u5 + 5 * i;
        float rnd = sm.next1D();
        size_t index = std::min((size_t)std::floor(nl * rnd), nl - 1);
        const MeshData &lm = sc.meshes[sc.lightMeshes[index]];
        LightQueryRecord lRec(V3(ref3[3 * i], ref3[3 * i + 1], ref3[3 * i + 2]));
        V3 Ls = lightSample(sc.lights[lm.light], lm, lRec, sm);
        o[0] = (float)sc.lightMeshes[index];
        o[1] = lRec.p.x; o[2] = lRec.p.y; o[3] = lRec.p.z; o[4] = lRec.n.x; o[5] = lRec.n.y; o[6] = lRec.n.z;
        o[7] = lRec.wi.x; o[8] = lRec.wi.y; o[9] = lRec.wi.z; o[10] = lRec.shadowRay.tmax; o[11] = lRec.pdf;
        o[12] = Ls.x; o[13] = Ls.y; o[14] = Ls.z;
    }
    return KZ_OK;
}

/* -------------------------------------------------------------- integrator */
static inline float powerHeuristic(float a, float b) {   /* integrator.cpp:340-344 */
    a *= a; b *= b;
    return a > 0.f ? a / (a + b) : 0.f;
}

struct PathCounters { uint64_t ext = 0, shadow = 0, vertices = 0; };

static bool rayIntersect(const SceneData &sc, const kz_ray &ray, Intersection &its, PathCounters &pc) {
    ++pc.ext;
    HitRec h = sc.accel.traceBvh(ray);
    if (h.geom == KZ_INVALID_ID) return false;
    fillIntersection(sc, h, its);
    return true;
}

/* integrator.cpp:195-338 */
static V3 Li(kzo_scene *s, Sampler &sampler, const kz_ray &ray_, PathCounters &pc) {
    const SceneData &sc = s->sc;
    const kz_integrator_desc &I = sc.integrator;
    const float eps = I.trace_bias;
    kz_ray ray = ray_;
    V3 L(0.f), throughput(1.f);
    float eta = 1.f;
    float bsdfWeight = 1.f;
    Intersection its;
    if (!rayIntersect(sc, ray, its, pc)) return L;
    {
        const MeshData &m = sc.meshes[its.mesh];
        if (m.light >= 0 && !sc.lights[m.light].primary_visibility) {
            V3 d(ray.d[0], ray.d[1], ray.d[2]);
            V3 o = its.p + eps * d;
            kz_ray newRay{{o.x, o.y, o.z}, kEpsilon, {d.x, d.y, d.z}, INFINITY};   /* Ray3f(o,d): ray.h:35-38 */
            rayIntersect(sc, newRay, its, pc);
        }
    }
    int depth = 0;
    while (depth < I.max_depth) {
        const MeshData &mesh = sc.meshes[its.mesh];
        V3 rayO(ray.o[0], ray.o[1], ray.o[2]), rayD(ray.d[0], ray.d[1], ray.d[2]);
        if (mesh.light >= 0) {
            LightQueryRecord lRec(rayO, its.p, its.shFrame.n);
            L += bsdfWeight * throughput * lightEval(sc.lights[mesh.light], lRec);
            break;
        }
        if (depth >= 3) {
            float probability = std::min(maxcoeff(throughput) * eta * eta, 0.95f);
            if (probability <= sampler.next1D()) break;
            throughput /= probability;
        }
        ++pc.vertices;
        /* light sampling; scene.h:45-56 */
        float rnd = sampler.next1D();
        size_t nl = sc.lightMeshes.size();
        if (nl > 0) {
            size_t index = std::min((size_t)std::floor(nl * rnd), nl - 1);
            const MeshData &lm = sc.meshes[sc.lightMeshes[index]];
            const kz_light_desc &light = sc.lights[lm.light];
            LightQueryRecord lRec(its.p);
            lRec.uv = its.uv;
            V3 Ls = lightSample(light, lm, lRec, sampler) / (1.f / nl);
            float lightPdfV = lightPdf(lm, lRec);
            lRec.shadowRay.tmin = eps;
            lRec.shadowRay.tmax -= eps;
            int seg = 0;
            bool occluded = occludedWalk(s, lRec.shadowRay, eps, &seg);
            pc.shadow += (uint64_t)seg;
            if (!occluded) {
                BSDFQueryRecord bRec(its.toLocal(-rayD), its.toLocal(lRec.wi), ESolidAngle);
                bRec.its = its;
                bRec.uv = its.uv;
                V3 f = bsdfEval(sc, mesh.bsdf, bRec);
                float bsdfPdfV = bsdfPdf(sc, mesh.bsdf, bRec);
                float lightWeight = powerHeuristic(lightPdfV, bsdfPdfV);
                L += throughput * Ls * f * lightWeight;
            }
        }
        if (I.regularization)
            its.accumulatedRoughness += bsdfRegularize(sc, mesh.bsdf, its.uv) * I.accumulated_roughness;
        /* BSDF sampling */
        BSDFQueryRecord bRec(its.shFrame.toLocal(-rayD));
        bRec.uv = its.uv;
        bRec.its = its;
        KZO_DRAW_BSDF_SAMPLES(sampler, s1, s2);
        V3 bsdfColor = bsdfSample(sc, mesh.bsdf, bRec, s1, s2);
        throughput *= bsdfColor;
        eta *= bRec.eta;
        V3 wo = its.toWorld(bRec.wo);
        ray = kz_ray{{its.p.x, its.p.y, its.p.z}, eps, {wo.x, wo.y, wo.z}, INFINITY};
        float bsdfPdfV = bsdfPdf(sc, mesh.bsdf, bRec);
        int prevBsdfMeasure = bRec.measure;
        /* A zero weight leaves bRec.wo unset in the reference (uninitialised direction) while every
         * later contribution is multiplied by throughput == 0: the path is dead.  Stop here; the
         * CUDA path does the same (DESIGN.md "Equivalences"). */
        if (iszero(throughput)) break;
        if (!rayIntersect(sc, ray, its, pc)) {
            L += throughput * backgroundColor(sc, wo);
            break;
        }
        const MeshData &nm = sc.meshes[its.mesh];
        if (nm.light >= 0) {
            LightQueryRecord lRec_(V3(ray.o[0], ray.o[1], ray.o[2]), its.p, its.shFrame.n);
            float lightPdf_ = lightPdf(nm, lRec_);
            bsdfWeight = powerHeuristic(bsdfPdfV, lightPdf_);
        }
        if (prevBsdfMeasure == EDiscrete) bsdfWeight = 1.f;
        depth++;
    }
    return L;
}

/* ---- normals / ao / whitted / path_mats (integrator.cpp:11-181) --------------------------------- */
static V3 squareToUniformHemisphere(V2 sample) {          /* warp.cpp:68-79 */
    float z = sample.x;
    float tmp = std::sqrt(1.0f - z * z);
    float phi = 2.0f * kPi * sample.y;
    return V3(std::cos(phi) * tmp, std::sin(phi) * tmp, z);
}
static bool anyHit(const SceneData &sc, const kz_ray &ray, PathCounters &pc) {       /* Scene::rayIntersect(ray): closest-hit query, boolean */
    ++pc.shadow;
    return sc.accel.traceBvh(ray).geom != KZ_INVALID_ID;
}
static V3 LiAlt(kzo_scene *s, Sampler &sampler, const kz_ray &ray_, PathCounters &pc) {
    const SceneData &sc = s->sc;
    const int type = sc.integrator.type;
    kz_ray ray = ray_;
    Intersection its;
    if (type == KZ_INTEGRATOR_NORMALS) {                                               /* integrator.cpp:19-29 */
        if (!rayIntersect(sc, ray, its, pc)) return V3(0.f);
        return V3(std::fabs(its.geoFrame.n.x), std::fabs(its.geoFrame.n.y), std::fabs(its.geoFrame.n.z));
    }
    if (type == KZ_INTEGRATOR_AO) {                                                    /* integrator.cpp:43-62 */
        if (!rayIntersect(sc, ray, its, pc)) return V3(0.f);
        V3 sample = squareToUniformHemisphere(sampler.next2D());
        V3 point = its.toWorld(sample);
        kz_ray shadowRay{{its.p.x, its.p.y, its.p.z}, kEpsilon, {point.x, point.y, point.z}, INFINITY};
        if (!anyHit(sc, shadowRay, pc)) {
            V3 n = normalized(its.shFrame.n);
            point = normalized(point);
            float cosTheta = dot(point, n);          /* shFrame.cosTheta(toLocal(point)) with the re-normalised n */
            return V3(cosTheta / kPi) / (0.5f * kInvPi);
        }
        return V3(0.f);
    }
    if (type == KZ_INTEGRATOR_WHITTED) {                                               /* integrator.cpp:80-128, recursion unrolled */
        /* The reference recurses: Li = reflect * Li(next) / 0.95, so the factors are applied from the LEAF upwards, one multiply and
         * one divide per level.  `chain` keeps the per-level factors and `unwind` folds them in that order (a running product from the
         * top would round differently); `weight` only serves the dead-path test. */
        V3 weight(1.f);
        std::vector<V3> chain;
        auto unwind = [&](V3 leaf) { for (size_t k = chain.size(); k-- > 0;) leaf = chain[k] * leaf / 0.95f; return leaf; };
        for (int depth = 0; depth < 4096; ++depth) {
            if (!rayIntersect(sc, ray, its, pc)) return unwind(V3(0.f));
            const MeshData &mesh = sc.meshes[its.mesh];
            V3 rayO(ray.o[0], ray.o[1], ray.o[2]), rayD(ray.d[0], ray.d[1], ray.d[2]);
            V3 Le(0.f);
            if (mesh.light >= 0) { LightQueryRecord leRec(rayO, its.p, its.shFrame.n); Le = lightEval(sc.lights[mesh.light], leRec); }
            if (sc.bsdfs[mesh.bsdf].type == KZ_BSDF_DIFFUSE) {                         /* the only BSDF with isDiffuse() == true, bsdf.cpp:77 */
                float rnd = sampler.next1D();
                size_t nl = sc.lightMeshes.size();
                if (nl == 0) return unwind(Le);
                size_t index = std::min((size_t)std::floor(nl * rnd), nl - 1);
                const MeshData &lm = sc.meshes[sc.lightMeshes[index]];
                LightQueryRecord rec(its.p);
                V3 Ls = lightSample(sc.lights[lm.light], lm, rec, sampler);
                if (anyHit(sc, rec.shadowRay, pc)) Ls = V3(0.f);
                float cosTheta = its.shFrame.toLocal(rec.wi).z;
                if (cosTheta < 0.f) cosTheta = 0.f;
                BSDFQueryRecord bRec(its.toLocal(-rayD), its.toLocal(rec.wi), ESolidAngle);
                V3 f = bsdfEval(sc, mesh.bsdf, bRec);
                V3 Lr = f * Ls * cosTheta;
                return unwind(Le + Lr / (1.0f / nl));
            }
            BSDFQueryRecord bRec(its.toLocal(-rayD));
            bRec.its = its; bRec.uv = its.uv;          /* the reference leaves bRec.its/uv default here; every BSDF that reads them would see zeros */
            KZO_DRAW_BSDF_SAMPLES(sampler, s1, s2);
            V3 refl = bsdfSample(sc, mesh.bsdf, bRec, s1, s2);
            if (!((double)sampler.next1D() < 0.95)) return unwind(V3(0.f));            /* `next1D() < 0.95`: a double comparison upstream */
            chain.push_back(refl);
            weight = weight * refl / 0.95f;
            if (iszero(weight)) return V3(0.f);
            V3 wo = its.toWorld(bRec.wo);
            ray = kz_ray{{its.p.x, its.p.y, its.p.z}, kEpsilon, {wo.x, wo.y, wo.z}, INFINITY};
        }
        return V3(0.f);
    }
    /* path_mats, integrator.cpp:142-173 */
    V3 color(0.f), t(1.f);
    for (int depth = 0; depth < 4096; ++depth) {
        if (!rayIntersect(sc, ray, its, pc)) return color;
        const MeshData &mesh = sc.meshes[its.mesh];
        V3 rayO(ray.o[0], ray.o[1], ray.o[2]), rayD(ray.d[0], ray.d[1], ray.d[2]);
        if (mesh.light >= 0) { LightQueryRecord lRecE(rayO, its.p, its.shFrame.n); color += t * lightEval(sc.lights[mesh.light], lRecE); }
        float probability = std::min(t.x, 0.95f);
        if (sampler.next1D() >= probability) return color;
        t /= probability;
        ++pc.vertices;
        BSDFQueryRecord bRec(its.shFrame.toLocal(-rayD));
        bRec.uv = its.uv; bRec.its = its;
        bRec.its.accumulatedRoughness = 0.f;
        KZO_DRAW_BSDF_SAMPLES(sampler, s1, s2);
        V3 f = bsdfSample(sc, mesh.bsdf, bRec, s1, s2);
        t *= f;
        if (iszero(t)) return color;               /* dead path, as in Li above */
        V3 wo = its.toWorld(bRec.wo);
        ray = kz_ray{{its.p.x, its.p.y, its.p.z}, kEpsilon, {wo.x, wo.y, wo.z}, INFINITY};
    }
    return color;
}

/* ------------------------------------------------------------------- film */
/* block.cpp:56-85 on the full bordered frame (offset 0): the per-tile blocks of the reference
 * never clip a splat (border = ceil(r-0.5)), so tile-local and whole-frame puts are identical
 * up to summation order. */
static void filmPut(const SceneData &sc, int border, float *frame, V2 pos_, V3 value) {
    if (!colorValid(value)) return;
    const kz_filter_desc &F = sc.filter;
    int cols = sc.camera.width + 2 * border, rows = sc.camera.height + 2 * border;
    float px = pos_.x - 0.5f - (0 - border), py = pos_.y - 0.5f - (0 - border);
    float lookup = 32 / F.radius;
    int x0 = std::max(0, (int)std::ceil(px - F.radius)), y0 = std::max(0, (int)std::ceil(py - F.radius));
    int x1 = std::min(cols - 1, (int)std::floor(px + F.radius)), y1 = std::min(rows - 1, (int)std::floor(py + F.radius));
    float wx[16], wy[16];
    for (int x = x0, i = 0; x <= x1 && i < 16; ++x) wx[i++] = F.table[(int)(std::fabs(x - px) * lookup)];
    for (int y = y0, i = 0; y <= y1 && i < 16; ++y) wy[i++] = F.table[(int)(std::fabs(y - py) * lookup)];
    for (int y = y0, yr = 0; y <= y1; ++y, ++yr)
        for (int x = x0, xr = 0; x <= x1; ++x, ++xr) {
            float *p = frame + 4 * ((size_t)y * cols + x);
            float w = wx[xr] * wy[yr];      /* Color4f(value) * wX * wY: ((c*wx)*wy) */
            p[0] += value.x * wx[xr] * wy[yr];
            p[1] += value.y * wx[xr] * wy[yr];
            p[2] += value.z * wx[xr] * wy[yr];
            p[3] += 1.0f * wx[xr] * wy[yr];
            (void)w;
        }
}

extern "C" int kzo_frame_dims(const kzo_scene *s, int32_t *w, int32_t *h, int32_t *b) {
    if (!s) return fail(KZ_ERR_INVALID, "null argument");
    *w = s->sc.camera.width; *h = s->sc.camera.height; *b = s->border;
    return KZ_OK;
}

/* renderer.cpp:20-69.  Threads own disjoint row bands and private frames that are summed at
 * the end (the reference merges per-tile blocks under a mutex, block.cpp:87-96). */
extern "C" int kzo_render(kzo_scene *s, const kz_render_req *req, int threads, float *frame) {
    if (!s || !req || !frame) return fail(KZ_ERR_INVALID, "null argument");
    const SceneData &sc = s->sc;
    int b = s->border, cols = sc.camera.width + 2 * b, rows = sc.camera.height + 2 * b;
    size_t fsz = (size_t)4 * cols * rows;
    if (req->clear_frame) std::fill(frame, frame + fsz, 0.f);
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    int nrows = req->y1 - req->y0;
    std::atomic<int> nextRow{0};
    std::vector<std::vector<float>> priv((size_t)threads);
    std::vector<std::thread> pool;
    std::vector<PathCounters> counters((size_t)threads);
    std::vector<uint64_t> npaths((size_t)threads, 0);
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            std::vector<float> &f = priv[t];
            f.assign(fsz, 0.f);
            Sampler sampler; sampler.cfg = &sc.sampler;
            for (;;) {
                int r = nextRow.fetch_add(1);
                if (r >= nrows) break;
                int y = req->y0 + r;
                for (int x = req->x0; x < req->x1; ++x)
                    for (int j = req->spp_begin; j < req->spp_end; ++j) {
                        sampler.generateSample(x, y, j);
                        V2 pix = sampler.nextPixel2D();
                        V2 pixelSample{float(x) + pix.x, float(y) + pix.y};
                        V2 apertureSample = sampler.next2D();
                        kz_ray ray = cameraRay(sc.camera, pixelSample, apertureSample);
                        V3 value = V3(1.0f) * (sc.integrator.type == KZ_INTEGRATOR_PATH_MIS ? Li(s, sampler, ray, counters[t]) : LiAlt(s, sampler, ray, counters[t]));
                        filmPut(sc, b, f.data(), pixelSample, value);
                        ++npaths[t];
                    }
            }
        });
    for (auto &th : pool) th.join();
    for (int t = 0; t < threads; ++t) {
        for (size_t i = 0; i < fsz; ++i) frame[i] += priv[t][i];
        s->paths += npaths[t]; s->raysExt += counters[t].ext; s->raysShadow += counters[t].shadow; s->vertices += counters[t].vertices;
    }
    return KZ_OK;
}

/* block.cpp:39-45, color.h:93-98, common.cpp:352-366, bitmap.cpp:46-54 */
extern "C" int kzo_resolve(kzo_scene *s, const float *frame, float *rgb_linear, uint8_t *srgb8) {
    if (!s || !frame) return fail(KZ_ERR_INVALID, "null argument");
    int b = s->border, W = s->sc.camera.width, H = s->sc.camera.height, cols = W + 2 * b;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float *p = frame + 4 * ((size_t)(y + b) * cols + (x + b));
            V3 c = p[3] != 0 ? V3(p[0], p[1], p[2]) / p[3] : V3(0.0f);
            size_t o = 3 * ((size_t)y * W + x);
            if (rgb_linear) { rgb_linear[o] = c.x; rgb_linear[o + 1] = c.y; rgb_linear[o + 2] = c.z; }
            if (srgb8) {
                V3 t = toSRGB(c);
                srgb8[o] = (uint8_t)clampf(255.f * t.x, 0.f, 255.f);
                srgb8[o + 1] = (uint8_t)clampf(255.f * t.y, 0.f, 255.f);
                srgb8[o + 2] = (uint8_t)clampf(255.f * t.z, 0.f, 255.f);
            }
        }
    return KZ_OK;
}

extern "C" int kzo_stats(kzo_scene *s, kz_stats *out) {
    if (!s || !out) return fail(KZ_ERR_INVALID, "null argument");
    std::memset(out, 0, sizeof(*out));
    out->paths = s->paths; out->rays_extension = s->raysExt; out->rays_shadow = s->raysShadow; out->vertices = s->vertices;
    out->bvh_nodes = s->sc.accel.nodes.size();
    return KZ_OK;
}

/* ----------------------------------------------------------------- probes */
extern "C" uint64_t kzo_hash_pixel_seed(int32_t x, int32_t y, uint64_t seed) { return hashPixelSeed(x, y, seed); }
extern "C" uint64_t kzo_hash_pixel_dim_seed(int32_t x, int32_t y, uint32_t dim, uint64_t seed) { return hashPixelDimSeed(x, y, dim, seed); }
extern "C" uint64_t kzo_mix_bits(uint64_t v) { return mixBits(v); }
extern "C" uint32_t kzo_permute(uint32_t i, uint32_t l, uint32_t p) { return permute(i, l, p); }
extern "C" void kzo_pcg32_stream(uint64_t seed, uint64_t delta, uint32_t *out, int n) {
    Pcg32 r; r.seed(seed); r.advance(delta);
    for (int i = 0; i < n; ++i) out[i] = r.nextUInt();
}
extern "C" float kzo_pcg32_float(uint64_t seed, uint64_t delta) {
    Pcg32 r; r.seed(seed); r.advance(delta);
    return r.nextFloat();
}
extern "C" int kzo_bsdf_query(kzo_scene *s, int bsdf, int mode, const float wi[3], const float wo[3], const float uv[2],
                              float accR, float sample1, const float sample2[2], float out[8]) {
    if (!s || bsdf < 0 || bsdf >= (int)s->sc.bsdfs.size()) return fail(KZ_ERR_INVALID, "bad bsdf");
    Intersection its;
    its.shFrame.s = V3(1, 0, 0); its.shFrame.t = V3(0, 1, 0); its.shFrame.n = V3(0, 0, 1);
    its.geoFrame = its.shFrame;
    its.dpdu = V3(1, 0, 0); its.dpdv = V3(0, 1, 0);
    its.uv = V2{uv[0], uv[1]};
    its.accumulatedRoughness = accR;
    for (int i = 0; i < 8; ++i) out[i] = 0.f;
    if (mode == 2) {
        BSDFQueryRecord bRec(V3(wi[0], wi[1], wi[2]));
        bRec.uv = its.uv; bRec.its = its;
        V3 w = bsdfSample(s->sc, bsdf, bRec, sample1, V2{sample2[0], sample2[1]});
        out[0] = w.x; out[1] = w.y; out[2] = w.z;
        if (!iszero(w)) { out[3] = bRec.wo.x; out[4] = bRec.wo.y; out[5] = bRec.wo.z; }
        out[6] = (float)bRec.measure;
        out[7] = iszero(w) ? 0.f : bsdfPdf(s->sc, bsdf, bRec);
        return KZ_OK;
    }
    BSDFQueryRecord bRec(V3(wi[0], wi[1], wi[2]), V3(wo[0], wo[1], wo[2]), ESolidAngle);
    bRec.uv = its.uv; bRec.its = its;
    if (mode == 0) { V3 f = bsdfEval(s->sc, bsdf, bRec); out[0] = f.x; out[1] = f.y; out[2] = f.z; }
    else out[0] = bsdfPdf(s->sc, bsdf, bRec);
    return KZ_OK;
}

/* ------------------------------------------------------------------ math probe (tests/golden/math_kat.json) */
/* One shading-math helper by name on raw floats: lets the CPU suite check the restatements against vectors produced by the
 * reference's own function bodies (oracle/ref_math_kat.cpp) on machines where the reference is not mounted. */
extern "C" int kzo_math_probe(const char *fn, const float *in, int n_in, float *out, int *n_out) {
    if (!fn || !in || !out || !n_out) return fail(KZ_ERR_INVALID, "null argument");
    const std::string f(fn);
    auto v3 = [&](int o) { return V3(in[o], in[o + 1], in[o + 2]); };
    auto v2 = [&](int o) { return V2{in[o], in[o + 1]}; };
    auto put3 = [&](V3 v) { out[0] = v.x; out[1] = v.y; out[2] = v.z; *n_out = 3; };
    auto put1 = [&](float v) { out[0] = v; *n_out = 1; };
    auto need = [&](int n) { return n_in >= n; };
    if (f == "roughnessToAlpha" && need(2)) { V2 a = roughnessToAlpha(in[0], in[1]); out[0] = a.x; out[1] = a.y; *n_out = 2; }
    else if (f == "lambda" && need(5)) put1(ggxLambda(v3(0), v2(3)));
    else if (f == "smithG1" && need(8)) put1(smithG1(v3(0), v3(3), v2(6)));
    else if (f == "smithG2" && need(11)) put1(smithG2(v3(0), v3(3), v3(6), v2(9)));
    else if (f == "ggxNDF" && need(5)) put1(ggxNDF(v3(0), v2(3)));
    else if (f == "ggxVNDF" && need(8)) put1(ggxVNDF(v3(0), v3(3), v2(6)));
    else if (f == "sampleVNDF" && need(7)) put3(sampleGGXVNDF(v3(0), v2(3), v2(5)));
    else if (f == "schlick" && need(4)) put3(schlickFresnel(v3(0), in[3]));
    else if (f == "ggxSmithBRDF" && need(11)) put3(ggxSmithBRDF(v3(0), v3(3), v3(6), in[9], in[10]));
    else if (f == "coordinateSystem" && need(3)) { V3 b, c; coordinateSystem(v3(0), b, c); out[0] = b.x; out[1] = b.y; out[2] = b.z; out[3] = c.x; out[4] = c.y; out[5] = c.z; *n_out = 6; }
    else if (f == "frameToLocal" && need(6)) put3(Frame(v3(0)).toLocal(v3(3)));
    else if (f == "frameToWorld" && need(6)) put3(Frame(v3(0)).toWorld(v3(3)));
    else if (f == "reflect" && need(6)) put3(reflect(v3(0), v3(3)));
    else if (f == "refract" && need(7)) put3(refractDir(v3(0), v3(3), in[6]));
    else if (f == "fresnel" && need(3)) put1(fresnelExtInt(in[0], in[1], in[2]));
    else if (f == "fresnelDielectric" && need(2)) { float t = 0.f; out[0] = fresnelDielectric(in[0], in[1], t); out[1] = t; *n_out = 2; }
    else if (f == "squareToUniformDisk" && need(2)) { V2 d = squareToUniformDisk(v2(0)); out[0] = d.x; out[1] = d.y; *n_out = 2; }
    else if (f == "squareToCosineHemisphere" && need(2)) put3(squareToCosineHemisphere(v2(0)));
    else if (f == "squareToBeckmann" && need(3)) put3(squareToBeckmann(v2(0), in[2]));
    else if (f == "squareToBeckmannPdf" && need(4)) put1(squareToBeckmannPdf(v3(0), in[3]));
    else if (f == "dpdfSample" && need(2)) {            /* in = weights..., sample value (dpdf.h:35-104 append / normalize / sample) */
        std::vector<float> cdf(1, 0.f);
        for (int i = 0; i + 1 < n_in; ++i) cdf.push_back(cdf.back() + in[i]);
        const float norm = 1.0f / cdf.back();
        for (size_t i = 1; i < cdf.size(); ++i) cdf[i] *= norm;
        cdf.back() = 1.0f;
        put1((float)cdfSample(cdf, in[n_in - 1]));
    } else if ((f == "kissEval" || f == "kissPdf" || f == "kissSample") && need(22)) {
        /* in = base rgb, roughness, metallic, anisotropy, specular, specularTint, sheen, sheenTint, clearcoat, clearcoatRoughness,
         *      wi, wo, accumulatedRoughness, sample1, sample2 -- constant textures (bsdf.cpp:1175-1371) */
        SceneData sc;
        kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
        t.color[0] = in[0]; t.color[1] = in[1]; t.color[2] = in[2]; sc.textures.push_back(t);
        t.color[0] = t.color[1] = t.color[2] = in[3]; sc.textures.push_back(t);
        t.color[0] = t.color[1] = t.color[2] = in[4]; sc.textures.push_back(t);
        kz_bsdf_desc m; memset(&m, 0, sizeof(m)); m.type = KZ_BSDF_KISS; m.base_color = 0; m.roughness = 1; m.metallic = 2;
        m.anisotropy = in[5]; m.specular = in[6]; m.specular_tint = in[7]; m.sheen = in[8]; m.sheen_tint = in[9]; m.clearcoat = in[10]; m.clearcoat_roughness = in[11];
        if (f == "kissSample") {
            BSDFQueryRecord r(v3(12)); r.its.accumulatedRoughness = in[18];
            const V3 w = kissSample(sc, m, r, in[19], v2(20));
            const bool z = iszero(w);
            out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = z ? 0.f : r.wo.x; out[4] = z ? 0.f : r.wo.y; out[5] = z ? 0.f : r.wo.z; *n_out = 6;
        } else {
            BSDFQueryRecord r(v3(12), v3(15), ESolidAngle); r.its.accumulatedRoughness = in[18];
            if (f == "kissEval") put3(kissEval(sc, m, r)); else put1(kissPdf(sc, m, r));
        }
    } else if ((f == "diffuseEval" || f == "diffusePdf" || f == "diffuseSample") && need(11)) {     /* in = albedo, wi, wo, sample2 (bsdf.cpp:27-75) */
        SceneData sc;
        kz_bsdf_desc m; memset(&m, 0, sizeof(m)); m.type = KZ_BSDF_DIFFUSE; m.albedo[0] = in[0]; m.albedo[1] = in[1]; m.albedo[2] = in[2];
        sc.bsdfs.push_back(m);
        if (f == "diffuseSample") {
            BSDFQueryRecord r(v3(3));
            const V3 w = bsdfSample(sc, 0, r, 0.5f, v2(9));
            const bool z = iszero(w);
            out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = z ? 0.f : r.wo.x; out[4] = z ? 0.f : r.wo.y; out[5] = z ? 0.f : r.wo.z; *n_out = 6;
        } else {
            BSDFQueryRecord r(v3(3), v3(6), ESolidAngle);
            if (f == "diffuseEval") put3(bsdfEval(sc, 0, r)); else put1(bsdfPdf(sc, 0, r));
        }
    } else if ((f == "samplerIndependent" || f == "samplerStratified" || f == "samplerCorrelated") && need(6)) {
        /* in = px, py, sample index, requested sampleCount, seed & 0xFFFFFF, seed >> 24; out = nextPixel2D, next2D, then 4 x (5 x next1D, next2D) */
        SamplerCfg cfg; memset(&cfg.d, 0, sizeof(cfg.d));
        const uint32_t requested = (uint32_t)in[3];
        cfg.d.seed = (uint64_t)in[4] | ((uint64_t)in[5] << 24);
        cfg.d.sample_count = requested;
        if (f == "samplerStratified") {               /* sampler.cpp:83-93 */
            int res = 4; while ((uint32_t)(res * res) < requested) res++;
            cfg.d.type = KZ_SAMPLER_STRATIFIED; cfg.d.res_x = cfg.d.res_y = res; cfg.d.sample_count = (uint32_t)(res * res);
        } else if (f == "samplerCorrelated") {        /* sampler.cpp:178-189 */
            const int ry = (int)std::sqrt((double)requested), rx = (int)((requested + ry - 1) / ry);
            cfg.d.type = KZ_SAMPLER_CORRELATED; cfg.d.res_x = rx; cfg.d.res_y = ry; cfg.d.sample_count = (uint32_t)(rx * ry);
        } else cfg.d.type = KZ_SAMPLER_INDEPENDENT;
        Sampler sm; sm.cfg = &cfg; sm.generateSample((int32_t)in[0], (int32_t)in[1], (int)in[2]);
        int k = 0;
        V2 a = sm.nextPixel2D(); out[k++] = a.x; out[k++] = a.y;
        a = sm.next2D(); out[k++] = a.x; out[k++] = a.y;
        for (int v = 0; v < 4; ++v) { for (int j = 0; j < 5; ++j) out[k++] = sm.next1D(); a = sm.next2D(); out[k++] = a.x; out[k++] = a.y; }
        *n_out = k;
    } else if ((f == "extraEval" || f == "extraPdf" || f == "extraSample") && need(18)) {
        /* in = bsdf type, albedo / kd rgb, roughness, anisotropy, intIOR, extIOR, conductor (0 Au, 1 Cu, 2 Cr), wi, wo, sample1, sample2
         * (bsdf.cpp:98-276,629-1145: dielectric, mirror, lambertian, ggx, roughconductor, roughplastic, roughdielectric) */
        SceneData sc;
        kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
        t.color[0] = in[1]; t.color[1] = in[2]; t.color[2] = in[3]; sc.textures.push_back(t);
        kz_bsdf_desc m; memset(&m, 0, sizeof(m)); m.type = (int)in[0]; m.base_color = 0; m.roughness = m.metallic = m.normal_map = m.nested = -1;
        m.int_ior = in[6]; m.ext_ior = in[7]; m.anisotropy = in[5];
        m.alpha = m.type == KZ_BSDF_GGX ? in[4] : std::max(0.001f, in[4] * in[4]);          /* the constructors' max(MIN_ALPHA, sqr(roughness)) */
        static const float E[3][3] = {{0.1431189557f, 0.3749570432f, 1.4424785571f}, {0.2004376970f, 0.9240334304f, 1.1022119527f}, {4.3696828663f, 2.9167024892f, 1.6547005413f}};
        static const float K[3][3] = {{3.9831604247f, 2.3857207478f, 1.6032152899f}, {3.9129485033f, 2.4528477015f, 2.1421879552f}, {5.2064337956f, 4.2313645277f, 3.7549467933f}};
        const int mat = std::min(2, std::max(0, (int)in[8]));
        for (int c = 0; c < 3; ++c) { m.eta[c] = E[mat][c]; m.k[c] = K[mat][c]; m.albedo[c] = in[1 + c]; }
        sc.bsdfs.push_back(m);
        if (f == "extraSample") {
            BSDFQueryRecord r(v3(9)); r.uv = V2{0.5f, 0.5f};
            const V3 w = bsdfSample(sc, 0, r, in[15], v2(16));
            const bool z = iszero(w);
            out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = z ? 0.f : r.wo.x; out[4] = z ? 0.f : r.wo.y; out[5] = z ? 0.f : r.wo.z;
            out[6] = z ? 0.f : r.eta; out[7] = z ? 0.f : (float)(r.measure == EDiscrete); *n_out = 8;
        } else {
            BSDFQueryRecord r(v3(9), v3(12), ESolidAngle); r.uv = V2{0.5f, 0.5f};
            if (f == "extraEval") put3(bsdfEval(sc, 0, r)); else put1(bsdfPdf(sc, 0, r));
        }
    } else if ((f == "texColorRamp" || f == "texBlend" || f == "texBackgroundUV" || f == "texBackgroundDir" || (f == "sceneBackground" && need(20))) && need(18)) {
        /* in = three constant colours, ramp min / max, background intensity, blend mode, then five "child missing" flags (mask, input1, ramp's nested,
         * input2, background's nested): background(blend(mask = colorramp(c0), c1, c2)), texture.cpp:104-270 */
        SceneData sc;
        auto constant = [&](int at) { kz_texture_desc t; memset(&t, 0, sizeof(t)); t.type = KZ_TEX_CONSTANT; t.child[0] = t.child[1] = t.child[2] = -1;
                                      t.color[0] = in[at]; t.color[1] = in[at + 1]; t.color[2] = in[at + 2]; sc.textures.push_back(t); return (int)sc.textures.size() - 1; };
        const int n0 = constant(0), n1 = constant(3), n2 = constant(6);
        kz_texture_desc tr; memset(&tr, 0, sizeof(tr)); tr.type = KZ_TEX_COLORRAMP; tr.a = in[9]; tr.b = in[10]; tr.child[0] = in[15] != 0.f ? -1 : n0; tr.child[1] = tr.child[2] = -1;
        sc.textures.push_back(tr); const int nr = (int)sc.textures.size() - 1;
        kz_texture_desc tb; memset(&tb, 0, sizeof(tb)); tb.type = KZ_TEX_BLEND; tb.mode = (int)in[12];
        tb.child[0] = in[13] != 0.f ? -1 : nr; tb.child[1] = in[14] != 0.f ? -1 : n1; tb.child[2] = in[16] != 0.f ? -1 : n2;
        sc.textures.push_back(tb); const int nb = (int)sc.textures.size() - 1;
        kz_texture_desc tg; memset(&tg, 0, sizeof(tg)); tg.type = KZ_TEX_BACKGROUND; tg.a = in[11]; tg.child[0] = in[17] != 0.f ? -1 : nb; tg.child[1] = tg.child[2] = -1;
        sc.textures.push_back(tg); const int ng = (int)sc.textures.size() - 1;
        const V2 uv{0.25f, 0.75f};                                    /* constant leaves: the lookup position does not matter */
        if (f == "texColorRamp") put3(evalTextureUV(sc, nr, uv));
        else if (f == "texBlend") put3(evalTextureUV(sc, nb, uv));
        else if (f == "texBackgroundUV") put3(evalTextureUV(sc, ng, uv));
        else {
            kz_texture_desc td = tg; td.child[0] = n2; sc.textures.push_back(td);
            if (f == "texBackgroundDir") put3(evalTextureDir(sc, (int)sc.textures.size() - 1, V3(0.f, 0.f, 1.f)));
            else {      /* scene.cpp:54-79 over that node: in[18] = no background, in[19] odd = a NaN component in the direction */
                sc.background = in[18] != 0.f ? -1 : (int)sc.textures.size() - 1;
                V3 dir(0.f, 0.f, 1.f);
                const int k = (int)in[19];
                if (k == 1) dir.x = std::nanf(""); else if (k == 3) dir.y = std::nanf(""); else if (k == 5) dir.z = std::nanf("");
                put3(backgroundColor(sc, dir));
            }
        }
    } else if (f == "pmj02bnTileSize" && need(1)) {                   /* sampler.cpp:291: tile = 1 << (log4(65536) - log4(roundUpPow4(spp))) */
        const int spp = (int)in[0];
        put1((float)(1 << (log2i_int(65536) / 2 - log2i_int(roundUpPow4(spp)) / 2)));
    } else if (f == "toSRGB" && need(3)) put3(toSRGB(v3(0)));
    else if (f == "toLinearRGB" && need(3)) put3(toLinearRGB(v3(0)));
    else if (f == "luminance" && need(3)) put1(luminance(v3(0)));
    else return fail(KZ_ERR_INVALID, "unknown probe or too few inputs: " + f);
    return KZ_OK;
}
