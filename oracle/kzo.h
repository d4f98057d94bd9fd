/*
 * kzo.h -- CPU ORACLE for the nano-kazen render hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libkzgpu.so, the kazen host binary) never
 * links, loads or calls it; there is no CPU fallback in the product.
 *
 * It mirrors include/kzgpu.h function by function over the same POD tables, so tests
 * feed byte-identical inputs to the oracle and to the CUDA path.
 *
 * Parity pins (see DESIGN.md "Oracle"):
 *   - Hash / MixBits / pcg32 / permute: PINNED against known-answer vectors generated
 *     from the reference's own include/kazen/hash.h, pcg32.h and the permute() body of
 *     src/kazen/common.cpp:316-344 compiled in place (oracle/Makefile target _ref/ref_kat,
 *     vectors committed as tests/golden/sampler_kat.json).
 *   - Sampler composition, camera, post-intersection, kiss/diffuse/normalmap BSDFs, area
 *     light, Mesh::sample, path_mis, film: restated from the cited reference lines; the
 *     reference has no tests or numeric fixtures for them, checked against the shipped
 *     golden PNG means only.
 *   - Ray/triangle arithmetic lives in Embree 3.13.0 (find_package(embree 3.13.0),
 *     CMakeLists.txt:12), which is NOT in /root/reference and not installed: PARITY
 *     UNPINNED.  Restated from Embree's published robust ("Pluecker") single-ray triangle
 *     intersector; hit identity = argmin t over all triangles passing that test.
 *   - OpenImageIO texture filtering and the pbrt-v4 pmj02bn / blue-noise tables (missing
 *     blobs): PARITY UNPINNED.
 */
#ifndef KZO_H
#define KZO_H
#include "../include/kzgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kzo_scene kzo_scene;

/* Copies every table it needs; the caller may free the desc afterwards. */
int  kzo_scene_create(const kz_scene_desc *desc, kzo_scene **out);
void kzo_scene_destroy(kzo_scene *s);
const char *kzo_last_error(void);

/* accel.cpp:63-110 + Embree robust triangle test. brute=1: test every triangle (defines
 * hit identity); brute=0: conservative BVH2 (must return identical results; tested). */
int  kzo_trace(kzo_scene *s, const kz_ray *rays, size_t n, int shadow, int brute, int threads, kz_hit *hits);
int  kzo_occluded(kzo_scene *s, const kz_ray *rays, size_t n, float trace_bias, int threads,
                  uint8_t *occluded, uint8_t *segments);
int  kzo_sample_dump(kzo_scene *s, const int32_t *pixel_sample_triples, size_t n, const char *pattern, float *out);
int  kzo_camera_rays(kzo_scene *s, const float *samples4, size_t n, kz_ray *out);
int  kzo_render(kzo_scene *s, const kz_render_req *req, int threads, float *frame_rgbw);
int  kzo_frame_dims(const kzo_scene *s, int32_t *width, int32_t *height, int32_t *border);
int  kzo_resolve(kzo_scene *s, const float *frame_rgbw, float *rgb_linear, uint8_t *srgb8);
int  kzo_stats(kzo_scene *s, kz_stats *out);

/* Single-function probes used by the known-answer tests. */
uint64_t kzo_hash_pixel_seed(int32_t x, int32_t y, uint64_t seed);                 /* Hash(Point2i, uint64)          */
uint64_t kzo_hash_pixel_dim_seed(int32_t x, int32_t y, uint32_t dim, uint64_t seed);/* Hash(Point2i, uint32, uint64)  */
uint64_t kzo_mix_bits(uint64_t v);
uint32_t kzo_permute(uint32_t i, uint32_t l, uint32_t p);
/* seed(s); advance(delta); out[k] = nextUInt() for k < n */
void     kzo_pcg32_stream(uint64_t seed, uint64_t delta, uint32_t *out, int n);
float    kzo_pcg32_float(uint64_t seed, uint64_t delta);
/* One BSDF query in the local frame: mode 0 = eval (rgb), 1 = pdf (out[0]), 2 = sample
 * (out = weight rgb, wo xyz, measure flag).  accumulated_roughness feeds its.accumulatedRoughness. */
int  kzo_bsdf_query(kzo_scene *s, int bsdf, int mode, const float wi[3], const float wo[3], const float uv[2],
                    float accumulated_roughness, float sample1, const float sample2[2], float out[8]);
/* Filter table as ImageBlock tabulates it (block.cpp:13-21, rfilter.cpp:10-102).
 * kind: 0 gaussian(radius, stddev) 1 mitchell(radius,B,C) 2 tent 3 box */
void kzo_filter_table(int kind, float p0, float p1, float p2, float *radius_out, float table33[33]);
/* sampleToCamera as Camera::activate builds it (camera.cpp:35-62), row-major. */
void kzo_camera_matrix(int width, int height, float fov_deg, float near_clip, float far_clip, float out16[16]);
/* Area CDF of a light mesh as Mesh::activate builds it (mesh.cpp:24-45, dpdf.h:35-81). */
int  kzo_light_cdf(kzo_scene *s, int mesh, float *cdf_out /* n_triangles+1 */, float *normalization);

#ifdef __cplusplus
}
#endif
#endif
