/* ORACLE (test infrastructure) -- closest-hit queries.
 *
 * The reference delegates this to Embree 3.13.0 built with RTC_SCENE_FLAG_ROBUST
 * (src/kazen/accel.cpp:29-58, rtcIntersect1 at :98).  Embree is a third-party dependency
 * (CMakeLists.txt:12 find_package(embree 3.13.0 REQUIRED)) that is neither vendored under
 * /root/reference nor installed here: PARITY UNPINNED.  What follows restates Embree's
 * published robust single-ray triangle intersector (kernels/geometry/triangle_intersector_pluecker.h,
 * PlueckerIntersector1 + PlueckerHitM::finalize, and stable_triangle_normal in
 * kernels/geometry/intersector_epilog / common/math) from its public algorithm:
 *
 *   v_i = p_i - org;  e0 = v2-v0, e1 = v0-v1, e2 = v1-v2
 *   U = dot(cross(e0, v2+v0), D), V = dot(cross(e1, v0+v1), D), W = dot(cross(e2, v1+v2), D)
 *   UVW = U+V+W; eps = FLT_EPSILON*|UVW|; accept if min(U,V,W) >= -eps || max(U,V,W) <= eps
 *   Ng = stable_triangle_normal(e0,e1,e2); den = 2*dot(Ng,D); T = 2*dot(v0,Ng); t = rcp(den)*T
 *   accept if tnear <= t <= tfar and den != 0;  u = min(U*rcp(UVW),1), v = min(V*rcp(UVW),1)
 *
 * with Embree's AVX2 operation shapes: dot(a,b) = fma(a.x,b.x, fma(a.y,b.y, a.z*b.z)),
 * cross(a,b).x = fms(a.y,b.z, a.z*b.y) (fused multiply-subtract).  Embree's rcp() is a
 * hardware estimate plus one Newton step (not correctly rounded, CPU dependent); here it is
 * the correctly rounded 1/x, which is why the contract on t is "<= 2 ulp", not bit-exact.
 *
 * Hit identity is defined BVH-independently: the hit is the triangle with the smallest t
 * among all triangles passing the test within [tnear, tfar]; exact-t ties (traversal-order
 * dependent in Embree) are broken by smallest (geom_id, prim_id).  kzo_trace(brute=1)
 * evaluates exactly that definition; the BVH2 below is only an accelerator for it and must
 * return bit-identical results (tests/test_oracle.py::test_bvh_matches_brute).
 */
#ifndef KZO_ACCEL_H
#define KZO_ACCEL_H
#include "kzo_math.h"
#include "../include/kzgpu.h"
#include <vector>
#include <cfloat>
#include <atomic>
#include <functional>
#include <thread>

namespace kzo {

struct Tri { V3 p0, p1, p2; uint32_t geom, prim; };

struct HitRec { float t, u, v; uint32_t prim, geom; };

inline float fmsf(float a, float b, float c) { return fmaf(a, b, -c); }          /* a*b - c, fused */
inline float edot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
inline V3 ecross(V3 a, V3 b) {
    return V3(fmsf(a.y, b.z, a.z * b.y), fmsf(a.z, b.x, a.x * b.z), fmsf(a.x, b.y, a.y * b.x));
}
inline V3 stableTriangleNormal(V3 a, V3 b, V3 c) {
    float ab_x = a.z * b.y, ab_y = a.x * b.z, ab_z = a.y * b.x;
    float bc_x = b.z * c.y, bc_y = b.x * c.z, bc_z = b.y * c.x;
    V3 cross_ab(fmsf(a.y, b.z, ab_x), fmsf(a.z, b.x, ab_y), fmsf(a.x, b.y, ab_z));
    V3 cross_bc(fmsf(b.y, c.z, bc_x), fmsf(b.z, c.x, bc_y), fmsf(b.x, c.y, bc_z));
    bool sx = std::fabs(ab_x) < std::fabs(bc_x);
    bool sy = std::fabs(ab_y) < std::fabs(bc_y);
    bool sz = std::fabs(ab_z) < std::fabs(bc_z);
    return V3(sx ? cross_ab.x : cross_bc.x, sy ? cross_ab.y : cross_bc.y, sz ? cross_ab.z : cross_bc.z);
}

/* Returns true and fills (t,u,v) when the triangle is hit within [tnear, tfar]. */
inline bool plueckerIntersect(V3 org, V3 dir, float tnear, float tfar, const Tri &tri, float &t_out, float &u_out, float &v_out) {
    V3 v0 = tri.p0 - org, v1 = tri.p1 - org, v2 = tri.p2 - org;
    V3 e0 = v2 - v0, e1 = v0 - v1, e2 = v1 - v2;
    float U = edot(ecross(e0, v2 + v0), dir);
    float V = edot(ecross(e1, v0 + v1), dir);
    float W = edot(ecross(e2, v1 + v2), dir);
    float UVW = U + V + W;
    float eps = FLT_EPSILON * std::fabs(UVW);
    float mn = std::fmin(U, std::fmin(V, W)), mx = std::fmax(U, std::fmax(V, W));
    if (!(mn >= -eps || mx <= eps)) return false;
    V3 Ng = stableTriangleNormal(e0, e1, e2);
    float d = edot(Ng, dir);
    float den = d + d;
    float T0 = edot(v0, Ng);
    float T = T0 + T0;
    float t = (1.0f / den) * T;
    if (!(tnear <= t && t <= tfar)) return false;
    if (den == 0.0f) return false;
    float rcpUVW = std::fabs(UVW) < 1e-18f ? 0.0f : 1.0f / UVW;     /* min_rcp_input */
    t_out = t;
    u_out = std::fmin(U * rcpUVW, 1.0f);
    v_out = std::fmin(V * rcpUVW, 1.0f);
    return true;
}

/* Deterministic closest-hit update: smaller t wins; equal t -> smaller (geom, prim). */
inline void considerHit(HitRec &best, float t, float u, float v, uint32_t geom, uint32_t prim) {
    bool better = t < best.t ||
                  (t == best.t && (geom < best.geom || (geom == best.geom && prim < best.prim)));
    if (better) { best.t = t; best.u = u; best.v = v; best.geom = geom; best.prim = prim; }
}

struct Bvh2Node {
    float lo[3], hi[3];
    int32_t left;    /* internal: index of left child (right = left+1); leaf: first triangle */
    int32_t count;   /* 0 = internal, >0 = leaf triangle count */
    int32_t axis;    /* internal: split axis (left child holds the lower centroids) */
};

struct Accel {
    std::vector<Tri> tris;          /* scene order: mesh by mesh, face by face */
    std::vector<Tri> ordered;       /* BVH leaf order */
    std::vector<Bvh2Node> nodes;
    float slack = 0.f;              /* absolute box inflation, see build() */

    void build();
    HitRec traceBrute(const kz_ray &r) const;
    HitRec traceBvh(const kz_ray &r) const;
};

inline HitRec missRec(float tfar) { return HitRec{tfar, 0.f, 0.f, KZ_INVALID_ID, KZ_INVALID_ID}; }

inline HitRec Accel::traceBrute(const kz_ray &r) const {
    V3 org(r.o[0], r.o[1], r.o[2]), dir(r.d[0], r.d[1], r.d[2]);
    HitRec best = missRec(r.tmax);
    for (const Tri &tri : tris) {
        float t, u, v;
        if (plueckerIntersect(org, dir, r.tmin, best.t, tri, t, u, v)) considerHit(best, t, u, v, tri.geom, tri.prim);
    }
    return best;
}

inline void Accel::build() {
    size_t n = tris.size();
    ordered = tris;
    nodes.clear();
    if (n == 0) return;
    struct Ref { float lo[3], hi[3], c[3]; uint32_t idx; };
    std::vector<Ref> refs(n);
    float maxAbs = 0.f;
    for (size_t i = 0; i < n; ++i) {
        const Tri &t = tris[i];
        const float px[3][3] = {{t.p0.x, t.p0.y, t.p0.z}, {t.p1.x, t.p1.y, t.p1.z}, {t.p2.x, t.p2.y, t.p2.z}};
        for (int a = 0; a < 3; ++a) {
            refs[i].lo[a] = std::min(px[0][a], std::min(px[1][a], px[2][a]));
            refs[i].hi[a] = std::max(px[0][a], std::max(px[1][a], px[2][a]));
            refs[i].c[a] = 0.5f * (refs[i].lo[a] + refs[i].hi[a]);
            maxAbs = std::max(maxAbs, std::max(std::fabs(refs[i].lo[a]), std::fabs(refs[i].hi[a])));
        }
        refs[i].idx = (uint32_t)i;
    }
    /* The triangle test accepts points up to ~eps*edge outside the triangle and subtracts the
     * ray origin with one rounding per vertex; the boxes are inflated per ray in traceBvh by
     * an amount that covers both (see there). */
    slack = maxAbs;
    /* median split; subtrees are independent, so the upper levels fork a thread per right half (the result -- hit bytes -- does
     * not depend on node numbering).  Nodes come from a preallocated array through an atomic cursor. */
    nodes.assign(2 * n + 1, Bvh2Node());
    std::atomic<int> next{1};
    std::vector<Tri> out(n);
    std::function<void(int, size_t, size_t, int)> split = [&](int node, size_t jb, size_t je, int fork_levels) {
        struct Job { int node; size_t b, e; };
        std::vector<Job> stack;
        std::vector<std::thread> forked;
        stack.push_back(Job{node, jb, je});
        while (!stack.empty()) {
            Job j = stack.back(); stack.pop_back();
            float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (size_t i = j.b; i < j.e; ++i)
                for (int a = 0; a < 3; ++a) {
                    lo[a] = std::min(lo[a], refs[i].lo[a]); hi[a] = std::max(hi[a], refs[i].hi[a]);
                    clo[a] = std::min(clo[a], refs[i].c[a]); chi[a] = std::max(chi[a], refs[i].c[a]);
                }
            Bvh2Node &nd = nodes[j.node];
            for (int a = 0; a < 3; ++a) { nd.lo[a] = lo[a]; nd.hi[a] = hi[a]; }
            size_t cnt = j.e - j.b;
            int axis = 0;
            for (int a = 1; a < 3; ++a) if (chi[a] - clo[a] > chi[axis] - clo[axis]) axis = a;
            if (cnt <= 4 || !(chi[axis] > clo[axis])) {
                nd.left = (int32_t)j.b; nd.count = (int32_t)cnt;
                continue;
            }
            size_t mid = (j.b + j.e) / 2;
            std::nth_element(refs.begin() + j.b, refs.begin() + mid, refs.begin() + j.e,
                             [axis](const Ref &x, const Ref &y) { return x.c[axis] < y.c[axis]; });
            int l = next.fetch_add(2);
            nd.left = l; nd.count = 0; nd.axis = axis;
            if (fork_levels > 0 && cnt > (1u << 16)) {
                --fork_levels;
                const int fl = fork_levels;
                forked.emplace_back([&split, l, mid, j, fl]() { split(l + 1, mid, j.e, fl); });
                stack.push_back(Job{l, j.b, mid});
            } else {
                stack.push_back(Job{l, j.b, mid});
                stack.push_back(Job{l + 1, mid, j.e});
            }
        }
        for (std::thread &t : forked) t.join();
    };
    {
        unsigned hw = std::thread::hardware_concurrency();
        int levels = 0;
        while ((1u << levels) < std::max(1u, hw) && levels < 6) ++levels;
        split(0, 0, n, levels);
    }
    nodes.resize((size_t)next.load());
    for (size_t i = 0; i < n; ++i) out[i] = tris[refs[i].idx];
    ordered.swap(out);
}

inline HitRec Accel::traceBvh(const kz_ray &r) const {
    HitRec best = missRec(r.tmax);
    if (nodes.empty()) return best;
    V3 org(r.o[0], r.o[1], r.o[2]), dir(r.d[0], r.d[1], r.d[2]);
    /* NaN rays fail every comparison of the triangle test: a miss by construction (and not worth walking the whole tree) */
    if (std::isnan(org.x + org.y + org.z) || std::isnan(dir.x) || std::isnan(dir.y) || std::isnan(dir.z) || std::isnan(r.tmin) || std::isnan(r.tmax)) {
        bool inf_only = !std::isnan(org.x) && !std::isnan(org.y) && !std::isnan(org.z);
        if (!inf_only || std::isnan(dir.x) || std::isnan(dir.y) || std::isnan(dir.z) || std::isnan(r.tmin) || std::isnan(r.tmax)) return best;
    }
    /* Conservative culling in double precision: inflate boxes by 1e-5 of the magnitude of
     * everything involved (far more than the triangle test's few-ulp tolerance), and the
     * t-interval by a relative 1e-5.  The BVH only prunes; identity comes from the test. */
    double mag = std::max((double)slack, (double)std::max(std::fabs(org.x), std::max(std::fabs(org.y), std::fabs(org.z))));
    double pad = 1e-5 * mag + 1e-30;
    double o[3] = {org.x, org.y, org.z}, d[3] = {dir.x, dir.y, dir.z};
    double inv[3];
    for (int a = 0; a < 3; ++a) inv[a] = d[a] != 0.0 ? 1.0 / d[a] : 0.0;
    int stack[128]; int sp = 0; stack[sp++] = 0;
    while (sp) {
        const Bvh2Node &nd = nodes[stack[--sp]];
        double t0 = (double)r.tmin, t1 = (double)best.t;
        if (std::isfinite(t1)) t1 = t1 + 1e-5 * std::fabs(t1) + 1e-30;
        t0 = t0 - 1e-5 * std::fabs(t0) - 1e-30;
        bool hitbox = true;
        for (int a = 0; a < 3 && hitbox; ++a) {
            double lo = (double)nd.lo[a] - pad, hi = (double)nd.hi[a] + pad;
            if (d[a] == 0.0) {
                if (o[a] < lo || o[a] > hi) hitbox = false;
            } else {
                /* reciprocal instead of two divisions: the extra rounding (1 ulp of a double) is far inside the 1e-9 margin below */
                double ta = (lo - o[a]) * inv[a], tb = (hi - o[a]) * inv[a];
                if (ta > tb) std::swap(ta, tb);
                ta -= 1e-9 * std::fabs(ta); tb += 1e-9 * std::fabs(tb);
                if (ta > t0) t0 = ta;
                if (tb < t1) t1 = tb;
                if (t0 > t1) hitbox = false;
            }
        }
        if (!hitbox) continue;
        if (nd.count > 0) {
            for (int i = 0; i < nd.count; ++i) {
                const Tri &tri = ordered[nd.left + i];
                float t, u, v;
                if (plueckerIntersect(org, dir, r.tmin, best.t, tri, t, u, v)) considerHit(best, t, u, v, tri.geom, tri.prim);
            }
        } else if (d[nd.axis] >= 0.0) {      /* near child last on the stack = visited first (order does not change the result) */
            stack[sp++] = nd.left + 1; stack[sp++] = nd.left;
        } else {
            stack[sp++] = nd.left; stack[sp++] = nd.left + 1;
        }
    }
    return best;
}

}  // namespace kzo
#endif
