/* tests/hostemu/emu.cpp -- DEVELOPMENT/TEST HARNESS, NOT A PRODUCT PATH.
 *
 * The build container has no GPU.  This file compiles the device routines of
 * nano-kazen_b200/csrc (kz_traverse.h, kz_sampler.h, kz_shade.h, kz_path.h: the very source the
 * CUDA kernels wrap) with g++ and runs them item by item on the CPU, so the CPU-side test suite
 * can check the traversal / sampler / wavefront logic against the oracle before GPU time is
 * spent.  libkzgpu.so never links this, and kzgpu_* has no CPU path: it fails with
 * KZ_ERR_NO_DEVICE when there is no CUDA device.
 */
#include "../../nano-kazen_b200/csrc/kz_path.h"
#include "../../nano-kazen_b200/csrc/kz_host_scene.h"
#include <atomic>
#include <thread>

struct kzemu {
    KzHostScene hs;
    kzbvh::Built bvh;
    KzCounters total{0, 0, 0, 0};
};

template <typename F> static void pfor(size_t n, F f) {
    int threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
    if (n < 256) { f(0, n); return; }
    std::atomic<size_t> next{0};
    size_t chunk = std::max<size_t>(64, n / ((size_t)threads * 32));
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&]() { for (;;) { size_t b = next.fetch_add(chunk); if (b >= n) break; f(b, std::min(n, b + chunk)); } });
    for (auto &th : pool) th.join();
}

extern "C" {

int kzemu_create(const kz_scene_desc *d, kzemu **out) {
    kzemu *e = new kzemu();
    if (!e->hs.flatten(d)) { fprintf(stderr, "kzemu: %s\n", e->hs.error.c_str()); delete e; return KZ_ERR_INVALID; }
    kzbvh::buildHostSah(e->hs.tris, 0, e->bvh);
    e->hs.sc.nodes = e->bvh.nodes.data();
    e->hs.sc.tris = e->bvh.tris.data();
    e->hs.sc.n_nodes = (uint32_t)e->bvh.nodes.size();
    e->hs.sc.n_tris = (uint32_t)(e->bvh.tris.size() / 3);
    e->hs.sc.scene_max_abs = e->bvh.max_abs;
    *out = e;
    return KZ_OK;
}
void kzemu_destroy(kzemu *e) { delete e; }
int kzemu_bvh_info(kzemu *e, uint64_t *nodes, uint64_t *tris, int *depth) {
    *nodes = e->bvh.nodes.size(); *tris = e->bvh.tris.size() / 3; *depth = e->bvh.depth;
    return KZ_OK;
}

int kzemu_trace(kzemu *e, const kz_ray *rays, size_t n, kz_hit *hits) {
    const KzScene &sc = e->hs.sc;
    pfor(n, [&](size_t b, size_t en) {
        KzStackRef stk;
        for (size_t i = b; i < en; ++i) {
            const kz_ray &r = rays[i];
            KzHit h = kz_trace(sc, stk, r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], r.tmin, r.tmax, false);
            hits[i].t = h.t; hits[i].u = h.u; hits[i].v = h.v; hits[i].prim_id = h.prim; hits[i].geom_id = h.geom;
        }
    });
    return KZ_OK;
}

/* kz_warp_trace (kz_kernels.cuh), restated lane by lane: 32 lanes share a fetch cursor, every iteration the lanes still at work make one
 * node step, the lanes that hold triangles test them while at least 1/`den` of the lanes at work hold some -- otherwise a lane with a node
 * group in hand puts its triangle group on its stack (two entries) and one without waits with it in registers -- and lanes whose groups are
 * exhausted pop; finished lanes are refilled once `nw` lane-iterations were lost.  This is the only caller of kz_trav_postpone and of the
 * triangle-group branch of kz_trav_pop on the host.  out_events[12] = {postponed groups, groups waited with, refills, then warp-level iteration / lane counts}. */
int kzemu_trace_warp(kzemu *e, const kz_ray *rays, size_t n, kz_hit *hits, int den, int nw, uint64_t *out_events) {
    const KzScene &sc = e->hs.sc;
    struct Lane { KzTrav t; KzLocalStack ls; bool active = false, finished = false, wait = false; size_t item = 0; };
    std::vector<Lane> L(32);
    KzStackRef stk;
    size_t cursor = 0; bool exhausted = false;
    uint64_t ev_post = 0, ev_wait = 0, ev_refill = 0;
    uint64_t c_node_it = 0, c_node_ln = 0, c_tri_it = 0, c_tri_ln = 0, c_pop_it = 0, c_pop_ln = 0, c_iter = 0, c_iter_ln = 0, c_post_it = 0;
    for (Lane &l : L) { l.t.sp = 0; l.t.ng_y = 0u; l.t.tg_y = 0u; }
    for (;;) {
        for (Lane &l : L) if (l.finished) {
            const KzHit &h = l.t.best;
            hits[l.item].t = h.t; hits[l.item].u = h.u; hits[l.item].v = h.v; hits[l.item].prim_id = h.prim; hits[l.item].geom_id = h.geom;
            l.active = false; l.finished = false;
        }
        if (!exhausted) {
            int idle = 0; for (Lane &l : L) idle += !l.active;
            if (idle) {
                ++ev_refill;
                for (Lane &l : L) if (!l.active && cursor < n) {
                    l.item = cursor++;
                    const kz_ray &r = rays[l.item];
                    kz_trav_init(sc, l.t, r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], r.tmin, r.tmax);
                    l.active = true;
                }
                exhausted = cursor >= n;
            }
        }
        bool any = false; for (Lane &l : L) any |= l.active;
        if (!any) break;
        int lost = 0;
        for (;;) {
            std::vector<Lane *> S;
            for (Lane &l : L) if (l.active && !l.finished) S.push_back(&l);
            if (S.empty()) break;
            { int k = 0; for (Lane *l : S) if (l->t.tg_y == 0u && l->t.ng_y > 0x00FFFFFFu) { kz_trav_node(sc, l->t, stk, l->ls); ++k; }
              if (k) { ++c_node_it; c_node_ln += (uint64_t)k; } }
            ++c_iter; c_iter_ln += S.size();
            const int total = (int)S.size();
            for (;;) {
                std::vector<Lane *> T;
                for (Lane *l : S) if (l->t.tg_y != 0u && !l->wait) T.push_back(l);
                if (T.empty()) break;
                if ((int)T.size() * den < total) {
                    ++c_post_it;
                    for (Lane *l : T) {
                        if (l->t.ng_y > 0x00FFFFFFu && l->t.sp < KZ_POSTPONE_SP_LIMIT) { kz_trav_postpone(l->t, stk, l->ls); ++ev_post; }
                        else { l->wait = true; ++ev_wait; }          /* leaves the loop with its group in registers */
                    }
                    break;
                }
                for (Lane *l : T) kz_trav_tri(sc, l->t);
                ++c_tri_it; c_tri_ln += T.size();
            }
            { int k = 0;
              for (Lane *l : S) {
                l->wait = false;
                if (l->t.tg_y == 0u && l->t.ng_y <= 0x00FFFFFFu) {
                    if (l->t.sp == 0) l->finished = true; else { kz_trav_pop(l->t, stk, l->ls); ++k; }
                }
              }
              if (k) { ++c_pop_it; c_pop_ln += (uint64_t)k; } }
            if (!exhausted) { lost += 32 - total; if (lost >= nw) break; }
        }
    }
    if (out_events) {
        out_events[0] = ev_post; out_events[1] = ev_wait; out_events[2] = ev_refill;
        /* warp-level iteration counts and the lanes that took part (tools/warp_sim.py) */
        out_events[3] = c_iter; out_events[4] = c_iter_ln; out_events[5] = c_node_it; out_events[6] = c_node_ln; out_events[7] = c_tri_it; out_events[8] = c_tri_ln;
        out_events[9] = c_pop_it; out_events[10] = c_pop_ln; out_events[11] = c_post_it;
    }
    return KZ_OK;
}

/* A scheduling policy that is NOT in the product (tools/warp_sim.py evaluates it against kzemu_trace_warp): every lane keeps the triangle
 * groups it meets in a list of its own and goes on with node steps; the leaf tests run when at least 1/`den` of the lanes at work
 * have groups listed (or nothing else is left to do), one triangle per lane and iteration, until fewer than 1/`den2` of the lanes still hold
 * triangles.  Same hits as the plain loop (checked by the caller); out_events as kzemu_trace_warp. */
int kzemu_trace_warp_lists(kzemu *e, const kz_ray *rays, size_t n, kz_hit *hits, int den, int den2, int nw, uint64_t *out_events) {
    const KzScene &sc = e->hs.sc;
    struct Grp { uint32_t x, y, m; };
    struct Lane { KzTrav t; KzLocalStack ls; bool active = false, finished = false; size_t item = 0; std::vector<Grp> pend; };
    std::vector<Lane> L(32);
    KzStackRef stk;
    size_t cursor = 0; bool exhausted = false;
    uint64_t ev_refill = 0, c_node_it = 0, c_node_ln = 0, c_tri_it = 0, c_tri_ln = 0, c_pop_it = 0, c_pop_ln = 0, c_iter = 0, c_iter_ln = 0, c_listed = 0, c_fire = 0;
    for (Lane &l : L) { l.t.sp = 0; l.t.ng_y = 0u; l.t.tg_y = 0u; }
    for (;;) {
        for (Lane &l : L) if (l.finished) {
            const KzHit &h = l.t.best;
            hits[l.item].t = h.t; hits[l.item].u = h.u; hits[l.item].v = h.v; hits[l.item].prim_id = h.prim; hits[l.item].geom_id = h.geom;
            l.active = false; l.finished = false;
        }
        if (!exhausted) {
            int idle = 0; for (Lane &l : L) idle += !l.active;
            if (idle) {
                ++ev_refill;
                for (Lane &l : L) if (!l.active && cursor < n) {
                    l.item = cursor++;
                    const kz_ray &r = rays[l.item];
                    kz_trav_init(sc, l.t, r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], r.tmin, r.tmax);
                    l.active = true;
                }
                exhausted = cursor >= n;
            }
        }
        bool any = false; for (Lane &l : L) any |= l.active;
        if (!any) break;
        int lost = 0;
        for (;;) {
            std::vector<Lane *> S;
            for (Lane &l : L) if (l.active && !l.finished) S.push_back(&l);
            if (S.empty()) break;
            const int total = (int)S.size();
            ++c_iter; c_iter_ln += S.size();
            { int k = 0;
              for (Lane *l : S) if (l->t.ng_y > 0x00FFFFFFu) {
                  kz_trav_node(sc, l->t, stk, l->ls); ++k;
                  if (l->t.tg_y != 0u) { l->pend.push_back(Grp{l->t.tg_x, l->t.tg_y, l->t.tg_m}); l->t.tg_y = 0u; ++c_listed; }
              }
              if (k) { ++c_node_it; c_node_ln += (uint64_t)k; } }
            int holders = 0, busy = 0;
            for (Lane *l : S) { holders += !l->pend.empty(); busy += (l->t.ng_y > 0x00FFFFFFu || l->t.sp > 0); }
            if (holders && (holders * den >= total || busy == 0)) {
                ++c_fire;
                for (;;) {
                    int k = 0;
                    for (Lane *l : S) {
                        if (l->t.tg_y == 0u && !l->pend.empty()) { const Grp g = l->pend.back(); l->pend.pop_back(); l->t.tg_x = g.x; l->t.tg_y = g.y; l->t.tg_m = g.m; }
                        k += l->t.tg_y != 0u;
                    }
                    if (k == 0) break;
                    if (k * den2 < total && busy > 0) {       /* the tail goes back on the lists */
                        for (Lane *l : S) if (l->t.tg_y != 0u) { l->pend.push_back(Grp{l->t.tg_x, l->t.tg_y, l->t.tg_m}); l->t.tg_y = 0u; }
                        break;
                    }
                    for (Lane *l : S) if (l->t.tg_y != 0u) kz_trav_tri(sc, l->t);
                    ++c_tri_it; c_tri_ln += (uint64_t)k;
                }
            }
            { int k = 0;
              for (Lane *l : S) if (l->t.ng_y <= 0x00FFFFFFu) {
                  if (l->t.sp > 0) { kz_trav_pop(l->t, stk, l->ls); ++k; }
                  else if (l->pend.empty()) l->finished = true;
              }
              if (k) { ++c_pop_it; c_pop_ln += (uint64_t)k; } }
            if (!exhausted) { lost += 32 - total; if (lost >= nw) break; }
        }
    }
    if (out_events) {
        out_events[0] = c_listed; out_events[1] = c_fire; out_events[2] = ev_refill;
        out_events[3] = c_iter; out_events[4] = c_iter_ln; out_events[5] = c_node_it; out_events[6] = c_node_ln; out_events[7] = c_tri_it; out_events[8] = c_tri_ln;
        out_events[9] = c_pop_it; out_events[10] = c_pop_ln; out_events[11] = 0;
    }
    return KZ_OK;
}

/* Traversal statistics of the plain per-ray loop (tuning aid): out = {node steps, triangle tests, triangle groups, accepted hits, max stack}. */
int kzemu_trace_stats(kzemu *e, const kz_ray *rays, size_t n, uint64_t *out) {
    const KzScene &sc = e->hs.sc;
    std::atomic<uint64_t> a_nodes{0}, a_tris{0}, a_groups{0}, a_hits{0}, a_sp{0};
    pfor(n, [&](size_t b, size_t en) {
        KzStackRef stk;
        uint64_t nodes = 0, tris = 0, groups = 0, hits = 0, maxsp = 0;
        for (size_t i = b; i < en; ++i) {
            const kz_ray &r = rays[i];
            KzTrav t; KzLocalStack ls;
            kz_trav_init(sc, t, r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], r.tmin, r.tmax);
            if (sc.n_nodes == 0) continue;
            for (;;) {
                if (t.ng_y > 0x00FFFFFFu) { kz_trav_node(sc, t, stk, ls); ++nodes; }
                else { t.tg_x = t.ng_x; t.tg_y = t.ng_y; t.ng_x = 0u; t.ng_y = 0u; }
                if (t.tg_y) ++groups;
                while (t.tg_y != 0u) { const float before = t.best.t; const uint32_t bp = t.best.prim; kz_trav_tri(sc, t); ++tris; if (t.best.t != before || t.best.prim != bp) ++hits; }
                if ((uint64_t)t.sp > maxsp) maxsp = t.sp;
                if (t.ng_y <= 0x00FFFFFFu) { if (t.sp == 0) break; kz_trav_pop(t, stk, ls); }
            }
        }
        a_nodes += nodes; a_tris += tris; a_groups += groups; a_hits += hits;
        uint64_t cur = a_sp.load(); while (maxsp > cur && !a_sp.compare_exchange_weak(cur, maxsp)) {}
    });
    out[0] = a_nodes; out[1] = a_tris; out[2] = a_groups; out[3] = a_hits; out[4] = a_sp;
    return KZ_OK;
}

int kzemu_occluded(kzemu *e, const kz_ray *rays, size_t n, float trace_bias, uint8_t *occ, uint8_t *segments) {
    const KzScene &sc = e->hs.sc;
    pfor(n, [&](size_t b, size_t en) {
        KzStackRef stk;
        for (size_t i = b; i < en; ++i) {
            const kz_ray &r = rays[i];
            int seg;
            occ[i] = kz_occluded_walk(sc, stk, mk3(r.o[0], r.o[1], r.o[2]), mk3(r.d[0], r.d[1], r.d[2]), r.tmin, r.tmax, trace_bias, &seg) ? 1 : 0;
            if (segments) segments[i] = (uint8_t)std::min(seg, 255);
        }
    });
    return KZ_OK;
}

/* The shadow walk with the early stop of the device's shadow / occlusion jobs (KzWalk::settles in kz_kernels.cuh, restated): a segment
 * ends at the first OPAQUE hit the traversal comes across once no invisible emitter can lie in front of it.  Must answer exactly as
 * kzemu_occluded (the closest-hit walk of integrator.cpp:259-294): same flags, same segment counts.  stops[0] counts the early stops. */
int kzemu_occluded_early(kzemu *e, const kz_ray *rays, size_t n, float trace_bias, uint8_t *occ, uint8_t *segments, uint64_t *stops) {
    const KzScene &sc = e->hs.sc;
    std::atomic<uint64_t> n_stop{0};
    pfor(n, [&](size_t b, size_t en) {
        KzStackRef stk; uint64_t my_stops = 0;
        for (size_t i = b; i < en; ++i) {
            const kz_ray &r = rays[i];
            kz3 o = mk3(r.o[0], r.o[1], r.o[2]); const kz3 d = mk3(r.d[0], r.d[1], r.d[2]);
            float tmin = r.tmin, tmax = r.tmax;
            int seg = 0; bool occluded = false;
            for (;;) {
                ++seg;
                KzTrav t; KzLocalStack ls;
                kz_trav_init(sc, t, o.x, o.y, o.z, d.x, d.y, d.z, tmin, tmax);
                bool stopped = false;
                while (!stopped && sc.n_nodes) {
                    if (t.ng_y > 0x00FFFFFFu) kz_trav_node(sc, t, stk, ls);
                    while (t.tg_y != 0u) {
                        if (!kz_trav_tri(sc, t)) continue;
                        bool settles = sc.integrator.type != KZ_INTEGRATOR_PATH_MIS;
                        if (!settles) {
                            const uint32_t fl = sc.meshes[t.best.geom].flags;
                            if (!((fl & KZ_MESH_IS_LIGHT) && !(fl & KZ_MESH_LIGHT_VISIBLE)))
                                settles = sc.n_invisible_lights == 0 || kz_trav_misses_box(t, sc.inv_light_lo, sc.inv_light_hi, t.best.t);
                        }
                        if (settles) { stopped = true; ++my_stops; break; }
                    }
                    if (stopped) break;
                    if (t.ng_y <= 0x00FFFFFFu) { if (t.sp == 0) break; kz_trav_pop(t, stk, ls); }
                }
                const KzHit h = t.best;
                if (h.geom == KZ_INVALID_ID) break;
                const uint32_t fl = sc.meshes[h.geom].flags;
                if (!(fl & KZ_MESH_IS_LIGHT) || (fl & KZ_MESH_LIGHT_VISIBLE) || sc.integrator.type != KZ_INTEGRATOR_PATH_MIS) { occluded = true; break; }
                o = o + d * (h.t + trace_bias); tmin = trace_bias; tmax = tmax - h.t;
                if (seg > 4096) break;
            }
            occ[i] = occluded ? 1 : 0;
            if (segments) segments[i] = (uint8_t)std::min(seg, 255);
        }
        n_stop += my_stops;
    });
    if (stops) stops[0] = n_stop.load();
    return KZ_OK;
}

int kzemu_sample_dump(kzemu *e, const int32_t *triples, size_t n, const char *pattern, float *out) {
    const KzScene &sc = e->hs.sc;
    size_t per = 0;
    for (const char *p = pattern; *p; ++p) per += (*p == '1') ? 1 : 2;
    for (size_t i = 0; i < n; ++i) {
        KzSampler sm;
        kz_sampler_start(sc, sm, triples[3 * i], triples[3 * i + 1], (uint32_t)triples[3 * i + 2]);
        float *o = out + i * per;
        for (const char *p = pattern; *p; ++p) {
            if (*p == '1') *o++ = kz_next1d(sc, sm);
            else { kz2 v = (*p == 'P') ? kz_next_pixel2d(sc, sm) : kz_next2d(sc, sm); *o++ = v.x; *o++ = v.y; }
        }
    }
    return KZ_OK;
}

int kzemu_camera_rays(kzemu *e, const float *s4, size_t n, kz_ray *out) {
    for (size_t i = 0; i < n; ++i) {
        KzF4 ro, rd;
        kz_camera_ray(e->hs.sc.camera, mk2(s4[4 * i], s4[4 * i + 1]), mk2(s4[4 * i + 2], s4[4 * i + 3]), ro, rd);
        out[i] = kz_ray{{ro.x, ro.y, ro.z}, ro.w, {rd.x, rd.y, rd.z}, rd.w};
    }
    return KZ_OK;
}

int kzemu_bsdf_query(kzemu *e, int mesh_bsdf, int mode, const float wi[3], const float wo[3], const float uv[2], float accR,
                     float sample1, const float sample2[2], float out[8]) {
    /* builds a fake intersection with an identity shading frame on a mesh slot that uses `mesh_bsdf` */
    const KzScene &sc = e->hs.sc;
    int mesh = -1;
    for (uint32_t g = 0; g < sc.n_meshes; ++g) if (sc.meshes[g].bsdf == mesh_bsdf) { mesh = (int)g; break; }
    if (mesh < 0) return KZ_ERR_INVALID;
    KzIts its;
    its.sh.s = mk3(1, 0, 0); its.sh.t = mk3(0, 1, 0); its.sh.n = mk3(0, 0, 1);
    its.geo_n = its.sh.n; its.dpdu = mk3(1, 0, 0); its.uv = mk2(uv[0], uv[1]); its.mesh = mesh; its.acc_rough = accR;
    its.p = mk3(0.f);
    KzBsdfCtx bc = bsdf_ctx(sc, its);
    for (int i = 0; i < 8; ++i) out[i] = 0.f;
    if (mode == 2) {
        kz3 w_o; float pdf, eta_s; int measure;
        kz3 w = bsdf_sample(bc, its, mk3(wi[0], wi[1], wi[2]), sample1, mk2(sample2[0], sample2[1]), &w_o, &pdf, &measure, &eta_s);
        out[0] = w.x; out[1] = w.y; out[2] = w.z;
        if (!iszero(w)) { out[3] = w_o.x; out[4] = w_o.y; out[5] = w_o.z; out[7] = pdf; }
        out[6] = (float)measure;
        return KZ_OK;
    }
    kz3 f; float pdf;
    bsdf_eval_pdf(bc, its, mk3(wi[0], wi[1], wi[2]), mk3(wo[0], wo[1], wo[2]), &f, &pdf);
    if (mode == 0) { out[0] = f.x; out[1] = f.y; out[2] = f.z; } else out[0] = pdf;
    return KZ_OK;
}

/* Batch-synchronous wavefront over all requested paths, stage functions called item by item. */
int kzemu_render(kzemu *e, const kz_render_req *req, float *frame_rgbw) {
    const KzScene &sc = e->hs.sc;
    const int b = sc.border, cols = sc.camera.width + 2 * b, rows = sc.camera.height + 2 * b;
    const size_t fsz = (size_t)cols * rows;
    KzF4 *frame = reinterpret_cast<KzF4 *>(frame_rgbw);
    if (req->clear_frame) memset(frame, 0, fsz * sizeof(KzF4));
    const int w = req->x1 - req->x0, hgt = req->y1 - req->y0, nS = req->spp_end - req->spp_begin;
    const size_t B = (size_t)w * hgt * nS;
    std::vector<KzBlockA> sa(B); std::vector<KzBlockB> sb(B); std::vector<KzBlockC> scv(B);
    KzPathState st{sa.data(), sb.data(), scv.data()};
    std::vector<uint32_t> q(B), qn, qs;
    size_t slot = 0;
    for (int y = req->y0; y < req->y1; ++y)
        for (int x = req->x0; x < req->x1; ++x)
            for (int s = req->spp_begin; s < req->spp_end; ++s) {
                kz_raygen_item(sc, st, (uint32_t)slot, x, y, (uint32_t)s);
                q[slot] = (uint32_t)slot;
                ++slot;
            }
    std::mutex mu;
    KzCounters tot{B, 0, 0, 0};
    const int max_pass = sc.integrator.type == KZ_INTEGRATOR_PATH_MIS ? sc.integrator.max_depth : (sc.integrator.type <= KZ_INTEGRATOR_AO ? 0 : 4095);
    for (int bounce = 0; bounce <= max_pass && !q.empty(); ++bounce) {
        qn.clear(); qs.clear();
        pfor(q.size(), [&](size_t bb, size_t ee) {
            KzStackRef stk; KzCounters c{0, 0, 0, 0};
            std::vector<uint32_t> ln, ls;
            for (size_t i = bb; i < ee; ++i) {
                kz_extend_item(sc, stk, st, q[i], bounce, c);
                uint32_t fl = kz_shade_item(sc, st, q[i], bounce, c);
                if (fl & KZ_SHADE_CONTINUE) ln.push_back(q[i]);
                if (fl & KZ_SHADE_SHADOW) ls.push_back(q[i]);
            }
            for (uint32_t s : ls) kz_shadow_item(sc, stk, st, s, c);
            std::lock_guard<std::mutex> lk(mu);
            qn.insert(qn.end(), ln.begin(), ln.end());
            tot.rays_ext += c.rays_ext; tot.rays_shadow += c.rays_shadow; tot.vertices += c.vertices;
        });
        q.swap(qn);
    }
    for (size_t i = 0; i < B; ++i) kz_accumulate_item(sc, st, (uint32_t)i, frame, sc.filter.table);
    e->total.paths += tot.paths; e->total.rays_ext += tot.rays_ext; e->total.rays_shadow += tot.rays_shadow; e->total.vertices += tot.vertices;
    return KZ_OK;
}

/* The film accumulation of the device, restated lane by lane: k_raygen's path order (kz_kernels.cuh: blocks of `spp_group` sample
 * indices, 8x4-pixel tiles, 32-path units), every "warp" a run of `units_per_warp` units, the taps of a tile summed in a region by their
 * offset from the path's own pixel and handed to the frame when the tile changes (k_accumulate).  Every path carries the radiance
 * L = hash(pixel, sample) instead of a traced one.  Fills `direct` (kz_accumulate_item per path) and `tiled`; returns the number of
 * (lane, offset) steps in which two lanes of a unit touched the same region texel -- the scheme is race free iff that is 0. */
int kzemu_splat_orders(kzemu *e, const kz_render_req *req, int spp_group, int units_per_warp, float *direct_rgbw, float *tiled_rgbw, uint64_t *conflicts) {
    const KzScene &sc = e->hs.sc;
    const int b = sc.border, cols = sc.camera.width + 2 * b, rows = sc.camera.height + 2 * b;
    KzF4 *fd = reinterpret_cast<KzF4 *>(direct_rgbw), *ft = reinterpret_cast<KzF4 *>(tiled_rgbw);
    memset(fd, 0, (size_t)cols * rows * sizeof(KzF4)); memset(ft, 0, (size_t)cols * rows * sizeof(KzF4));
    const int w = req->x1 - req->x0, h = req->y1 - req->y0, nS = req->spp_end - req->spp_begin;
    const uint32_t tiles_x = (uint32_t)(w + 7) / 8u, tiles_y = (uint32_t)(h + 3) / 4u, npx = tiles_x * tiles_y * 32u;
    const uint64_t total = (uint64_t)npx * nS;
    std::vector<KzBlockA> sa(total); std::vector<KzBlockB> sb(total); std::vector<KzBlockC> scv(total);
    KzPathState st{sa.data(), sb.data(), scv.data()};
    const uint32_t G = (uint32_t)std::min(nS, std::max(1, spp_group));
    for (uint64_t i = 0; i < total; ++i) {                       /* k_raygen's decode of the path index */
        uint64_t g = i;
        const uint64_t per_block = (uint64_t)npx * G;
        const uint32_t blk = (uint32_t)(g / per_block); g -= (uint64_t)blk * per_block;
        const uint32_t left = (uint32_t)nS - blk * G, grp = left < G ? left : G;
        const uint32_t tile = (uint32_t)(g / (32u * grp)), r = (uint32_t)(g - (uint64_t)tile * (32u * grp));
        const uint32_t s_local = blk * G + (r >> 5), in_tile = r & 31u;
        const int x = req->x0 + (int)((tile % tiles_x) * 8u + (in_tile & 7u)), y = req->y0 + (int)((tile / tiles_x) * 4u + (in_tile >> 3));
        if (x < req->x1 && y < req->y1) {
            kz_raygen_item(sc, st, (uint32_t)i, x, y, (uint32_t)(req->spp_begin + (int)s_local));
            const uint32_t hsh = (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (s_local + 1u) * 83492791u;
            st.b[i].rad.L = mkf4((float)(hsh & 255u) / 64.f, (float)((hsh >> 8) & 255u) / 64.f, (float)((hsh >> 16) & 255u) / 64.f, 1.f);
            if ((hsh >> 24) == 7u) st.b[i].rad.L.x = -1.f;          /* an invalid value now and then (block.cpp:58-62) */
        } else st.b[i].smp.pix = 0xFFFFFFFFu;
    }
    for (uint64_t i = 0; i < total; ++i) if (st.b[i].smp.pix != 0xFFFFFFFFu) kz_accumulate_item(sc, st, (uint32_t)i, fd, sc.filter.table);
    /* ---- k_accumulate ---- */
    const float radius = sc.filter.radius, lookup = 32 / radius;
    const int lo = -(int)floorf(radius + 0.5f), hi = (int)ceilf(radius + 0.5f) - 1, tw = 8 + hi - lo, th = 4 + hi - lo;
    std::vector<KzF4> region((size_t)tw * th, KzF4{0.f, 0.f, 0.f, 0.f});
    std::vector<int> touched((size_t)tw * th);
    uint64_t bad = 0;
    const uint64_t n_units = (total + 31u) >> 5;
    auto flush = [&](int cx, int cy) {
        for (int k = 0; k < tw * th; ++k) {
            const KzF4 v = region[k];
            const int X = cx + b + lo + k % tw, Y = cy + b + lo + k / tw;
            if ((v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) && X >= 0 && X < cols && Y >= 0 && Y < rows) { KzF4 &p = ft[(size_t)Y * cols + X]; p.x += v.x; p.y += v.y; p.z += v.z; p.w += v.w; }
            region[k] = KzF4{0.f, 0.f, 0.f, 0.f};
        }
    };
    for (uint64_t u0 = 0; u0 < n_units; u0 += (uint64_t)units_per_warp) {            /* one warp's run */
        bool open = false; int cur_x = 0, cur_y = 0;
        for (uint64_t u = u0; u < std::min<uint64_t>(n_units, u0 + units_per_warp); ++u) {
            KzSplat sp[32]; bool ok[32]; int leader = -1;
            for (int l = 0; l < 32; ++l) {
                const uint64_t i = (u << 5) + l;
                ok[l] = i < total && st.b[i].smp.pix != 0xFFFFFFFFu && kz_splat_of(sc, st, (uint32_t)i, sp[l]);
                if (ok[l] && (sp[l].x0 < sp[l].ipx + b + lo || sp[l].x1 > sp[l].ipx + b + hi || sp[l].y0 < sp[l].ipy + b + lo || sp[l].y1 > sp[l].ipy + b + hi)) {
                    kz_accumulate_item(sc, st, (uint32_t)i, ft, sc.filter.table); ok[l] = false;
                }
                if (ok[l] && leader < 0) leader = l;
            }
            if (leader < 0) continue;
            const int tx = req->x0 + ((sp[leader].ipx - req->x0) & ~7), ty = req->y0 + ((sp[leader].ipy - req->y0) & ~3);
            if (open && (tx != cur_x || ty != cur_y)) flush(cur_x, cur_y);
            cur_x = tx; cur_y = ty; open = true;
            for (int dy = lo; dy <= hi; ++dy)
                for (int dx = lo; dx <= hi; ++dx) {
                    std::fill(touched.begin(), touched.end(), 0);
                    for (int l = 0; l < 32; ++l) {
                        if (!ok[l]) continue;
                        const int X = sp[l].ipx + b + dx, Y = sp[l].ipy + b + dy;
                        if (Y < sp[l].y0 || Y > sp[l].y1 || X < sp[l].x0 || X > sp[l].x1) continue;
                        const float wy = sc.filter.table[(int)(fabsf((float)Y - sp[l].py) * lookup)], wx = sc.filter.table[(int)(fabsf((float)X - sp[l].px) * lookup)];
                        const int ry = sp[l].ipy - ty - lo + dy, rx = sp[l].ipx - tx - lo + dx;
                        if (ry < 0 || ry >= th || rx < 0 || rx >= tw) { ++bad; continue; }
                        if (touched[(size_t)ry * tw + rx]++) ++bad;
                        KzF4 &p = region[(size_t)ry * tw + rx];
                        p.x += sp[l].value.x * wx * wy; p.y += sp[l].value.y * wx * wy; p.z += sp[l].value.z * wx * wy; p.w += 1.0f * wx * wy;
                    }
                }
        }
        if (open) flush(cur_x, cur_y);
    }
    *conflicts = bad;
    return KZ_OK;
}

/* the per-item bodies of k_intersection_dump / k_light_sample_dump (kz_kernels.cuh) */
int kzemu_intersection_dump(kzemu *e, const kz_ray *rays, size_t n, float *out24) {
    const KzScene &sc = e->hs.sc;
    pfor(n, [&](size_t b, size_t en) {
        KzStackRef stk;
        for (size_t i = b; i < en; ++i) {
            const kz_ray &r = rays[i];
            const KzHit h = kz_trace(sc, stk, r.o[0], r.o[1], r.o[2], r.d[0], r.d[1], r.d[2], r.tmin, r.tmax, false);
            float *o = out24 + 24 * i;
            for (int k = 0; k < 24; ++k) o[k] = 0.f;
            o[0] = h.t; o[1] = -1.f;
            if (h.geom == KZ_INVALID_ID) continue;
            KzIts its; its.acc_rough = 0.f;
            fill_intersection(sc, h, its, mk3(0.f));
            o[1] = (float)its.mesh;
            o[2] = its.p.x; o[3] = its.p.y; o[4] = its.p.z; o[5] = its.uv.x; o[6] = its.uv.y;
            o[7] = its.geo_n.x; o[8] = its.geo_n.y; o[9] = its.geo_n.z;
            o[10] = its.sh.s.x; o[11] = its.sh.s.y; o[12] = its.sh.s.z;
            o[13] = its.sh.t.x; o[14] = its.sh.t.y; o[15] = its.sh.t.z;
            o[16] = its.sh.n.x; o[17] = its.sh.n.y; o[18] = its.sh.n.z;
            o[19] = its.dpdu.x; o[20] = its.dpdu.y; o[21] = its.dpdu.z;
        }
    });
    return KZ_OK;
}
int kzemu_light_sample_dump(kzemu *e, const float *ref3, const float *u5, size_t n, float *out16) {
    const KzScene &sc = e->hs.sc;
    for (size_t i = 0; i < n; ++i) {
        float *o = out16 + 16 * i;
        for (int k = 0; k < 16; ++k) o[k] = 0.f;
        o[0] = -1.f;
        if (sc.n_light_meshes <= 0) continue;
        const float *u = u5 + 5 * i;
        const KzEmitterSample es = kz_sample_emitter(sc, mk3(ref3[3 * i], ref3[3 * i + 1], ref3[3 * i + 2]), u[0], u[1], u[2], u[3]);
        o[0] = (float)es.mesh;
        o[1] = es.p.x; o[2] = es.p.y; o[3] = es.p.z; o[4] = es.n.x; o[5] = es.n.y; o[6] = es.n.z;
        o[7] = es.wi.x; o[8] = es.wi.y; o[9] = es.wi.z; o[10] = es.dist; o[11] = es.pdf;
        if (es.pdf > 0.f && !isnan(es.pdf) && !isinf(es.pdf)) {
            const kz_light_desc l = sc.lights[es.light];
            const kz3 Ls = mk3(l.radiance[0], l.radiance[1], l.radiance[2]) / es.pdf;
            o[12] = Ls.x; o[13] = Ls.y; o[14] = Ls.z;
        }
    }
    return KZ_OK;
}

int kzemu_stats(kzemu *e, kz_stats *out) {
    memset(out, 0, sizeof(*out));
    out->paths = e->total.paths; out->rays_extension = e->total.rays_ext; out->rays_shadow = e->total.rays_shadow; out->vertices = e->total.vertices;
    out->bvh_nodes = e->bvh.nodes.size(); out->bvh_bytes = e->bvh.nodes.size() * 80 + e->bvh.tris.size() * 16;
    return KZ_OK;
}

}  // extern "C"
