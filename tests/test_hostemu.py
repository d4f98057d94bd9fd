"""The device routines of nano-kazen_b200/csrc compiled for the host (tests/hostemu) against the
oracle: this is what can be verified about the CUDA source on a box without a GPU.  The GPU suite
repeats the same comparisons through the C ABI on the real device."""
import numpy as np
import pytest

import scenes


def _pair(kzo, emu, sb):
    d = sb.desc()
    return kzo.Oracle(d), emu.Emu(d)


def test_trace_bit_exact(kzo, emu):
    sb = scenes.cornell_scene(32, 32, 4)
    O, E = _pair(kzo, emu, sb)
    rays = np.concatenate([scenes.primary_rays(64, 39.0, (0, 0, -3.4)), scenes.incoherent_rays(20000, extent=0.95)])
    a, b = O.trace(rays, brute=True), E.trace(rays)
    assert a.tobytes() == b.tobytes()
    O.close(); E.close()


@pytest.mark.parametrize("n", [1, 2, 7, 5000, 60000])
def test_soup_trace_bit_exact(kzo, emu, n):
    sb = scenes.soup_scene(n)
    O, E = _pair(kzo, emu, sb)
    rays = np.concatenate([scenes.primary_rays(64), scenes.incoherent_rays(8000)])
    a, b = O.trace(rays, brute=(n <= 5000)), E.trace(rays)
    assert a.tobytes() == b.tobytes()
    nodes, tris, depth = E.bvh_info()
    assert n <= tris <= 2 * n and nodes >= 1          # stored references: the SAH builder may split a triangle in two (KZ_SAH_PRESPLIT)
    O.close(); E.close()


def test_occluded_walk(kzo, emu):
    sb = scenes.cornell_scene(16, 16, 4)
    O, E = _pair(kzo, emu, sb)
    rays = scenes.incoherent_rays(20000, extent=0.97)
    rays["tmax"] *= 1.5
    (oa, sa), (ob, sb_) = O.occluded(rays, 1e-3), E.occluded(rays, 1e-3)
    assert np.array_equal(oa, ob) and np.array_equal(sa, sb_)
    assert sa.max() >= 2          # some rays stepped through an invisible light
    O.close(); E.close()


@pytest.mark.parametrize("kind", ["independent", "stratified", "correlated"])
def test_sampler_bit_exact(kzo, emu, kind):
    sb = scenes.cornell_scene(64, 64, 30, kind)
    O, E = _pair(kzo, emu, sb)
    rng = np.random.default_rng(3)
    tr = np.stack([rng.integers(0, 4000, 500), rng.integers(0, 4000, 500), rng.integers(0, sb.sampler.sample_count, 500)], 1).astype(np.int32)
    pat = "P2" + "1111" * 2 + "12" * 3 + "1"
    a, b = O.sample_dump(tr, pat), E.sample_dump(tr, pat)
    assert a.tobytes() == b.tobytes()
    O.close(); E.close()


def test_pmj02bn_with_synthetic_tables(kzo, emu):
    """the pbrt tables are missing from the reference mount (parity unpinned); the code path is
    checked oracle-vs-device with seeded stand-in tables"""
    rng = np.random.default_rng(5)
    bn = rng.integers(0, 65536, (48, 128, 128), dtype=np.uint16)
    pm = rng.integers(0, 2 ** 32, (5, 65536, 2), dtype=np.uint32)
    sb = scenes.cornell_scene(16, 16, 16, "stratified")
    sb.set_sampler("pmj02bn", 16, tables=(bn, pm))
    O, E = _pair(kzo, emu, sb)
    tr = np.array([[x, y, j] for x in (0, 5, 130) for y in (1, 77) for j in (0, 3, 15)], np.int32)
    a, b = O.sample_dump(tr, "P2121212121212"), E.sample_dump(tr, "P2121212121212")
    assert a.tobytes() == b.tobytes()
    assert (a >= 0).all() and (a < 1).all()
    O.close(); E.close()


@pytest.mark.parametrize("thin", [None, (0.05, 3.0)])
def test_camera_rays(kzo, emu, thin):
    sb = scenes.cornell_scene(64, 48, 4, thinlens=thin)
    O, E = _pair(kzo, emu, sb)
    rng = np.random.default_rng(2)
    s4 = rng.uniform(0, 1, (500, 4)).astype(np.float32) * np.array([64, 48, 1, 1], np.float32)
    a, b = O.camera_rays(s4), E.camera_rays(s4)
    for k in ("o", "d", "tmin", "tmax"):
        assert np.allclose(a[k], b[k], rtol=2e-6, atol=1e-7)
    O.close(); E.close()


def test_bsdf_queries(kzo, emu):
    sb = scenes.cornell_scene(8, 8, 1, with_texture=True, normalmap=True)
    O, E = _pair(kzo, emu, sb)
    rng = np.random.default_rng(9)
    for bsdf in range(len(sb.bsdfs)):
        if not any(m.bsdf == bsdf for m in sb.meshes):
            continue
        for _ in range(40):
            wi = rng.normal(size=3); wi[2] = abs(wi[2]) + 0.05; wi /= np.linalg.norm(wi)
            wo = rng.normal(size=3); wo[2] = abs(wo[2]) * rng.choice([1, 1, 1, -1]) + 0.01; wo /= np.linalg.norm(wo)
            uv = rng.uniform(0, 1, 2); acc = float(rng.choice([0.0, 0.3]))
            s1 = float(rng.uniform()); s2 = rng.uniform(0, 1, 2)
            for mode in (0, 1, 2):
                a = O.bsdf_query(bsdf, mode, wi, wo, uv, acc, s1, s2)
                b = E.bsdf_query(bsdf, mode, wi, wo, uv, acc, s1, s2)
                n = 3 if mode == 0 else (1 if mode == 1 else 6)
                assert np.allclose(a[:n], b[:n], rtol=2e-4, atol=1e-6), (bsdf, mode, a, b)
    O.close(); E.close()


@pytest.mark.parametrize("cfg", [
    dict(sampler="stratified"),
    dict(sampler="correlated", visible_light=True),
    dict(sampler="independent", with_texture=True, normalmap=True, regularization=True),
    dict(sampler="stratified", thinlens=(0.05, 3.2), background=(0.3, 0.4, 0.5), max_depth=3),
])
def test_render_matches_oracle(kzo, emu, cfg):
    """same samples, same arithmetic up to libm/FMA: images agree far below Monte Carlo noise"""
    sb = scenes.cornell_scene(32, 32, 16, **cfg)
    O, E = _pair(kzo, emu, sb)
    fo, fe = O.render(), E.render()
    ro, _ = O.resolve(fo); re, _ = O.resolve(fe)
    assert np.allclose(fo[..., 3], fe[..., 3], rtol=1e-5, atol=1e-6)        # identical splat weights
    assert scenes.rel_mse(re, ro).max() < 1e-6
    so, se = O.stats(), E.stats()
    assert so["paths"] == se["paths"] and so["rays_extension"] == se["rays_extension"] and so["vertices"] == se["vertices"]
    O.close(); E.close()


def test_render_is_shardable(kzo, emu):
    """sum of disjoint sample-index shards / rectangles == the whole render"""
    sb = scenes.cornell_scene(24, 24, 16, "stratified")
    O, E = _pair(kzo, emu, sb)
    whole = E.render()
    parts = E.render(0, 5) + E.render(5, 16)
    assert np.allclose(whole, parts, rtol=1e-5, atol=1e-6)
    tiles = E.render(rect=(0, 0, 10, 24)) + E.render(rect=(10, 0, 24, 11)) + E.render(rect=(10, 11, 24, 24))
    assert np.allclose(whole, tiles, rtol=1e-5, atol=1e-6)
    O.close(); E.close()


def test_extra_bsdf_queries(kzo, emu):
    """dielectric / mirror / lambertian / ggx / roughconductor / roughplastic / roughdielectric + normal-mapped
    roughplastic, eval / pdf / sample, device routines vs oracle"""
    sb = scenes.gallery_scene(8, 8, 1)
    O, E = _pair(kzo, emu, sb)
    rng = np.random.default_rng(13)
    checked = 0
    for bsdf in range(len(sb.bsdfs)):
        if not any(m.bsdf == bsdf for m in sb.meshes):
            continue
        for _ in range(60):
            wi = rng.normal(size=3); wi[2] = (abs(wi[2]) + 0.05) * rng.choice([1, 1, 1, -1]); wi /= np.linalg.norm(wi)
            wo = rng.normal(size=3); wo[2] = (abs(wo[2]) + 0.02) * rng.choice([1, 1, -1]); wo /= np.linalg.norm(wo)
            uv = rng.uniform(0, 1, 2); s1 = float(rng.uniform()); s2 = rng.uniform(0.001, 0.999, 2)
            for mode in (0, 1, 2):
                a = O.bsdf_query(bsdf, mode, wi, wo, uv, 0.0, s1, s2)
                b = E.bsdf_query(bsdf, mode, wi, wo, uv, 0.0, s1, s2)
                n = 3 if mode == 0 else (1 if mode == 1 else 7)
                assert np.allclose(a[:n], b[:n], rtol=5e-4, atol=2e-6), (sb.bsdfs[bsdf].type, mode, a, b)
                checked += 1
    assert checked > 1000
    O.close(); E.close()


def test_extra_bsdf_render_matches_oracle(kzo, emu):
    sb = scenes.gallery_scene(40, 30, 16)
    O, E = _pair(kzo, emu, sb)
    fo, fe = O.render(), E.render()
    ro, _ = O.resolve(fo); re, _ = O.resolve(fe)
    assert np.allclose(fo[..., 3], fe[..., 3], rtol=1e-5, atol=1e-6)
    assert scenes.rel_mse(re, ro).max() < 1e-5
    assert ro.mean() > 0.02
    so, se = O.stats(), E.stats()
    assert so["paths"] == se["paths"] and abs(so["vertices"] - se["vertices"]) <= 2e-4 * so["vertices"]
    O.close(); E.close()


@pytest.mark.parametrize("kind", ["normals", "ao", "whitted", "path_mats"])
def test_other_integrators_match_oracle(kzo, emu, kind):
    """the other integrator plugins through the same wavefront"""
    sb = scenes.cornell_scene(32, 32, 16, "stratified", visible_light=True)
    sb.set_integrator(kind=kind)
    O, E = _pair(kzo, emu, sb)
    fo, fe = O.render(), E.render()
    ro, _ = O.resolve(fo); re, _ = O.resolve(fe)
    assert np.allclose(fo[..., 3], fe[..., 3], rtol=1e-5, atol=1e-6)
    # whitted: the reference's shadow ray spans [0, dist] exactly (light.cpp:24), so it grazes the shading point and the light
    # sample itself; whether those end points count as hits flips with the last bit of its.p / dist -> statistical parity only
    assert scenes.rel_mse(re, ro).max() < (5e-3 if kind == "whitted" else 1e-6)
    assert ro.mean() > (0.005 if kind == "whitted" else 0.05)
    O.close(); E.close()


def test_exact_ties_resolve_by_ids_cpu(kzo, emu):
    from test_gpu_parity import _tie_scene
    sb = _tie_scene()
    O, E = _pair(kzo, emu, sb)
    rays = scenes.primary_rays(64, 40.0, (0, 0, -3.0))
    a = O.trace(rays, brute=True)
    assert a.tobytes() == O.trace(rays, brute=False).tobytes() == E.trace(rays).tobytes()
    hit = a["geom_id"] != 0xFFFFFFFF
    assert hit.any() and (a["geom_id"][hit] == 0).all()
    O.close(); E.close()


def test_latlong_environment_map_cpu(kzo, emu):
    rng = np.random.default_rng(17)
    env = rng.uniform(0.0, 2.0, (16, 32, 3)).astype(np.float32)
    sb = scenes.cornell_scene(32, 24, 16, "stratified", max_depth=4)
    sb.background = sb.tex_background(1.5, sb.tex_image(env, srgb=False))
    sb.meshes = [m for i, m in enumerate(sb.meshes) if i != 1 and m.light < 0]
    O, E = _pair(kzo, emu, sb)
    ro, _ = O.resolve(O.render()); re, _ = O.resolve(E.render())
    assert ro.mean() > 0.05 and scenes.rel_mse(re, ro).max() < 1e-6
    O.close(); E.close()


def _probe_rays(n_primary=48, n_incoherent=6000):
    return np.concatenate([scenes.primary_rays(n_primary, 39.0, (0, 0, -3.4)), scenes.incoherent_rays(n_incoherent, extent=0.95)])


def check_intersection_dump(O, X, rays, rtol=2e-5):
    """A3 field by field (accel.cpp:113-236): t and the mesh bit-exact, positions / uv / frames / dpdu to rounding."""
    a, b = O.intersection_dump(rays), X.intersection_dump(rays)
    assert a[:, :2].tobytes() == b[:, :2].tobytes()
    hit = a[:, 1] >= 0
    assert hit.sum() > rays.shape[0] // 4
    np.testing.assert_allclose(b[hit, 2:7], a[hit, 2:7], rtol=rtol, atol=2e-6)          # p, uv
    np.testing.assert_allclose(b[hit, 7:19], a[hit, 7:19], rtol=0, atol=3e-5)            # unit vectors of the two frames
    scale = np.abs(a[hit, 19:22]).max(axis=1, keepdims=True) + 1e-6
    np.testing.assert_allclose(b[hit, 19:22] / scale, a[hit, 19:22] / scale, rtol=0, atol=1e-4)   # dpdu (any magnitude)


def check_light_sample_dump(O, X, n=4000, seed=5, rtol=2e-5):
    """L1 / L2 / S1 field by field (scene.h:45-56, mesh.cpp:108-133, light.cpp:16-51) on shared random numbers."""
    rng = np.random.default_rng(seed)
    ref = rng.uniform(-0.9, 0.9, (n, 3)).astype(np.float32)
    u5 = rng.uniform(0, 1, (n, 5)).astype(np.float32)
    u5[:8, 0] = [0.0, 0.49999997, 0.5, 0.99999994, 0.25, 0.75, 0.0, 0.5]            # light-pick edges
    u5[:4, 1] = [0.0, 0.99999994, 0.5, 0.50000006]                                  # CDF edges
    a, b = O.light_sample_dump(ref, u5), X.light_sample_dump(ref, u5)
    assert a[:, 0].tobytes() == b[:, 0].tobytes()                                    # same emitter picked
    assert (a[:, 0] >= 0).all() and len(np.unique(a[:, 0])) >= 2
    np.testing.assert_allclose(b[:, 1:11], a[:, 1:11], rtol=rtol, atol=2e-6)         # p, n, wi, dist
    usable = a[:, 11] > 0
    assert usable.sum() > n // 4 and ((b[:, 11] > 0) == usable).all()
    np.testing.assert_allclose(b[usable, 11:15], a[usable, 11:15], rtol=1e-4)        # pdf, Le / pdf


@pytest.mark.parametrize("kw", [dict(), dict(with_texture=True, normalmap=True)])
def test_intersection_record_cpu(kzo, emu, kw):
    O, E = _pair(kzo, emu, scenes.cornell_scene(16, 16, 4, **kw))
    check_intersection_dump(O, E, _probe_rays())
    O.close(); E.close()


def test_intersection_record_gallery_cpu(kzo, emu):
    """meshes with uvs + normals, normals only, and neither (all three shading-frame cases of accel.cpp:190-236)"""
    O, E = _pair(kzo, emu, scenes.gallery_scene(16, 12, 4))
    rays = np.concatenate([scenes.primary_rays(64, 50.0, (0, 0.1, -3.2)), scenes.incoherent_rays(6000, extent=0.9)])
    check_intersection_dump(O, E, rays)
    O.close(); E.close()


def test_emitter_sample_cpu(kzo, emu):
    O, E = _pair(kzo, emu, scenes.cornell_scene(16, 16, 4))
    check_light_sample_dump(O, E)
    O.close(); E.close()


@pytest.mark.parametrize("kind,kw", [("gaussian", {}), ("mitchell", {}), ("tent", {}), ("box", {}), ("gaussian", {"radius": 2.5}), ("gaussian", {"radius": 3.0})])
def test_tiled_film_accumulation_cpu(emu, kind, kw):
    """k_accumulate's order of summation (kz_kernels.cuh), restated lane by lane in tests/hostemu: the taps of a pixel tile's sample
    indices summed in a region by their offset from the path's own pixel, the region added to the frame when the tile changes.  For
    every reconstruction filter, aligned and unaligned request rectangles, runs of sample indices that do and do not divide the
    sample count and warps whose runs split tiles: the frame equals the path-by-path splats up to the order of the float additions,
    and no two lanes of a unit ever meet in one texel at one offset (the reason it needs no atomics)."""
    sb = scenes.cornell_scene(61, 37, 6, "stratified")
    sb.set_filter(kind, **kw)
    E = emu.Emu(sb.desc())
    for rect in ((0, 0, 61, 37), (3, 2, 58, 31), (17, 5, 26, 9)):
        for spp, group, run in ((6, 64, 53), (6, 4, 3), (5, 2, 1), (1, 1, 7)):
            fd, ft, bad = E.splat_orders(rect, spp, group, run)
            assert bad == 0
            assert fd[..., 3].sum() > 0
            scale = np.abs(fd).max()
            assert np.abs(fd - ft).max() <= 2e-5 * scale, (kind, rect, spp, group, run)
    E.close()


@pytest.mark.parametrize("den,nw", [(5, 8), (1, 8), (32, 4), (3, 64)])
def test_warp_cooperative_loop_cpu(emu, den, nw):
    """kz_warp_trace (kz_kernels.cuh) restated lane by lane in tests/hostemu -- lane refill from a cursor, triangle groups postponed to
    the stack (two entries) or waited with in registers, pops of either kind of entry: whatever the schedule, the hits are those of
    the plain per-ray loop, bit for bit (ties included: the soup has duplicated references)."""
    for sb, rays in ((scenes.soup_scene(60000), np.concatenate([scenes.primary_rays(48), scenes.incoherent_rays(6000)])),
                     (scenes.cornell_scene(16, 16, 4), scenes.incoherent_rays(5000, extent=0.95))):
        E = emu.Emu(sb.desc())
        a = E.trace(rays)
        b, ev = E.trace_warp(rays, den, nw)
        assert a.tobytes() == b.tobytes()
        if den < 32:                                               # den 32: a group is never put aside (the plain schedule)
            assert ev["postponed"] > 0 and ev["waited"] > 0        # both ways of putting a group aside were exercised
        E.close()


def test_shadow_walk_early_stop_is_exact_cpu(emu):
    """The device's shadow and occlusion jobs end a segment at the first opaque hit once no invisible emitter can lie in front of it
    (KzWalk::settles, restated in tests/hostemu with the same kz_trav_misses_box): flags AND segment counts must equal the closest-hit
    walk of integrator.cpp:259-294 for every ray -- also for rays aimed at, grazing and starting inside the bounds of the invisible
    emitters, with visible emitters, and for the integrators whose shadow rays take any hit."""
    rng = np.random.default_rng(11)
    for kw in ({}, {"visible_light": True}):
        sb = scenes.cornell_scene(16, 16, 4, **kw)
        for kind in ("path_mis", "ao"):
            sb.set_integrator(kind=kind)
            E = emu.Emu(sb.desc())
            rays = scenes.incoherent_rays(40000, extent=0.97, seed=5)
            rays["tmax"] *= 1.5
            # a third of the rays end on (or just beyond) the ceiling emitters, as the shadow rays of the integrator do
            k = rays.shape[0] // 3
            tgt = np.stack([rng.uniform(-0.65, 0.35, k), rng.choice([0.9, 0.98, 1.05], k), rng.uniform(-0.35, 0.95, k)], 1).astype(np.float32)
            dvec = tgt - rays["o"][:k]
            ln = np.linalg.norm(dvec, axis=1).astype(np.float32)
            rays["d"][:k] = dvec / ln[:, None]
            rays["tmax"][:k] = ln * rng.choice([0.999, 1.0, 1.2], k).astype(np.float32)
            o0, s0 = E.occluded(rays, 1e-3)
            o1, s1, stops = E.occluded_early(rays, 1e-3)
            assert np.array_equal(o0, o1) and np.array_equal(s0, s1), (kw, kind)
            assert stops > 1000
            if kind == "path_mis" and not kw:
                assert s0.max() >= 2          # some rays stepped through an invisible emitter
            E.close()
