"""Synthetic scenes + ray batches shared by the tests and bench.py (all seeded, numpy only)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "nano-kazen_b200"))
import pykazen as pk  # noqa: E402


def soup_triangles(n, seed=0x5EED):
    """centres uniform in [-1,1]^3, three offsets uniform in [-s,s]^3, s = 2 n^(-1/3).
    (numpy PCG64 stream instead of the PCG32 the survey suggests: same distribution, vectorised.)"""
    rng = np.random.default_rng(seed + n)
    c = rng.uniform(-1, 1, (n, 1, 3)).astype(np.float32)
    s = np.float32(2.0 * n ** (-1.0 / 3.0))
    off = rng.uniform(-s, s, (n, 3, 3)).astype(np.float32)
    P = (c + off).reshape(-1, 3)
    F = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
    return P, F


def soup_scene(n, seed=0x5EED):
    sb = pk.SceneBuilder()
    P, F = soup_triangles(n, seed)
    sb.mesh(P, F, sb.bsdf_diffuse((0.5, 0.5, 0.5)))
    sb.set_camera(64, 64, 40.0, pk.lookat((0, 0, -3), (0, 0, 0), (0, 1, 0)))
    return sb


def primary_rays(res, fov=40.0, origin=(0, 0, -3.0), tmin=1e-4):
    """Pinhole at `origin` looking +z, res x res pixel centres, [tmin, inf)."""
    ys, xs = np.mgrid[0:res, 0:res]
    t = np.tan(np.radians(fov) / 2)
    dx = ((xs + 0.5) / res * 2 - 1) * t
    dy = (1 - (ys + 0.5) / res * 2) * t
    d = np.stack([dx, dy, np.ones_like(dx)], -1).reshape(-1, 3)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = np.zeros(res * res, pk.RAY_DTYPE)
    r["o"] = np.asarray(origin, np.float32); r["d"] = d.astype(np.float32)
    r["tmin"] = tmin; r["tmax"] = np.inf
    return r


def incoherent_rays(n, seed=0xBEEF, extent=1.0):
    """origin uniform in [-extent,extent]^3, direction uniform on the sphere, tnear 1e-3, tfar U[0.1,2]."""
    rng = np.random.default_rng(seed)
    r = np.zeros(n, pk.RAY_DTYPE)
    r["o"] = rng.uniform(-extent, extent, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r["d"] = d.astype(np.float32)
    r["tmin"] = 1e-3
    r["tmax"] = rng.uniform(0.1, 2.0, n).astype(np.float32)
    return r


def quad(p0, p1, p2, p3):
    P = np.array([p0, p1, p2, p3], np.float32)
    F = np.array([[0, 1, 2], [3, 0, 2]], np.uint32)      # kazen's quad split (mesh.cpp:245-251)
    return P, F


def orient(P, F, ref, away=True):
    """flip triangle winding so geometric normals point away from (or towards) `ref`"""
    P = np.asarray(P, np.float32); F = np.array(F, np.uint32)
    for k in range(F.shape[0]):
        a, b, c = P[F[k, 0]], P[F[k, 1]], P[F[k, 2]]
        n = np.cross(b - a, c - a)
        s = np.dot(n, (a + b + c) / 3 - np.asarray(ref, np.float32))
        if (s < 0) == away:
            F[k, 1], F[k, 2] = F[k, 2], F[k, 1]
    return P, F


def uv_sphere(center, radius, nu=24, nv=12):
    """smooth-shaded sphere with normals and uvs"""
    P, N, UV, F = [], [], [], []
    for j in range(nv + 1):
        th = np.pi * j / nv
        for i in range(nu + 1):
            ph = 2 * np.pi * i / nu
            n = np.array([np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph)])
            P.append(np.asarray(center) + radius * n); N.append(n); UV.append([i / nu, 1 - j / nv])
    for j in range(nv):
        for i in range(nu):
            a = j * (nu + 1) + i; b = a + 1; c = a + nu + 1; d = c + 1
            if j != 0:
                F.append([a, b, c])
            if j != nv - 1:
                F.append([b, d, c])
    return np.array(P, np.float32), np.array(N, np.float32), np.array(UV, np.float32), np.array(F, np.uint32)


def cornell_scene(width=64, height=64, spp=16, sampler="stratified", with_texture=False, normalmap=False,
                  visible_light=False, regularization=False, thinlens=None, background=None, max_depth=5):
    """Small closed box: diffuse walls, kiss sphere (smooth normals + uvs), quad mesh light (invisible by default),
    plus a second, flat-shaded kiss block without normals/uvs."""
    sb = pk.SceneBuilder()
    white = sb.bsdf_diffuse((0.73, 0.73, 0.73)); red = sb.bsdf_diffuse((0.65, 0.05, 0.05)); green = sb.bsdf_diffuse((0.12, 0.45, 0.15))
    # floor, ceiling, back, left, right  (box [-1,1]^3, open towards -z)
    for (a, b, c, d, m) in [
        ((-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1), white),
        ((-1, 1, -1), (-1, 1, 1), (1, 1, 1), (1, 1, -1), white),
        ((-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1), white),
        ((-1, -1, -1), (-1, -1, 1), (-1, 1, 1), (-1, 1, -1), red),
        ((1, -1, -1), (1, 1, -1), (1, 1, 1), (1, -1, 1), green),
    ]:
        P, F = orient(*quad(a, b, c, d), (0, 0, 0), away=False)
        sb.mesh(P, F, m)
    if with_texture:
        rng = np.random.default_rng(7)
        img = rng.uniform(0.1, 0.9, (32, 64, 3)).astype(np.float32)
        base = sb.tex_image(img, scale=2.0, srgb=True)
        rough = sb.tex_colorramp(0.2, 0.6, sb.tex_image(img[:, :, ::-1].copy(), srgb=False))
        metal = sb.tex_blend("mix", sb.tex_constant((0.3, 0.3, 0.3)), sb.tex_constant((0, 0, 0)), sb.tex_constant((1, 1, 1)))
    else:
        base = sb.tex_constant((0.871, 0.376, 0.0)); rough = sb.tex_constant((0.35, 0, 1)); metal = sb.tex_constant((0.1, 0, 0))
    kiss = sb.bsdf_kiss(base, rough, metal, specular=0.5, specular_tint=0.2, clearcoat=0.5, clearcoat_roughness=0.3, sheen=0.2, sheen_tint=0.4)
    mat = kiss
    if normalmap:
        rng = np.random.default_rng(11)
        nm = np.zeros((16, 16, 3), np.float32)
        nm[..., 0] = 0.5 + 0.2 * rng.uniform(-1, 1, (16, 16)); nm[..., 1] = 0.5 + 0.2 * rng.uniform(-1, 1, (16, 16)); nm[..., 2] = 0.95
        mat = sb.bsdf_normalmap(sb.tex_image(nm, scale=3.0, srgb=False), kiss)
    P, N, UV, F = uv_sphere((-0.35, -0.55, 0.2), 0.45)
    sb.mesh(P, F, mat, normals=N, uvs=UV)
    # flat block without normals / uvs
    kiss2 = sb.bsdf_kiss(sb.tex_constant((0.2, 0.3, 0.8)), sb.tex_constant((0.6, 0, 0)), sb.tex_constant((0.8, 0, 0)))
    bx = np.array([[0.25, -1, -0.3], [0.75, -1, -0.3], [0.75, -1, 0.3], [0.25, -1, 0.3],
                   [0.25, -0.3, -0.3], [0.75, -0.3, -0.3], [0.75, -0.3, 0.3], [0.25, -0.3, 0.3]], np.float32)
    bf = np.array([[4, 5, 6], [7, 4, 6], [0, 1, 5], [4, 0, 5], [1, 2, 6], [5, 1, 6], [2, 3, 7], [6, 2, 7], [3, 0, 4], [7, 3, 4]], np.uint32)
    bx, bf = orient(bx, bf, (0.5, -0.65, 0.0), away=True)
    sb.mesh(bx, bf, kiss2)
    # ceiling light, two stacked quads so that shadow rays have an invisible light to step through
    lt = sb.light((17.0 * 0.8, 12.0 * 0.8, 4.0 * 0.8), primary_visibility=visible_light)
    P, F = quad((-0.3, 0.98, -0.3), (0.3, 0.98, -0.3), (0.3, 0.98, 0.3), (-0.3, 0.98, 0.3))
    sb.mesh(P, F, white, light=lt)
    lt2 = sb.light((2.0, 2.0, 3.0), primary_visibility=visible_light)
    P, F = quad((-0.6, 0.9, 0.5), (-0.2, 0.9, 0.5), (-0.2, 0.9, 0.9), (-0.6, 0.9, 0.9))
    sb.mesh(P, F, white, light=lt2)
    if background is not None:
        sb.background = sb.tex_background(1.0, sb.tex_constant(background))
    sb.set_camera(width, height, 39.0, pk.lookat((0, 0, -3.4), (0, 0, 0), (0, 1, 0)), near=0.1, far=100.0, thinlens=thinlens)
    sb.set_sampler(sampler, spp)
    sb.set_filter("gaussian")
    sb.set_integrator(max_depth=max_depth, regularization=regularization)
    return sb


def gallery_scene(width=64, height=48, spp=16, sampler="stratified", max_depth=6):
    """Every remaining BSDF plugin in one closed room: dielectric sphere, mirror wall, lambertian (textured) floor,
    ggx block, roughconductor / roughplastic / roughdielectric spheres, plus a normal-mapped roughplastic."""
    sb = pk.SceneBuilder()
    rng = np.random.default_rng(21)
    img = rng.uniform(0.2, 0.9, (16, 16, 3)).astype(np.float32)
    white = sb.bsdf_diffuse((0.7, 0.7, 0.7))
    lamb = sb.bsdf_lambertian(sb.tex_image(img, scale=4.0, srgb=False))
    mirror = sb.bsdf_mirror()
    glass = sb.bsdf_dielectric()
    ggx = sb.bsdf_ggx(sb.tex_constant((0.9, 0.6, 0.3)), roughness=0.4, anisotropy=0.2)
    gold = sb.bsdf_roughconductor(0.3, "Au")
    plastic = sb.bsdf_roughplastic(0.25, kd=(0.2, 0.3, 0.6))
    frosted = sb.bsdf_roughdielectric(0.35)
    nm = np.zeros((8, 8, 3), np.float32); nm[..., 0] = 0.5 + 0.15 * rng.uniform(-1, 1, (8, 8)); nm[..., 1] = 0.5 + 0.15 * rng.uniform(-1, 1, (8, 8)); nm[..., 2] = 0.95
    bumpy = sb.bsdf_normalmap(sb.tex_image(nm, scale=2.0, srgb=False), sb.bsdf_roughplastic(0.2, kd=(0.6, 0.2, 0.2)))
    for (a, b, c, d, m) in [
        ((-2, -1, -1), (2, -1, -1), (2, -1, 1.5), (-2, -1, 1.5), lamb),         # floor
        ((-2, 1.2, -1), (-2, 1.2, 1.5), (2, 1.2, 1.5), (2, 1.2, -1), white),    # ceiling
        ((-2, -1, 1.5), (2, -1, 1.5), (2, 1.2, 1.5), (-2, 1.2, 1.5), mirror),   # back wall: mirror
        ((-2, -1, -1), (-2, -1, 1.5), (-2, 1.2, 1.5), (-2, 1.2, -1), white),
        ((2, -1, -1), (2, 1.2, -1), (2, 1.2, 1.5), (2, -1, 1.5), white),
    ]:
        P, F = orient(*quad(a, b, c, d), (0, 0, 0.2), away=False)
        if m == lamb:
            sb.mesh(P, F, m, uvs=np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32), normals=np.tile(np.array([[0, 1, 0]], np.float32), (4, 1)))
        else:
            sb.mesh(P, F, m)
    for k, m in enumerate((glass, gold, plastic, frosted, bumpy)):
        P, N, UV, F = uv_sphere((-1.4 + 0.7 * k, -0.65, 0.3 + 0.15 * (k % 2)), 0.33, nu=20, nv=10)
        sb.mesh(P, F, m, normals=N, uvs=UV)
    bx = np.array([[-0.3, -1, -0.6], [0.3, -1, -0.6], [0.3, -1, -0.2], [-0.3, -1, -0.2], [-0.3, -0.8, -0.6], [0.3, -0.8, -0.6], [0.3, -0.8, -0.2], [-0.3, -0.8, -0.2]], np.float32)
    bf = np.array([[4, 5, 6], [7, 4, 6], [0, 1, 5], [4, 0, 5], [1, 2, 6], [5, 1, 6], [2, 3, 7], [6, 2, 7], [3, 0, 4], [7, 3, 4]], np.uint32)
    bx, bf = orient(bx, bf, (0, -0.9, -0.4), away=True)
    sb.mesh(bx, bf, ggx, uvs=np.zeros((8, 2), np.float32), normals=None)
    lt = sb.light((14.0, 13.0, 11.0))
    P, F = orient(*quad((-0.5, 1.18, 0.0), (0.5, 1.18, 0.0), (0.5, 1.18, 0.8), (-0.5, 1.18, 0.8)), (0, 0, 0.2), away=False)
    sb.mesh(P, F, white, light=lt)
    sb.background = sb.tex_background(1.0, sb.tex_constant((0.05, 0.06, 0.08)))
    sb.set_camera(width, height, 50.0, pk.lookat((0, 0.1, -3.2), (0, -0.3, 0.3), (0, 1, 0)), near=0.1, far=100.0)
    sb.set_sampler(sampler, spp)
    sb.set_filter("gaussian")
    sb.set_integrator(max_depth=max_depth)
    return sb


def studio_scene(width=512, height=512, spp=64, sampler="stratified"):
    """Procedural stand-in for kazen's smallest shipped scene (scene/2022_q1/WarmStudio: 15 872-triangle kiss ball on a
    2 048-triangle diffuse backdrop, lit by a 32-triangle invisible mesh light array, mitchell filter; BASELINE configs[0]).
    Same triangle counts, materials, light radiance and camera; the geometry is generated here, not copied."""
    sb = pk.SceneBuilder()
    ball = sb.bsdf_kiss(sb.tex_constant((0.871, 0.376, 0.0)), sb.tex_constant((0.0, 0.0, 1.0)), sb.tex_constant((0.0, 0.0, 0.0)),
                        anisotropy=0.0, specular=0.5, specular_tint=0.0, clearcoat=0.0, clearcoat_roughness=0.0, sheen=0.0, sheen_tint=0.0)
    P, N, UV, F = uv_sphere((0.0, 1.0071, 0.0), 0.988, nu=128, nv=63)
    assert F.shape[0] == 15872
    sb.mesh(P, F, ball, normals=N, uvs=UV)
    # backdrop: floor -> quarter cylinder -> wall, swept along z (32 x 32 quads = 2048 triangles), normals towards the camera side
    prof = []
    for k in range(33):
        t = k / 32.0
        if t < 0.4: prof.append((-7.8 + 8.2 * (t / 0.4) * 1.0 + 0.0, 0.0))                    # floor x from -7.8 .. 0.4 (camera is at x = -4.86 looking +x)
        elif t < 0.7:
            a = (t - 0.4) / 0.3 * (np.pi / 2); prof.append((0.4 + 3.0 * np.sin(a), 3.0 - 3.0 * np.cos(a)))
        else: prof.append((3.4, 3.0 + (t - 0.7) / 0.3 * 6.9))
    prof = np.array(prof)
    zs = np.linspace(-8.1, 8.1, 33)
    BP = np.array([[x, y, z] for z in zs for (x, y) in prof], np.float32)
    BF = []
    for j in range(32):
        for i in range(32):
            a = j * 33 + i; b = a + 1; c = a + 33; d = c + 1
            BF += [[a, b, d], [a, d, c]]
    BP, BF = orient(BP, np.array(BF, np.uint32), (-3.0, 3.0, 0.0), away=False)
    assert BF.shape[0] == 2048
    sb.mesh(BP, BF, sb.bsdf_diffuse((0.05, 0.05, 0.05)))
    # 4 x 4 array of 1 x 1 light panels, tilted towards the ball
    LP, LF = [], []
    for r in range(4):
        for c in range(4):
            x0 = -2.246 + 1.165 * c; y0 = 1.857 + 1.13 * r; z0 = -3.97 + 0.755 * r
            q = [(x0 + 1, y0, z0), (x0, y0, z0), (x0 + 1, y0 + 0.828, z0 + 0.56), (x0, y0 + 0.828, z0 + 0.56)]
            k = len(LP); LP += q; LF += [[k, k + 1, k + 3], [k, k + 3, k + 2]]
    LP, LF = orient(np.array(LP, np.float32), np.array(LF, np.uint32), (0.0, 1.0, 0.0), away=False)
    lt = sb.light((4.0 * 0.63827, 4.0 * 0.572175, 4.0 * 0.420238), primary_visibility=False)
    sb.mesh(LP, LF, sb.bsdf_diffuse((0.5, 0.5, 0.5)), light=lt)
    c2w = np.array([[4.371138828673793e-08, -4.371138828673793e-08, 1.0, -4.857681751251221], [0.0, 1.0, 4.371138828673793e-08, 0.9879673719406128],
                    [-1.0, -1.910685676922942e-15, 4.371138828673793e-08, 0.0], [0, 0, 0, 1]], np.float32)
    sb.set_camera(width, height, 49.13434207760448, c2w, near=0.10000000149011612, far=100.0)
    sb.set_sampler(sampler, spp)
    sb.set_filter("mitchell")
    sb.set_integrator(max_depth=5)
    return sb


def rel_mse(a, b, eps=1e-2):
    """per-channel relative MSE of image a against reference b"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return ((a - b) ** 2 / (b ** 2 + eps)).mean(axis=(0, 1))


# ------------------------------------------------------------------------------------------------ procedural soup from a seed
# BASELINE configs[4]: a 10^8-triangle procedural soup "generated on device from a seed".  The triangles are a pure
# function of (triangle index, seed) through a 32-bit integer hash, so the GPU (torch, int64 arithmetic masked to 32 bits) and the
# CPU (numpy uint32) produce bit-identical float32 coordinates without 3.6 GB of positions ever crossing PCIe.
# Same distribution as soup_triangles(): centres uniform in [-1,1]^3, three offsets uniform in [-s,s]^3, s = 2 n^(-1/3).
_HK1, _HK2 = 0x7FEB352D, 0x846CA68B


def _lowbias32_np(x):
    x = x.astype(np.uint32, copy=True)
    x ^= x >> np.uint32(16); x *= np.uint32(_HK1); x ^= x >> np.uint32(15); x *= np.uint32(_HK2); x ^= x >> np.uint32(16)
    return x


def hash_soup_numpy(n, seed=0x5EED, first=0, count=None, chunk=1 << 20):
    """positions (3*count, 3) float32 of triangles [first, first+count) of the n-triangle hash soup (chunks on a thread pool: numpy
    releases the GIL inside its loops, and 10^8 triangles are 3.6 GB of coordinates)"""
    count = n - first if count is None else count
    if count > chunk:
        from concurrent.futures import ThreadPoolExecutor
        out = np.empty((3 * count, 3), np.float32)

        def work(b):
            c = min(chunk, count - b)
            out[3 * b: 3 * (b + c)] = hash_soup_numpy(n, seed, first + b, c, chunk)
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 1) as ex:
            list(ex.map(work, range(0, count, chunk)))
        return out
    s = np.float32(2.0 * n ** (-1.0 / 3.0))
    kc = np.uint32((seed * 2654435761 + 0x9E3779B9) & 0xFFFFFFFF); ko = np.uint32((seed * 40503 + 0x85EBCA6B) & 0xFFFFFFFF)
    t = np.arange(first, first + count, dtype=np.uint32)
    with np.errstate(over="ignore"):
        cc = (t[:, None] * np.uint32(3) + np.arange(3, dtype=np.uint32)[None, :])                    # (count, 3)
        hc = _lowbias32_np(_lowbias32_np(cc) ^ kc)
        oc = (t[:, None] * np.uint32(9) + np.arange(9, dtype=np.uint32)[None, :])                    # (count, 9)
        ho = _lowbias32_np(_lowbias32_np(oc) ^ ko)
    uc = (hc >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    uo = (ho >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    c = uc * np.float32(2) - np.float32(1)
    off = (uo * np.float32(2) - np.float32(1)) * s
    return (c[:, None, :] + off.reshape(count, 3, 3)).reshape(-1, 3)


def hash_soup_torch(n, seed=0x5EED, device="cuda", chunk=1 << 24):
    """the same triangles generated on `device`: (3n, 3) float32 positions and (n, 3) int32 indices (torch tensors)"""
    import torch
    s = float(np.float32(2.0 * n ** (-1.0 / 3.0)))
    kc = (seed * 2654435761 + 0x9E3779B9) & 0xFFFFFFFF; ko = (seed * 40503 + 0x85EBCA6B) & 0xFFFFFFFF
    M = 0xFFFFFFFF

    def lb(x):
        x = x ^ (x >> 16); x = (x * _HK1) & M; x = x ^ (x >> 15); x = (x * _HK2) & M; x = x ^ (x >> 16)
        return x
    P = torch.empty((3 * n, 3), dtype=torch.float32, device=device)
    for first in range(0, n, chunk):
        cnt = min(chunk, n - first)
        t = torch.arange(first, first + cnt, dtype=torch.int64, device=device)
        cc = (t[:, None] * 3 + torch.arange(3, dtype=torch.int64, device=device)[None, :]) & M
        oc = (t[:, None] * 9 + torch.arange(9, dtype=torch.int64, device=device)[None, :]) & M
        hc = lb(lb(cc) ^ kc); ho = lb(lb(oc) ^ ko)
        uc = (hc >> 8).to(torch.float32) * (2.0 ** -24); uo = (ho >> 8).to(torch.float32) * (2.0 ** -24)
        c = uc * 2.0 - 1.0
        off = (uo * 2.0 - 1.0) * s
        P[3 * first: 3 * (first + cnt)] = (c[:, None, :] + off.reshape(cnt, 3, 3)).reshape(-1, 3)
    F = torch.arange(3 * n, dtype=torch.int32, device=device).reshape(-1, 3)
    return P, F


def big_scene(n, width=3840, height=2160, spp=1024, positions=None, indices=None, keep=None):
    """BASELINE configs[4]: n-triangle hash soup, diffuse, constant environment light as the only emitter (NEE draws
    are consumed but skipped, integrator.cpp:247-248), pinhole camera, stratified sampler.  positions / indices: numpy arrays or
    raw device pointers (ints) of arrays that already live on the GPU (kz_mesh_desc accepts both); default: numpy hash soup."""
    import ctypes as C
    sb = pk.SceneBuilder()
    bs = sb.bsdf_diffuse((0.5, 0.5, 0.5))
    if positions is None:
        positions = hash_soup_numpy(n); indices = np.arange(3 * n, dtype=np.uint32).reshape(-1, 3)
    if isinstance(positions, (int, np.integer)):
        m = pk.MeshDesc()
        m.positions = C.cast(int(positions), pk.c_float_p); m.indices = C.cast(int(indices), pk.c_u32_p)
        m.n_vertices, m.n_triangles, m.bsdf, m.light = 3 * n, n, bs, -1
        sb.meshes.append(m); sb._keep.append(keep)
    else:
        sb.mesh(positions, indices, bs)
    sb.background = sb.tex_background(1.0, sb.tex_constant((0.8, 0.9, 1.0)))
    sb.set_camera(width, height, 40.0, pk.lookat((0, 0, -3), (0, 0, 0), (0, 1, 0)))
    sb.set_sampler("stratified", spp)
    sb.set_filter("gaussian")
    sb.set_integrator(max_depth=5)
    return sb
