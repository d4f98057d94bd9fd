"""A small tour of every kernel for compute-sanitizer (memcheck): LBVH + SAH builds, batch trace, occlusion walk, sampler/camera
probes, path tracing with every BSDF class, the other integrators, mip lookup, resolve."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200")]
import numpy as np
import scenes, pykazen as pk
for builder in (pk.BUILD_HOST_SAH, pk.BUILD_LBVH):
    sb = scenes.soup_scene(3000)
    G = pk.Gpu(sb.desc(), builder=builder)
    r = np.concatenate([scenes.primary_rays(32), scenes.incoherent_rays(2000)])
    h = G.trace(r); print("trace", builder, int((h["geom_id"] != 0xFFFFFFFF).sum()))
    G.close()
sb = scenes.cornell_scene(32, 24, 4, "stratified", with_texture=True, normalmap=True, regularization=True, background=(0.1, 0.1, 0.2))
G = pk.Gpu(sb.desc(), builder=pk.BUILD_LBVH)
occ, seg = G.occluded(scenes.incoherent_rays(2000, extent=0.9), 1e-3); print("occluded", int(occ.sum()), int(seg.max()))
print("samples", G.sample_dump(np.array([[1, 2, 3]], np.int32), "P211").ravel()[:3])
print("camera", G.camera_rays(np.array([[3.5, 4.5, 0.5, 0.5]], np.float32))["d"])
f = G.render(); rgb, s8 = G.resolve(f); print("render", float(rgb.mean()))
print("mip", G.image_lookup(0, np.array([[0.3, 0.7]], np.float32), 2))
G.close()
sb = scenes.gallery_scene(32, 24, 4)
G = pk.Gpu(sb.desc()); print("gallery", float(G.resolve(G.render())[0].mean())); G.close()
for kind in ("normals", "ao", "whitted", "path_mats"):
    sb = scenes.cornell_scene(24, 16, 4, "correlated", visible_light=True); sb.set_integrator(kind=kind)
    G = pk.Gpu(sb.desc()); print(kind, float(G.resolve(G.render())[0].mean())); G.close()
print("done")
