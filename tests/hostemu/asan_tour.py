"""Tour of the device routines (compiled for the host with -fsanitize=address,undefined) over every scene family; run by
tests/test_sanitizers.py in a subprocess with libasan/libubsan preloaded.  compute-sanitizer is closed on the GPU pool, so this
is where out-of-bounds indexing / UB in the shared CUDA source is looked for."""
import os, sys, numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
[sys.path.insert(0, p) for p in (os.path.join(ROOT, "tests"), HERE, os.path.join(ROOT, "nano-kazen_b200"))]
import emu_py; emu_py.LIB = sys.argv[1]; emu_py.build = lambda: None
import scenes
for sb in (scenes.cornell_scene(24,16,4,"stratified",with_texture=True,normalmap=True,regularization=True,background=(0.1,0.1,0.2)), scenes.gallery_scene(24,16,4), scenes.studio_scene(32,18,4), scenes.soup_scene(20000)):
    E = emu_py.Emu(sb.desc())
    r = np.concatenate([scenes.primary_rays(24), scenes.incoherent_rays(3000)])
    h = E.trace(r); o,s = E.occluded(scenes.incoherent_rays(2000, extent=0.9), 1e-3)
    f = E.render(); print(E.bvh_info(), float(f[...,:3].sum()))
    E.close()
for kind in ("normals","ao","whitted","path_mats"):
    sb = scenes.cornell_scene(16,12,4,"correlated",visible_light=True); sb.set_integrator(kind=kind)
    E = emu_py.Emu(sb.desc()); print(kind, float(E.render()[...,:3].sum())); E.close()
print("asan tour done")
