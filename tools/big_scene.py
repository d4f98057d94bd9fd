"""BASELINE configs[4] stand-in: procedural triangle soup + constant environment light, path traced (LBVH accel built on the GPU).
usage: big_scene.py <tris> <width> <height> <spp>"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200")]
import numpy as np, torch
import scenes, pykazen as pk
n, W, H, spp = (int(a) for a in sys.argv[1:5])
t0 = time.time()
sb = scenes.soup_scene(n)
sb.background = sb.tex_background(1.0, sb.tex_constant((0.8, 0.9, 1.0)))
sb.set_camera(W, H, 40.0, pk.lookat((0, 0, -3), (0, 0, 0), (0, 1, 0)))
sb.set_sampler("stratified", spp)
d = sb.desc()
t1 = time.time()
G = pk.Gpu(d, builder=pk.BUILD_LBVH, lib_path=os.environ.get("KZGPU_LIB", pk.LIB_GPU))
t2 = time.time()
st0 = G.stats()
G.render_device(0, sb.sampler.sample_count); torch.cuda.synchronize(); G.stats(reset=True)      # warm-up at full size: the path pool is grown on demand
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); G.render_device(0, sb.sampler.sample_count); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); st = G.stats()
paths = W * H * sb.sampler.sample_count
print(f"{n} tris: scene gen {t1-t0:.1f}s, upload+LBVH {t2-t1:.2f}s ({st0['bvh_nodes']} nodes, {st0['bvh_bytes']/2**30:.2f} GiB); {W}x{H}x{sb.sampler.sample_count}spp in {ms:.1f} ms = "
      f"{paths/ms/1e3:.1f} Mpaths/s, {(st['rays_extension']+st['rays_shadow'])/ms/1e3:.1f} Mrays/s, {st['rays_extension']/paths:.2f} rays/path, trace {st['ms_trace']:.1f} ms shade {st['ms_shade']:.1f} ms, "
      f"GPU mem {torch.cuda.mem_get_info()[0]/2**30:.1f} GiB free")
G.close()
