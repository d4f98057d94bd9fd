"""One profiled frame of the wavefront path tracer for ncu (`--profile-from-start off`): the first frame warms up (pools, caches), the
second is bracketed by cudaProfilerStart/Stop.  usage: prof_paths.py warm|big [tris] [width height spp]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200")]
import numpy as np, torch
import scenes, pykazen as pk
which = sys.argv[1]
if which == "warm":
    xml = os.path.join(R, "tests", "data", "kazen_scenes", "2022_q1", "WarmStudio", "WarmStudio.xml")
    hs = pk.HostScene(xml, {"camera.width": "i:512", "camera.height": "i:512", "sampler.type": "s:stratified", "sampler.sampleCount": "i:64"})
    G = pk.Gpu(hs.desc, builder=pk.BUILD_HOST_SAH); W = H = 512; spp = 64
else:
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000_000
    W, H, spp = (int(a) for a in sys.argv[3:6]) if len(sys.argv) > 5 else (1920, 1080, 16)
    P, F = scenes.hash_soup_torch(n)
    sb = scenes.big_scene(n, W, H, spp, positions=P.data_ptr(), indices=F.data_ptr(), keep=(P, F))
    G = pk.Gpu(sb.desc(), builder=pk.BUILD_LBVH)
if os.environ.get("KZ_PROF_LANES"):
    G.configure("lanes", int(os.environ["KZ_PROF_LANES"]))
stream = torch.cuda.current_stream().cuda_stream
G.render_device(0, spp, stream=stream); torch.cuda.synchronize(); G.stats(reset=True)
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); G.render_device(0, spp, stream=stream); e1.record(); torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = G.stats()
print(f"{which}: {W}x{H}x{spp} in {e0.elapsed_time(e1):.2f} ms = {W*H*spp/e0.elapsed_time(e1)/1e3:.1f} Mpaths/s; rays ext {st['rays_extension']} shadow {st['rays_shadow']} vertices {st['vertices']} launches {st['kernel_launches']}")
G.close()
