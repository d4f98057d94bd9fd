"""Times k_trace for several builds of libkzgpu.so (tuning experiments): tools/variant_bench.py lib1.so lib2.so ..."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200")]
import numpy as np, torch
import scenes, pykazen as pk
n_tris = int(os.environ.get("TRIS", 1 << 20)); res = int(os.environ.get("RES", 2048)); ninc = int(os.environ.get("NINC", 1 << 22))
sb = scenes.soup_scene(n_tris); d = sb.desc()
prim, inc = scenes.primary_rays(res), scenes.incoherent_rays(ninc)
ref = None
stream = torch.cuda.current_stream().cuda_stream
for lib in sys.argv[1:] or [pk.LIB_GPU]:
    G = pk.Gpu(d, lib_path=lib, builder=pk.BUILD_LBVH if os.environ.get("LBVH") else pk.BUILD_HOST_SAH)
    out = []
    for rays in (prim, inc):
        dr = torch.from_numpy(rays.view(np.float32).reshape(-1, 8)).cuda(); dh = torch.empty((rays.shape[0], 5), dtype=torch.float32, device="cuda")
        for _ in range(2): G.trace_device(dr.data_ptr(), rays.shape[0], dh.data_ptr(), stream=stream)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): G.trace_device(dr.data_ptr(), rays.shape[0], dh.data_ptr(), stream=stream)
        e1.record(); torch.cuda.synchronize()
        out.append((rays.shape[0] * 5 / e0.elapsed_time(e1) / 1e3, dh.cpu().view(torch.int32)))
    if ref is None: ref = [o[1] for o in out]
    ok = all(torch.equal(a, o[1]) for a, o in zip(ref, out))
    crc = [int(o[1].to(torch.int64).sum().item()) & 0xFFFFFFFFFFFF for o in out]       # compare across processes (builder knobs are read once per process)
    print(f"hit checksums {crc[0]:012x} {crc[1]:012x}  bvh {G.stats()['bvh_nodes']} nodes {G.stats()['bvh_bytes'] / 2**20:.0f} MiB  build {G.stats()['ms_build']:.0f} ms")
    print(f"{os.path.basename(lib):22s} primary {out[0][0]:8.1f} Mrays/s   incoherent {out[1][0]:8.1f} Mrays/s   same_hits={ok}", flush=True)
    G.close()
