"""Path-tracing throughput vs. wavefront pool size (KZGPU_POOL_LOG2) on the bench's Cornell-class scene."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
[sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200")]
import numpy as np, torch
import scenes, pykazen as pk
W = H = 512; spp = 64
libs = [a for a in sys.argv[1:] if a.endswith(".so")] or [pk.LIB_GPU]
for lib in libs:
  for log2 in [int(a) for a in sys.argv[1:] if not a.endswith(".so")] or [22]:
      os.environ["KZGPU_POOL_LOG2"] = str(log2)
      sb = scenes.cornell_scene(W, H, spp, "stratified")
      G = pk.Gpu(sb.desc(), lib_path=lib)
      for _ in range(2):
          G.render_device(0, spp)
      torch.cuda.synchronize(); G.stats(reset=True)
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record(); h0 = time.perf_counter()
      for _ in range(3):
          G.render_device(0, spp)
      h1 = time.perf_counter()
      e1.record(); torch.cuda.synchronize()
      ms = e0.elapsed_time(e1) / 3
      st = G.stats()
      print(f"{os.path.basename(lib)} pool 2^{log2}: {ms:8.2f} ms/frame  {W*H*spp/ms/1e3:8.1f} Mpaths/s  trace {st['ms_trace']/3:7.2f} shade {st['ms_shade']/3:7.2f} launches {st['kernel_launches']//3}  host enqueue {(h1-h0)*1e3/3:6.2f} ms/frame", flush=True)
      G.close()
