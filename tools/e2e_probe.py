"""End-to-end rate of kzgpu_trace (pinned host buffers) on the bench workload for a few pipeline chunk counts (KZGPU_TRACE_CHUNKS)."""
import os, sys, time, subprocess
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    [sys.path.insert(0, os.path.join(R, p)) for p in ("tests", "nano-kazen_b200")]
    import numpy as np, torch
    import scenes, pykazen as pk
    sb = scenes.soup_scene(1 << 20); d = sb.desc()
    G = pk.Gpu(d)
    batches = []
    for r in (scenes.primary_rays(4096), scenes.incoherent_rays(1 << 24)):
        host = torch.from_numpy(r.view(np.float32).reshape(-1, 8)).pin_memory()
        batches.append((r.shape[0], host, torch.empty((r.shape[0], 5), dtype=torch.float32).pin_memory()))
    def step():
        for k, (n, h, o) in enumerate(batches): G.trace_host_ptr(h.data_ptr(), n, o.data_ptr(), shadow=(k == 1))
    step(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): step()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    print(f"chunks {os.environ.get('KZGPU_TRACE_CHUNKS', 'default')}: {dt*1e3:.2f} ms/step, {sum(b[0] for b in batches)/dt/1e6:.1f} Mrays/s end to end", flush=True)
    for graded in ("0", "1", "0", "1"):        # chunk sizes graded at both ends of a call (KZGPU_TRACE_GRADED), A/B/A/B in one process
        os.environ["KZGPU_TRACE_GRADED"] = graded
        step(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): step()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
        print(f"   graded={graded}: {dt*1e3:.2f} ms/step, {sum(b[0] for b in batches)/dt/1e6:.1f} Mrays/s", flush=True)
else:
    for k in sys.argv[1:] or ["16", "32", "64"]:
        subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, KZGPU_TRACE_CHUNKS=k))
