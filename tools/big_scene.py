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
# under torchrun: scene replicated, sample indices sharded over the ranks, ONE NCCL reduce of the frame
rank, world, local = (int(os.environ.get(k, d0)) for k, d0 in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G = pk.Gpu(d, devices=(local,), builder=pk.BUILD_LBVH, lib_path=os.environ.get("KZGPU_LIB", pk.LIB_GPU))
t2 = time.time()
st0 = G.stats()
fh, fw, _ = G.frame_shape()
frame = torch.zeros((fh, fw, 4), dtype=torch.float32, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
s0, s1 = pk.shard_range(0, sb.sampler.sample_count, rank, world)
def render():
    G.render_device(s0, s1, device=0, clear=True, stream=stream, frame_ptr=frame.data_ptr())
    if world > 1: dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
render(); torch.cuda.synchronize(); G.stats(reset=True)      # warm-up at full size: the path pool is grown on demand
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); render(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); st = G.stats()
if world > 1:
    t = torch.tensor([ms, float(st["rays_extension"]), float(st["rays_shadow"])], dtype=torch.float64, device="cuda")
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
    ms = float(tmax[0].item()); st["rays_extension"] = float(t[1].item()); st["rays_shadow"] = float(t[2].item())
paths = W * H * sb.sampler.sample_count
if rank == 0: print(f"{world} GPU(s), mean rgb {float(frame[..., :3].sum() / frame[..., 3].sum().clamp_min(1e-9)):.5f}; ", end="")
if rank == 0: print(f"{n} tris: scene gen {t1-t0:.1f}s, upload+LBVH {t2-t1:.2f}s ({st0['bvh_nodes']} nodes, {st0['bvh_bytes']/2**30:.2f} GiB); {W}x{H}x{sb.sampler.sample_count}spp in {ms:.1f} ms = "
      f"{paths/ms/1e3:.1f} Mpaths/s, {(st['rays_extension']+st['rays_shadow'])/ms/1e3:.1f} Mrays/s, {st['rays_extension']/paths:.2f} rays/path, trace {st['ms_trace']:.1f} ms shade {st['ms_shade']:.1f} ms, "
      f"GPU mem {torch.cuda.mem_get_info()[0]/2**30:.1f} GiB free")
G.close()
if world > 1: dist.destroy_process_group()
