#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric (Mrays/s / Mpaths/s) on the triangle-soup intersection workload
(configs[1]) plus a path-tracing leg (Cornell-class scene), one process per GPU.

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...        # the CPU port of the reference path on the host cores

A step = one pass of the intersection hot path (kzgpu_trace_device, the rtcIntersect1 replacement)
over one primary-ray batch and one incoherent shadow-ray batch, rays resident in HBM.  Multi-GPU is
weak scaling: scene replicated, every rank traces its own batches, no data-path collective; the
path-tracing leg shards sample indices and ends in ONE NCCL reduce of the accumulated frame.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "nano-kazen_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def b_ray(n_tris, shadow=False):
    """algorithmic bytes per ray = ray in + hit out + one root-to-leaf descent of
    80-byte 8-wide nodes + one 4-triangle leaf of 48-byte triangles."""
    return 32 + (8 if shadow else 20) + 80 * math.ceil(math.log(max(n_tris, 8) / 4.0, 8)) + 4 * 48


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def soup_workload(n_tris, n_primary_res, n_incoherent, seed_offset=0):
    import scenes
    sb = scenes.soup_scene(n_tris)
    prim = scenes.primary_rays(n_primary_res)
    inc = scenes.incoherent_rays(n_incoherent, seed=0xBEEF + seed_offset)
    return sb, prim, inc


def run_reference(args):
    """The reference arm: the CPU port (oracle) of the same path on all host cores, bounded sample."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kzo_py
    kzo_py.build()
    cores = os.cpu_count() or 1
    sb, prim, inc = soup_workload(args.tris, 512, 1 << 18)
    O = kzo_py.Oracle(sb.desc())
    n = prim.shape[0] + inc.shape[0]
    for _ in range(args.warmup):
        O.trace(prim[: 1 << 14]); O.trace(inc[: 1 << 14], shadow=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.trace(prim); O.trace(inc, shadow=True)
    dt = time.perf_counter() - t0
    val = args.steps * n / dt / 1e6
    sample = f"{prim.shape[0]} primary (512x512 pinhole grid) + {inc.shape[0]} incoherent shadow rays per step, {args.tris}-triangle soup"
    line = {"impl": "reference", "metric": "Mrays/s", "value": val, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"triangle-soup intersection microbench, {args.tris} tris, primary + shadow rays (BASELINE configs[1])", "tris": args.tris},
            "cpu_baseline": {"value": val, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "CPU port of the reference path (oracle: median-split BVH2 + Embree-robust Pluecker test, std::thread over all cores); "
                    "the reference itself (Embree 3.13 + TBB + OIIO) cannot be built in this image"}
    emit(line)


_JSON_FD = None


def claim_stdout():
    """Everything that libraries print on stdout (NCCL's version banner, ...) is sent to stderr; the ONE JSON line is
    written to the original stdout by emit()."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_JSON_FD if _JSON_FD is not None else 1, (json.dumps(line) + "\n").encode())


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tris", type=int, default=1 << 20)
    ap.add_argument("--primary-res", type=int, default=4096)
    ap.add_argument("--incoherent", type=int, default=1 << 24)
    ap.add_argument("--builder", default="sah", choices=["sah", "lbvh"])
    ap.add_argument("--no-paths", action="store_true", help="skip the path-tracing (Mpaths/s) leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return 0

    import torch
    import torch.distributed as dist
    import pykazen as pk
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    affinity = None
    if world > 1:
        # one process per GPU: run on (and first-touch the pinned ray / hit buffers from) the CPUs next to this GPU, when the
        # platform exposes a proper subset of them; the end-to-end leg is host-memory / PCIe bound with 8 ranks copying at once
        try:
            prop = torch.cuda.get_device_properties(local)
            bus, dom, dev = getattr(prop, "pci_bus_id", None), getattr(prop, "pci_domain_id", 0), getattr(prop, "pci_device_id", 0)
            if bus is not None:
                with open(f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist") as f:
                    cpus = set()
                    for part in f.read().strip().split(","):
                        if part:
                            a, _, b = part.partition("-")
                            cpus.update(range(int(a), int(b or a) + 1))
                allowed = os.sched_getaffinity(0)
                near = cpus & allowed
                if near and near != allowed:
                    os.sched_setaffinity(0, near)
                    affinity = f"{len(near)} of {len(allowed)} cpus (local to GPU {local})"
        except Exception:
            affinity = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ------------------------------------------------------------------ workload + accel
    sb, prim, inc = soup_workload(args.tris, args.primary_res, args.incoherent, seed_offset=rank)
    G = pk.Gpu(sb.desc(), devices=(local,), builder=pk.BUILD_HOST_SAH if args.builder == "sah" else pk.BUILD_LBVH)
    build_ms = G.stats()["ms_build"]
    stream = torch.cuda.current_stream().cuda_stream
    batches = []
    for r in (prim, inc):
        host = torch.from_numpy(r.view(np.float32).reshape(-1, 8)).pin_memory()
        batches.append({"n": r.shape[0], "host": host, "dev": host.cuda(), "hits": torch.empty((r.shape[0], 5), dtype=torch.float32, device="cuda"),
                        "host_hits": torch.empty((r.shape[0], 5), dtype=torch.float32).pin_memory()})
    rays_per_step = sum(b["n"] for b in batches)

    def step():
        for k, b in enumerate(batches):
            G.trace_device(b["dev"].data_ptr(), b["n"], b["hits"].data_ptr(), shadow=(k == 1), device=0, stream=stream)

    for _ in range(args.warmup):
        step()
    barrier()
    G.stats(reset=True)
    clocks = ClockSampler(local) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.stop() if clocks else None
    st = G.stats(reset=True)
    launches = int(sum_over_ranks(st["kernel_launches"]))
    value = world * args.steps * rays_per_step / (ms * 1e-3) / 1e6
    hit_frac = float((batches[0]["hits"][:, 4].view(torch.int32) != -1).float().mean().item())

    # roofline of the dominant kernel (k_trace): CUDA events inside the library around every launch
    peak, peak_kind = measured_peaks()
    trace_launches = args.steps * len(batches)               # one k_trace per batch
    kernel_ms = st["ms_trace"] / max(1, trace_launches)
    bytes_per_launch = 0.5 * (batches[0]["n"] * b_ray(args.tris) + batches[1]["n"] * b_ray(args.tris))   # both batches return 20-byte hits
    achieved = bytes_per_launch / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture of this workload, if there is one
        with open(os.path.join(ROOT, "profiles", "r4_traffic.json")) as f:
            for tj in json.load(f)["workloads"]:
                if tj["tris"] == args.tris and tj["builder"] == args.builder and tj["rays_per_launch"] == batches[0]["n"] == batches[1]["n"]:
                    traffic = tj["dram_bytes_per_launch_mean"]
    except Exception:
        traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel": "k_trace", "kernel_ms": kernel_ms, "bytes_per_ray": b_ray(args.tris), "peak_kind": peak_kind}

    # ------------------------------------------------------------------ e2e through the host-buffer C ABI
    def e2e_step():
        for k, b in enumerate(batches):
            G.trace_host_ptr(b["host"].data_ptr(), b["n"], b["host_hits"].data_ptr(), shadow=(k == 1))
    e2e_step()
    barrier()
    e2e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": world * e2e_steps * rays_per_step / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": rays_per_step * 32,
           "d2h_bytes_per_step": rays_per_step * 20, "steps": e2e_steps, "api": "kzgpu_trace (pinned host buffers)"}
    same = bool(torch.equal(batches[1]["host_hits"].view(torch.int32), batches[1]["hits"].cpu().view(torch.int32)))   # bit compare (miss ids are NaN patterns)

    # ------------------------------------------------------------------ path-tracing legs (Mpaths/s)
    def path_leg(make_scene, label):
        import scenes
        W = H = 512; spp = 64
        sbp = make_scene(scenes, W, H, spp)
        GP = pk.Gpu(sbp.desc(), devices=(local,), builder=pk.BUILD_HOST_SAH)
        fh, fw, _ = GP.frame_shape()
        frame = torch.zeros((fh, fw, 4), dtype=torch.float32, device="cuda")
        s0, s1 = pk.shard_range(0, spp, rank, world)

        def pstep():
            GP.render_device(s0, s1, device=0, clear=True, stream=stream, frame_ptr=frame.data_ptr())
            if world > 1:
                dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)      # the only collective: ImageBlock merge (block.cpp:87-96)
        for _ in range(2):
            pstep()
        barrier(); GP.stats(reset=True)
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        psteps = 3
        p0.record()
        for _ in range(psteps):
            pstep()
        p1.record()
        barrier()
        pms = max_over_ranks(p0.elapsed_time(p1)) / psteps
        ps = GP.stats()
        npaths = W * H * spp                                  # strong scaling: total work fixed
        rays_total = sum_over_ranks(ps["rays_extension"] + ps["rays_shadow"]) / psteps
        out = {"value": npaths / (pms * 1e-3) / 1e6, "unit": "Mpaths/s", "mrays_per_s": rays_total / (pms * 1e-3) / 1e6, "ms_per_frame": pms,
               "scene": label + f", {W}x{H}, {spp} spp, stratified, path_mis maxDepth 5",
               "rays_per_path": rays_total / npaths, "scaling": "strong (sample-index shards + one NCCL reduce)",
               "ms_trace": ps["ms_trace"] / psteps, "ms_shade": ps["ms_shade"] / psteps,
               "ms_note": "ms_trace / ms_shade are summed over the two concurrent lanes of a device (chunks overlap), so they exceed ms_per_frame",
               "mean_rgb": None}
        # B_path = n_ext*B_ray + n_sh*(B_ray - 12) + n_vtx*512 + 256, with the MEASURED per-path counts
        n_tris = sum(int(m.n_triangles) for m in sbp.meshes)
        n_ext = sum_over_ranks(ps["rays_extension"]) / psteps / npaths
        n_sh = sum_over_ranks(ps["rays_shadow"]) / psteps / npaths
        n_vtx = sum_over_ranks(ps["vertices"]) / psteps / npaths
        br = b_ray(n_tris)
        b_path = n_ext * br + n_sh * (br - 12) + n_vtx * 512 + 256
        peak, _kind = measured_peaks()
        ach = npaths * b_path / (pms * 1e-3) / 1e9 / world       # per GPU, against one GPU's HBM
        out["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                           "bytes_per_path": b_path, "bytes_per_ray": br, "tris": n_tris, "n_ext": n_ext, "n_sh": n_sh, "n_vtx": n_vtx,
                           "note": "whole wavefront (all stages), per GPU; the accel of this scene is L2 resident, so HBM does not bind (DESIGN 7)"}
        if rank == 0:
            rgb, _ = GP.resolve(frame.cpu().numpy())
            out["mean_rgb"] = float(rgb.mean())
        GP.close()
        # time-to-image beside it: the CPU port renders a bounded sample range of the SAME frame on all host cores
        if rank == 0 and world == 1 and not args.no_cpu:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import kzo_py
            OP = kzo_py.Oracle(sbp.desc())
            cs = 2                                                 # 2 of the 64 sample indices of every pixel
            t0 = time.perf_counter()
            cf = OP.render(0, cs)
            dt = time.perf_counter() - t0
            OP.close()
            cmp_ = cs * W * H / dt / 1e6
            out["cpu_baseline"] = {"value": cmp_, "unit": "Mpaths/s", "cores": os.cpu_count() or 1, "kind": "port",
                                   "sample": f"sample indices 0..{cs - 1} of all {W}x{H} pixels of the same frame ({cs * W * H} paths)",
                                   "time_to_image_s": {"cpu_extrapolated": npaths / (cmp_ * 1e6), "gpu": pms * 1e-3}}
        return out

    paths = paths_cornell = None
    if not args.no_paths:
        paths = path_leg(lambda sc, W, H, spp: sc.studio_scene(W, H, spp, "stratified"),
                         "WarmStudio stand-in (15872-tri kiss ball, 2048-tri diffuse backdrop, 32-tri invisible mesh light; BASELINE configs[0] class)")
        paths_cornell = path_leg(lambda sc, W, H, spp: sc.cornell_scene(W, H, spp, "stratified"),
                                 "cornell-class (kiss + diffuse + 2 invisible mesh lights)")

    # ------------------------------------------------------------------ CPU baseline (rank 0, N == 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import kzo_py
        import scenes
        O = kzo_py.Oracle(sb.desc())
        cp, ci = scenes.primary_rays(1024), inc[: 1 << 20]
        t0 = time.perf_counter()
        ho_p = O.trace(cp); ho_i = O.trace(ci, shadow=True)
        dt = time.perf_counter() - t0
        # the sample doubles as a full-size parity check of the timed kernel
        ok = ho_i.tobytes() == batches[1]["hits"][: 1 << 20].cpu().numpy().reshape(-1).view(pk.HIT_DTYPE).tobytes()
        cpu = {"value": (cp.shape[0] + ci.shape[0]) / dt / 1e6, "unit": "Mrays/s", "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{cp.shape[0]} primary (1024x1024 grid) + {ci.shape[0]} incoherent shadow rays of the same workload, once",
               "hits_match_gpu": bool(ok)}
        O.close()

    if rank == 0:
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"triangle-soup intersection microbench, {args.tris} tris, primary + shadow rays (BASELINE configs[1])",
                           "tris": args.tris, "rays_per_step_per_gpu": rays_per_step, "primary": batches[0]["n"], "incoherent_shadow": batches[1]["n"],
                           "builder": args.builder, "accel_build_ms": build_ms, "primary_hit_fraction": hit_frac,
                           "l2": "inputs larger than L2 (512 MiB of rays + 320 MiB of hits per batch)", "parallelism": f"replicas x{world}",
                           "cpu_affinity": affinity},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_matches_device": same, "gpu_launches": launches, "clocks": clk, "paths": paths, "paths_cornell": paths_cornell}
        emit(line)
    G.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
