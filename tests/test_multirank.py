"""N>1 host logic on CPU (gloo, world_size 2): sample-index sharding + ONE reduce of the bordered frame
(the cross-GPU analogue of ImageBlock::put(ImageBlock&), block.cpp:87-96).  The per-rank render is done
by the oracle here (test infrastructure); on the GPU box bench.py runs the same decomposition over NCCL."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
import numpy as np
import torch, torch.distributed as dist
for p in ("tests", "oracle", "nano-kazen_b200"):
    sys.path.insert(0, os.path.join(%(root)r, p))
import scenes, kzo_py, pykazen as pk
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sb = scenes.cornell_scene(24, 20, 9, "stratified")
O = kzo_py.Oracle(sb.desc())
spp = sb.sampler.sample_count
s0, s1 = pk.shard_range(0, spp, rank, world)
frame = torch.from_numpy(O.render(s0, s1, threads=2))
dist.reduce(frame, dst=0, op=dist.ReduceOp.SUM)
counts = torch.tensor([s1 - s0], dtype=torch.int64)
dist.all_reduce(counts)
if rank == 0:
    whole = O.render(0, spp, threads=2)
    ok = bool(np.allclose(frame.numpy(), whole, rtol=1e-5, atol=1e-6)) and int(counts.item()) == spp
    print("RESULT", ok, float(np.abs(frame.numpy() - whole).max()))
dist.destroy_process_group()
"""


def test_shard_range_partitions():
    sys.path.insert(0, os.path.join(ROOT, "nano-kazen_b200"))
    import pykazen as pk
    for n in (1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            parts = [pk.shard_range(5, 5 + n, r, world) for r in range(world)]
            assert parts[0][0] == 5 and parts[-1][1] == 5 + n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_sharded_render_gloo(tmp_path, kzo):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29731", str(script)], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")]
    assert line and line[0].split()[1] == "True", r.stdout


def test_reference_arm_prints_one_line_under_torchrun(kzo):
    """bench.py --impl reference: rank 0 alone runs and prints the JSON line, the other rank exits 0"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29732", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--tris", "20000", "--width", "256", "--height", "144"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "Mpaths/s" and j["value"] > 0 and j["cpu_baseline"]["kind"] == "port"
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["n_gpus"] == 2
