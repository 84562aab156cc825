"""Host logic of the multi-GPU path on CPU: shard plans, and a world_size-2 gloo run of
the [gradient | chi2] exchange."""
import os
import subprocess
import sys

import numpy as np
import torch

from gpuvmem_b200 import dist as gdist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_plan_channels_follow_reference_rule():
    plan = gdist.shard_plan(64, [1000] * 64, 8)
    assert len(plan) == 8
    for r, kw in enumerate(plan):
        assert kw["channels"] == list(range(r, 64, 8))
    allc = sorted(c for kw in plan for c in kw["channels"])
    assert allc == list(range(64))


def test_shard_plan_visibility_chunks_cover_everything_once():
    for Z, world in [(10_000_000, 8), (1_048_576, 4), (7, 2), (5, 8)]:
        plan = gdist.shard_plan(1, [Z], world)
        covered = np.zeros(Z, dtype=int)
        for kw in plan:
            a, b = kw["vis_slice"]
            covered[a:max(a, b)] += 1
        assert (covered == 1).all()
    assert gdist.shard_plan(3, [5, 5, 5], 1) == [dict()]


def test_split_join_f64_roundtrip():
    x = torch.tensor(123456789.123456789, dtype=torch.float64)
    assert abs(float(gdist.join_f64(gdist.split_f64(x))) - float(x)) < 1e-6


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gpuvmem_b200 import dist as gdist
rank, world, local = gdist.init_from_env(2)
assert world == 2 and dist.get_backend() == "gloo"
MN = 64
buf = torch.zeros(2 * MN + 2)
buf[:2 * MN] = torch.arange(2 * MN, dtype=torch.float32) * (rank + 1)
chi2 = torch.tensor(1000.25 * (rank + 1), dtype=torch.float64)
buf[2 * MN:] = gdist.split_f64(chi2)
gdist.allreduce_eval(buf)
assert torch.equal(buf[:2 * MN], torch.arange(2 * MN, dtype=torch.float32) * 3)
assert abs(float(gdist.join_f64(buf[2 * MN:])) - 3000.75) < 1e-9
plan = gdist.shard_plan(1, [1001], world)
a, b = plan[rank]["vis_slice"]
n = torch.tensor([b - a]); dist.all_reduce(n); assert int(n) == 1001
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


_WORKER_HOST = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from gpuvmem_b200 import dist as gdist
from gpuvmem_b200 import host
rank, world, local = gdist.init_from_env(2)
assert world == 2 and dist.get_backend() == "gloo"
# what bench.py / a launcher does before creating the per-rank Session: rank 0 makes the NCCL id,
# the launcher's rendezvous carries it, every rank derives its shard of the visibilities
box = [host.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(box, src=0)
assert isinstance(box[0], bytes) and len(box[0]) == 128
ids = [None, None]
dist.all_gather_object(ids, box[0])
assert ids[0] == ids[1]
for Z in ([1001], [700, 900, 1100]):          # visibility chunks / whole channels
    mine = sum(hi - lo for lo, hi in host.shard_plan(Z, world, rank))
    n = torch.tensor([mine]); dist.all_reduce(n); assert int(n) == sum(Z), (Z, int(n))
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_host_layer_rendezvous(tmp_path):
    script = tmp_path / "worker_host.py"
    script.write_text(_WORKER_HOST)
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29534", str(script), ROOT]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2
