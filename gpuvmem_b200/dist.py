"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch) for the one exchange step the path has.

The reference shards by frequency channel only (``gpu_idx = i % num_gpus``, reference
``src/functions.cu:4341``) and "reduces" by peer-to-peer stores into GPU 0 under a global
lock (``:4534-4549``). Here each rank owns either whole channels (same ``i % world`` rule)
or, when there are fewer channels than ranks, a contiguous chunk of every channel's
visibilities (chi2 and its gradient are plain sums over visibilities). Every rank keeps a
replica of the image and accumulates its partial gradient locally; ONE all-reduce of
``[2*M*N gradient | chi2]`` per evaluation replaces the serialised P2P accumulate.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(gpus_requested=1):
    """RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* come from torchrun. Returns (rank, world, local)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_plan(nchan, nvis_per_chan, world):
    """Per-rank keyword arguments for ``Engine.from_problem``.

    * ``nchan >= world``: channel ``i`` goes to rank ``i % world`` (the reference's rule).
    * otherwise: every channel is cut into ``world`` contiguous visibility chunks.
    """
    if world <= 1:
        return [dict()]
    if nchan >= world:
        return [dict(channels=[c for c in range(nchan) if c % world == r]) for r in range(world)]
    zmax = max(nvis_per_chan)
    per = -(-zmax // world)
    return [dict(vis_slice=(r * per, min((r + 1) * per, zmax))) for r in range(world)]


def split_f64(x64):
    """A float64 scalar tensor as two float32 (hi, lo) so it can ride in the fp32 gradient
    buffer through the same all-reduce; ``join_f64`` undoes it after the sum."""
    hi = x64.to(torch.float32)
    lo = (x64 - hi.to(torch.float64)).to(torch.float32)
    return torch.cat([hi.reshape(1), lo.reshape(1)])


def join_f64(pair):
    return pair[0].to(torch.float64) + pair[1].to(torch.float64)


def allreduce_eval(buf):
    """Sum ``[gradient | chi2_hi | chi2_lo]`` over ranks in place (no-op for one rank)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(buf)
    return buf
