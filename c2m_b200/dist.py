"""Multi-GPU plumbing for the warp path: one process per GPU, work partitioned by batch x frame.

Every (n) slice of the folded batch x frame axis is independent in forward and backward
(the reference folds T into the batch at src/modules/model.py:196-202 and
shards the batch with DistributedSampler, src/train.py:58-60), so the op needs no collective.
torch.distributed is used only to agree on timings / counts (and, in a training step, for the DDP
gradient all-reduce of the surrounding network, which has nothing to do with this op).
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_global: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [begin, end) block of the batch x frame axis owned by `rank`; the first
    n_global % world ranks get one extra frame."""
    if world <= 0 or not (0 <= rank < world) or n_global < 0:
        raise ValueError(f"bad shard request n={n_global} rank={rank} world={world}")
    base, extra = divmod(n_global, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_world() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1-process default)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def init(backend: str | None = None) -> Tuple[int, int, int]:
    """Initialise the default process group when launched under torchrun (env:// rendezvous)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def _reduce_device(device):
    """Where the one-element reduce tensor lives: NCCL reduces CUDA tensors only (the rank's current device), gloo
    takes CPU tensors."""
    if device is not None:
        return device
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return "cpu"


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank scalar (device-timed milliseconds) over all ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_reduce_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_reduce_device(device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def aggregate_throughput(frames_this_rank: int, ms_this_rank: float, device=None) -> Tuple[float, float, int]:
    """Whole-job frames/s = frames over all ranks / max-over-ranks time. Returns (frames_per_s,
    max_ms, total_frames)."""
    total = int(round(sum_over_ranks(frames_this_rank, device)))
    ms = max_over_ranks(ms_this_rank, device)
    return (total / (ms * 1e-3) if ms > 0 else 0.0), ms, total
