"""Data-parallel plumbing of the Stage-1 step: one process per GPU, batch split, ONE all-reduce of the flat
gradient buffer per step (the reference's DDP bucketing + SyncBatchNorm collectives of train_stage1.py:68-70 are
consciously replaced: BatchNorm statistics stay per rank = the published single-GPU recipe on every rank).

Device-agnostic host logic (NCCL on the GPUs, gloo in the CPU tests): nothing here touches libtris_sm100.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def world_size(pg=None) -> int:
    return dist.get_world_size(pg) if (dist.is_available() and dist.is_initialized()) else 1


def shard_seed(base: int, rank: int, index: int = 0) -> int:
    """Seed of the `index`-th synthetic batch of `rank` (reference seeds 1234, train_stage1.py:34-41; ranks draw disjoint
    streams exactly like DistributedSampler hands them disjoint samples, :107-111)."""
    return base + 1000 * rank + index


def shard_range(n_items: int, rank: int, world: int) -> range:
    """Round-robin shard of independent work items (validation refs, SURVEY 8e 'replicas only')."""
    return range(rank, n_items, world)


def broadcast_parameters(flat: torch.Tensor, pg=None, src: int = 0) -> None:
    """Rank `src`'s parameters to every rank (what DistributedDataParallel does at construction)."""
    if world_size(pg) > 1:
        dist.broadcast(flat, src=src, group=pg)


def all_reduce_gradients(flat_grad: torch.Tensor, pg=None) -> float:
    """SUM all-reduce of the flat gradient buffer in place; returns the factor the optimizer must apply (1/world) so
    that the averaging costs no extra pass (it is folded into the fused AdamW kernel)."""
    w = world_size(pg)
    if w > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=pg)
    return 1.0 / w


def max_over_ranks(value: float, device, pg=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if world_size(pg) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=pg)
    return float(t.item())
