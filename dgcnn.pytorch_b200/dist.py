"""Multi-GPU plumbing: one process per GPU, clouds sharded over the ranks.

Every cloud's kNN graph, gather and max are independent of every other cloud
(models/dgcnn.py:22-25 keeps gathers inside a cloud), so the batch dimension shards
with no data-path collective.  The only exchange on the EdgeConv path is the
training-mode BatchNorm statistics under SyncBatchNorm (main_partseg_dist.py:189):
  forward : all-reduce(SUM) of [sum e (Co), sum e^2 (Co), edge count (1)]  fp64
  backward: all-reduce(SUM) of [sum g (Co), sum g*xhat (Co)]              fp64
issued by ops.edgeconv_fwd_op / edgeconv_bwd_op on the compute stream between the gather
kernel and the finalize kernel.  Weights are replicated; DDP averages their gradients.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[begin, end) of the contiguous shard of ``n_items`` clouds owned by ``rank``
    (DistributedSampler-style: sizes differ by at most one, early ranks get the extra)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, extra = divmod(n_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from torchrun's environment; initialises the default
    process group when WORLD_SIZE > 1 (nccl on GPUs, gloo otherwise)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def allreduce_stats(stats: torch.Tensor, group=None) -> torch.Tensor:
    """The SyncBatchNorm exchange: element-wise SUM of the fp64 statistics buffer over the
    group, in place.  ``stats`` = [sum e (Co) | sum e^2 (Co) | count] forward, or
    [sum g (Co) | sum g*xhat (Co)] backward."""
    if stats.dtype != torch.float64:
        raise TypeError("BatchNorm statistics travel in fp64")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats
