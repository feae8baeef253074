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


class FlatGradSync:
    """Data-parallel gradient averaging over a flat fp32 buffer, overlapped with the backward.

    The parameters are laid out in one flat buffer in two buckets: ``late`` (those whose gradients
    are produced at the very end of the backward pass -- the EdgeConv layers, which come first in the
    network) and ``early`` (conv5 and the head: 95 % of the bytes, complete after the first few per
    cent of the backward).  Autograd produces each gradient in its own tensor (``.grad`` is None at
    the start of a step, so nothing is accumulated in place: no read-modify-write kernel per
    parameter); a bucket is packed into the flat buffer by ONE multi-tensor copy as soon as its last
    gradient exists.  The early bucket's all-reduce is issued on a side stream and runs under the
    EdgeConv backward; ``average()`` packs and reduces the small late bucket, joins the side stream
    and points every ``.grad`` at its (averaged) slice of the flat buffer for the optimizer.
    No bucketing threads; everything is capturable in a CUDA graph together with the step.
    Usage per step:  sync.zero() ; loss.backward() ; sync.average() ; optimizer.step()
    """

    def __init__(self, params, group=None, late=None, overlap: bool = True):
        """``late``: optional predicate(param_index, param) -> bool selecting the late bucket
        (default: the first 12 parameters = conv1..conv4 + their BatchNorms of a DGCNN)."""
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        if late is None:
            late = lambda i, p: i < 12  # noqa: E731
        self.late_ps = [p for i, p in enumerate(self.params) if late(i, p)]
        self.early_ps = [p for i, p in enumerate(self.params) if not late(i, p)]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.view = {}
        o = 0
        for p in self.late_ps + self.early_ps:
            self.view[p] = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()
        n_late = sum(p.numel() for p in self.late_ps)
        self.flat_late, self.flat_early = self.flat[:n_late], self.flat[n_late:]
        # NCCL averages inside the collective (no separate scaling kernel); other backends sum, then scale
        self._avg = bool(self.world > 1 and dist.get_backend(group) == "nccl" and hasattr(dist.ReduceOp, "AVG"))
        self.overlap = bool(overlap and self.world > 1 and self.early_ps and dev.type == "cuda")
        self._pending = len(self.early_ps)
        self._left = self._pending
        self._early_done = False
        self._side = torch.cuda.Stream(device=dev) if self.overlap else None
        if self.overlap:
            for p in self.early_ps:
                p.register_post_accumulate_grad_hook(self._on_early_grad)
        self.zero()

    def _pack(self, ps) -> None:
        """gradients of ``ps`` -> their slices of the flat buffer (one multi-tensor copy); a parameter
        that received no gradient contributes zeros"""
        have = [p for p in ps if p.grad is not None and p.grad.data_ptr() != self.view[p].data_ptr()]
        if have:
            # flattened on both sides: a 1x1-conv weight gradient in channels-last strides has the same element
            # order as the contiguous slice, and differing strides would push the multi-tensor copy onto its
            # slow per-tensor path (one un-vectorised 2 MB copy for conv5's weight)
            torch._foreach_copy_([self.view[p].view(-1) for p in have], [p.grad.reshape(-1) for p in have])
        for p in ps:
            if p.grad is None:
                self.view[p].zero_()

    def _on_early_grad(self, _param) -> None:
        self._left -= 1
        if self._left == 0:
            self._pack(self.early_ps)
            cur = torch.cuda.current_stream()
            self._side.wait_stream(cur)
            with torch.cuda.stream(self._side):
                self._reduce_mean(self.flat_early)
            self._early_done = True

    def _reduce_mean(self, buf: torch.Tensor) -> None:
        if self._avg:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.mul_(1.0 / self.world)

    def zero(self) -> None:
        """start of a step: every ``.grad`` is None, so autograd stores (not accumulates) gradients"""
        for p in self.params:
            p.grad = None
        self._left = self._pending
        self._early_done = False

    def average(self) -> None:
        if self._early_done:
            self._pack(self.late_ps)
            if self.flat_late.numel():
                self._reduce_mean(self.flat_late)
            torch.cuda.current_stream().wait_stream(self._side)
        else:   # no overlap (or a hook did not fire): pack and reduce everything here
            self._pack(self.params)
            if self.world > 1:
                self._reduce_mean(self.flat)
        for p in self.params:
            p.grad = self.view[p]


class PeerStatsExchange:
    """The SyncBatchNorm statistics exchange as ONE kernel over NVLink peer memory.

    The exchange is a few hundred fp64 values per BatchNorm layer, ten times per step: pure
    latency.  Instead of an NCCL collective every rank pushes its vector straight into a slot
    of every peer's buffer (plain stores over NVLink into memory mapped with
    torch.distributed._symmetric_memory) as 8-byte {32 data bits | 32-bit sequence number} words
    -- no flag, no fence: an aligned 8-byte store arrives atomically, so a word that carries the
    expected sequence number also carries its data -- and sums the vectors that arrive in its own
    buffer in rank order (csrc/peer_exchange.cu).  No host involvement, capturable in a CUDA
    graph, bit-identical results on every rank.

        ex = PeerStatsExchange.enable()        # after init_process_group, once per process
    """

    def __init__(self, group=None):
        import torch.distributed._symmetric_memory as symm_mem

        from . import _lib
        self.group = dist.group.WORLD if group is None else group
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        lib = _lib.load()
        nbytes = lib.ecb200_peer_buffer_bytes(self.world)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.bufs_dev = int(self.handle.buffer_ptrs_dev)     # device array of `world` base pointers
        self.seq = torch.zeros(1, dtype=torch.int64, device=dev)
        self.max_values = nbytes // (2 * self.world * 16)
        torch.cuda.synchronize()
        dist.barrier(self.group)                              # every buffer is zeroed before any push

    @classmethod
    def enable(cls, group=None) -> "PeerStatsExchange":
        """Create the exchange for ``group`` (default: WORLD) and route the EdgeConv / embed_pool
        BatchNorm statistics of that group through it."""
        from . import ops
        ex = cls(group)
        ops.set_peer_exchange(ops.register_group(ex.group), ex)
        return ex

    @classmethod
    def enable_collectively(cls, group=None) -> str:
        """``enable()`` on every rank of ``group``, then agree on the outcome: unless the symmetric-
        memory setup succeeded on ALL ranks the exchange is switched off everywhere and the
        statistics travel by the group's all-reduce (a rank pushing into peer memory while another
        waits in an NCCL all-reduce would hang both).  Returns a description of the transport."""
        from . import ops
        err = ""
        try:
            cls.enable(group)
        except Exception as exc:  # noqa: BLE001 - any failure selects the NCCL transport
            err = f"{type(exc).__name__}: {exc}"
        g = dist.group.WORLD if group is None else group
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(g) == "nccl" else "cpu"
        okf = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(okf, op=dist.ReduceOp.MIN, group=g)
        if int(okf.item()) == 1:
            return "one-kernel push exchange over NVLink peer memory (symmetric memory)"
        ops.set_peer_exchange(ops.register_group(g), None)
        return ("NCCL all-reduce (peer memory unavailable on at least one rank"
                + (f"; here: {err}" if err else "") + ")")[:200]

    def allreduce_(self, stats: torch.Tensor) -> torch.Tensor:
        """In-place SUM of an fp64 device vector over the group (the op the kernels use)."""
        from ctypes import c_void_p

        from . import _lib
        if stats.dtype != torch.float64 or not stats.is_cuda or not stats.is_contiguous():
            raise TypeError("PeerStatsExchange.allreduce_ takes a contiguous fp64 CUDA tensor")
        if stats.numel() > self.max_values:
            raise ValueError(f"at most {self.max_values} values per exchange, got {stats.numel()}")
        _lib.call("ecb200_peer_allreduce", c_void_p(stats.data_ptr()), stats.numel(), c_void_p(self.bufs_dev),
                  self.rank, self.world, c_void_p(self.seq.data_ptr()),
                  c_void_p(torch.cuda.current_stream().cuda_stream))
        return stats
