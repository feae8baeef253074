"""SyncBatchNorm for the small [B, F] activations of the classification head, over the same
statistics exchange the EdgeConv layers use (ops._allreduce_stats: one push kernel over NVLink
peer memory, or the process group's all-reduce).

torch.nn.SyncBatchNorm (main_partseg_dist.py:189 converts every BatchNorm) issues one
all_gather in the forward and one all_reduce in the backward per layer through NCCL; for the
head's bn6 / bn7 those are four latency-bound library collectives per step on [32, 512]
tensors.  Here the exchange is the fp64 vector [sum x | sum x^2 | count] (forward) and
[sum g | sum g*xhat] (backward), exactly the protocol of the fused EdgeConv BatchNorm, so the
whole step uses one transport.  Semantics are nn.SyncBatchNorm's: global batch statistics in
training mode, unbiased variance into running_var, num_batches_tracked += 1.
"""
from __future__ import annotations

from ctypes import c_void_p

import torch
import torch.nn as nn

from . import _lib, ops


class _SyncBNRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, momentum, eps, group):
        x = x.contiguous().float()
        M, F = x.shape
        dev = x.device
        with torch.cuda.device(dev):
            st = ops._stream(x)
            stats = torch.zeros(2 * F + 1, device=dev, dtype=torch.float64)
            affine = torch.empty(4, F, device=dev, dtype=torch.float32)
            mean, invstd, a, b = (c_void_p(affine.data_ptr() + 4 * F * r) for r in range(4))
            _lib.call("ecb200_colstats", ops._ptr(x), M, F, ops._ptr(stats), st)
            g32, b32 = gamma.detach().contiguous().float(), beta.detach().contiguous().float()
            ops._stats_to_affine(stats, group, g32, b32, eps, F, mean, invstd, a, b, st)   # exchange + finalize
            if running_mean is not None:
                _lib.call("ecb200_bn_update_running", ops._ptr(stats), F, float(momentum), ops._ptr(running_mean),
                          ops._ptr(running_var), ops._ptr(nbt), st)
        out = torch.addcmul(affine[3], x, affine[2])
        ctx.save_for_backward(x, affine, stats)
        ctx.group = group
        return out

    @staticmethod
    def backward(ctx, g):
        x, affine, stats = ctx.saved_tensors
        M, F = x.shape
        dev = x.device
        g = g.contiguous().float()
        with torch.cuda.device(dev):
            st = ops._stream(x)
            mean, invstd, a = (c_void_p(affine.data_ptr() + 4 * F * r) for r in range(3))
            total = torch.empty(2 * F, device=dev, dtype=torch.float64)
            dgamma = torch.empty(F, device=dev, dtype=torch.float32)
            dbeta = torch.empty(F, device=dev, dtype=torch.float32)
            dx = torch.empty_like(x)
            # local sums (also the parameter gradients) -> exchange -> dx: three launches
            _lib.call("ecb200_rows_bn_bwd_stats", ops._ptr(g), ops._ptr(x), mean, invstd, M, F, ops._ptr(total),
                      ops._ptr(dgamma), ops._ptr(dbeta), st)
            if ctx.group:
                ops._allreduce_stats(total, ctx.group)
            count = c_void_p(stats.data_ptr() + 8 * 2 * F)          # stats[2F] = global row count (fp64)
            _lib.call("ecb200_rows_bn_bwd_dx", ops._ptr(g), ops._ptr(x), mean, invstd, a, ops._ptr(total), count,
                      M, F, ops._ptr(dx), st)
        return dx, dgamma, dbeta, None, None, None, None, None, None


def batch_norm_rows(x: torch.Tensor, bn: nn.Module) -> torch.Tensor:
    """``bn(x)`` for x [B, F]; when ``bn`` is a SyncBatchNorm in training mode inside an
    initialised process group, the statistics travel over the EdgeConv statistics exchange
    instead of torch's NCCL all_gather / all_reduce."""
    from .dgcnn import _sync_group
    group = _sync_group(bn) if x.is_cuda and x.dim() == 2 else 0
    if not group:
        return bn(x)
    mom = -1.0 if bn.momentum is None else float(bn.momentum)
    track = bn.track_running_stats and bn.running_mean is not None
    return _SyncBNRows.apply(x, bn.weight, bn.bias, bn.running_mean if track else None,
                             bn.running_var if track else None, bn.num_batches_tracked if track else None,
                             mom, bn.eps, group)
