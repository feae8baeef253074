"""torch.library custom ops over the C ABI (include/edgeconv_b200.h).

PyTorch is plumbing here: it owns device memory and the stream; every piece of
arithmetic on the EdgeConv path is a kernel of libedgeconv_b200.so.  CPU tensors
are rejected -- the CPU path is the oracle under oracle/, which this package never
imports.

Ops (namespace ``edgeconv_b200``):
  knn(x, k) -> int32 [B,N,k]                       reference knn(), models/dgcnn.py:6-12
  graph_feature(x, idx, mode) -> edge tensor       get_graph_feature(), dgcnn.py:15-44
  edgeconv_fwd(...) -> out [B,Co,N] (+ saved)      conv{n} + max over k, dgcnn.py:54-73,:84-98
  edgeconv_bwd(...) -> dx, dW, dgamma, dbeta       their autograd
"""
from __future__ import annotations

import os
from ctypes import c_void_p
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib

GF_CONCAT, GF_KNN_ONLY, GF_DISP_ONLY, GF_CONCAT_CENTERED = 0, 1, 2, 3
MAX_K = 64
AMAX_SLOTS = 32          # ECB200_AMAX_SLOTS

# torch.distributed process groups cannot travel through an op schema: ops take an
# integer handle into this table (0 = no cross-rank BatchNorm statistics).
_GROUPS = {}
_PEER = {}      # group handle -> dist.PeerStatsExchange (statistics exchange over NVLink peer memory)


def register_group(group) -> int:
    """Handle for a process group whose ranks share BatchNorm statistics
    (SyncBatchNorm semantics, main_partseg_dist.py:189)."""
    if group is None:
        group = dist.group.WORLD
    for h, g in _GROUPS.items():
        if g is group:
            return h
    h = len(_GROUPS) + 1
    _GROUPS[h] = group
    return h


def set_peer_exchange(handle: int, exchange) -> None:
    """Route the BatchNorm-statistics all-reduce of group ``handle`` through a one-kernel exchange
    over peer memory (dist.PeerStatsExchange) instead of NCCL; None switches back."""
    if exchange is None:
        _PEER.pop(handle, None)
    else:
        _PEER[handle] = exchange


def _allreduce_stats(stats: Tensor, handle: int) -> None:
    """In-place SUM of the fp64 statistics vector over the ranks of group ``handle``."""
    peer = _PEER.get(handle)
    if peer is not None and stats.numel() <= peer.max_values:
        _lib.call("ecb200_peer_allreduce", _ptr(stats), stats.numel(), c_void_p(peer.bufs_dev), peer.rank,
                  peer.world, _ptr(peer.seq), _stream(stats))
    else:
        dist.all_reduce(stats, group=_GROUPS[handle])


def _stats_to_affine(stats: Tensor, handle: int, gamma: Tensor, beta: Tensor, eps: float, Co: int, mean, invstd,
                     a, b, st) -> None:
    """training-mode BatchNorm: (cross-rank SUM of ``stats`` under group ``handle``, then) the folded
    affine.  With the peer exchange the all-reduce and the finalize are ONE kernel."""
    peer = _PEER.get(handle) if handle else None
    if peer is not None and stats.numel() <= peer.max_values:
        _lib.call("ecb200_peer_allreduce_bn_finalize", _ptr(stats), Co, c_void_p(peer.bufs_dev), peer.rank,
                  peer.world, _ptr(peer.seq), _ptr(gamma), _ptr(beta), float(eps), mean, invstd, a, b, st)
        return
    if handle:
        dist.all_reduce(stats, group=_GROUPS[handle])
    _lib.call("ecb200_bn_finalize", _ptr(stats), _ptr(gamma), _ptr(beta), None, None, 1, float(eps), Co, mean,
              invstd, a, b, st)


def _bstats_to_coeffs(bstats: Tensor, handle: int, count_ptr, a, invstd, Co: int, dgamma: Tensor, dbeta: Tensor,
                      c1, c2, st) -> None:
    """training-mode BatchNorm backward: local [sum g | sum g*xhat] -> dgamma, dbeta (local) and the
    c1, c2 coefficients from the cross-rank sums; with the peer exchange in ONE kernel."""
    peer = _PEER.get(handle) if handle else None
    if handle:
        bglobal = bstats.clone()
        if peer is not None and bglobal.numel() <= peer.max_values:
            _lib.call("ecb200_peer_allreduce_bwd_finalize", _ptr(bstats), _ptr(bglobal), Co, c_void_p(peer.bufs_dev),
                      peer.rank, peer.world, _ptr(peer.seq), count_ptr, a, invstd, _ptr(dgamma), _ptr(dbeta),
                      c1, c2, st)
            return
        dist.all_reduce(bglobal, group=_GROUPS[handle])
    else:
        bglobal = bstats
    _lib.call("ecb200_bwd_finalize", _ptr(bstats), _ptr(bglobal), count_ptr, a, invstd, 1, Co, _ptr(dgamma),
              _ptr(dbeta), c1, c2, st)


def _ptr(t: Optional[Tensor]):
    return None if t is None else c_void_p(t.data_ptr())


def _stream(t: Tensor):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _check_cuda_f32(name: str, t: Tensor, dims: Optional[int] = None):
    if not t.is_cuda:
        raise RuntimeError(f"edgeconv_b200: {name} must be a CUDA tensor (no CPU fallback; "
                           f"got device {t.device})")
    if t.dtype != torch.float32:
        raise RuntimeError(f"edgeconv_b200: {name} must be float32, got {t.dtype}")
    if dims is not None and t.dim() != dims:
        raise ValueError(f"edgeconv_b200: {name} must have {dims} dimensions, got shape "
                         f"{tuple(t.shape)}")


# ------------------------------------------------------------------------------ kNN
@torch.library.custom_op("edgeconv_b200::knn", mutates_args=(), device_types="cuda")
def knn_op(x: Tensor, k: int, sorted: bool = True) -> Tensor:
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    if k > N or k < 1:
        # same trigger as Tensor.topk in the reference (dgcnn.py:11)
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")
    if k > MAX_K:
        raise RuntimeError(f"edgeconv_b200: k={k} exceeds the selector limit {MAX_K}")
    x = x.contiguous()
    kind = knn_tc_kind(C, N, k)
    if kind == "f16":
        # feature-space layers: tcgen05 / TMA distance tiles (knn_tc.cu), packed fp16 halves
        hh, hl, nb, xxs, cmax = split_f16_op(x, False)[:5]
        return knn_tc_f16_op(hh, hl, nb, xxs, cmax, B, N, k)
    if kind == "tf32":
        hi, lo, xx = split_tf32_op(x)
        return knn_tc_op(hi, lo, xx, B, N, k)
    if kind == "xyz":
        return knn_tc_xyz_op(x, k)
    with torch.cuda.device(x.device):
        # xyz layer (and any shape the tensor-core kernel does not take): FP32 FMA tiles
        xx = torch.empty(B * N, device=x.device, dtype=torch.float32)
        idx = torch.empty(B, N, k, device=x.device, dtype=torch.int32)
        st = _stream(x)
        _lib.call("ecb200_sqnorms", _ptr(x), B, C, N, _ptr(xx), st)
        _lib.call("ecb200_knn", _ptr(x), _ptr(xx), B, C, N, k, int(sorted), _ptr(idx), st)
    return idx


@torch.library.custom_op("edgeconv_b200::split_tf32", mutates_args=(), device_types="cuda")
def split_tf32_op(x: Tensor) -> List[Tensor]:
    """x [B,C,N] -> [hi [B*N,C], lo [B*N,C], xx [B*N]]: the point-major tf32 operand pair shared
    by the tensor-core kNN and the tensor-core point GEMM of a layer, and |x|^2."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    x = x.contiguous()
    with torch.cuda.device(x.device):
        hi = torch.empty(B * N, C, device=x.device, dtype=torch.float32)
        lo = torch.empty(B * N, C, device=x.device, dtype=torch.float32)
        xx = torch.empty(B * N, device=x.device, dtype=torch.float32)
        _lib.call("ecb200_split_tf32", _ptr(x), B, C, N, _ptr(hi), _ptr(lo), _ptr(xx), _stream(x))
    return [hi, lo, xx]


@split_tf32_op.register_fake
def _(x):
    B, C, N = x.shape
    return [x.new_empty((B * N, C)), x.new_empty((B * N, C)), x.new_empty((B * N,))]


@torch.library.custom_op("edgeconv_b200::knn_tc", mutates_args=(), device_types="cuda")
def knn_tc_op(hi: Tensor, lo: Tensor, xx: Tensor, B: int, N: int, k: int) -> Tensor:
    """kNN graph from the split operands: int32 [B,N,k], nearest first."""
    C = hi.shape[1]
    if k > N or k < 1:
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")
    with torch.cuda.device(hi.device):
        idx = torch.empty(B, N, k, device=hi.device, dtype=torch.int32)
        nbytes = _lib.load().ecb200_knn_tc_workspace_bytes(B, N, k)
        ws = torch.empty(nbytes, device=hi.device, dtype=torch.uint8)
        _lib.call("ecb200_knn_tc", _ptr(hi), _ptr(lo), _ptr(xx), B, C, N, k, 1, _ptr(idx), _ptr(ws), nbytes,
                  _stream(hi))
    return idx


@knn_tc_op.register_fake
def _(hi, lo, xx, B, N, k):
    return hi.new_empty((B, N, k), dtype=torch.int32)


# ---- packed-FP16 operands (kind::f16): same 11-bit significands as tf32 at twice the MMA rate ----
@torch.library.custom_op("edgeconv_b200::split_f16", mutates_args=(), device_types="cuda")
def split_f16_op(x: Tensor, with_tf32: bool, amax: Optional[Tensor] = None) -> List[Tensor]:
    """x [B,C,N] -> [hh, hl (fp16 [B*N,C]), nb (fp16 [B*N,64]), xxs [B*N], cmax [B], hi, lo (tf32-valued
    fp32 [B*N,C]), xx [B*N]].  hh/hl/nb/xxs/cmax are the operands of knn_tc_f16_op (the tensor scaled by a
    power of two into fp16's range; nb = the candidate rows of the folded -0.5|x_j|^2 term, cmax = the
    clouds' largest scaled squared norms); hi/lo/xx those of the tensor-core GEMMs, from the same pass
    (empty unless with_tf32).
    ``amax`` [AMAX_SLOTS]: max |x| when the producer of x already knows it (ecb200_edge_apply_amax)."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    x = x.contiguous()
    dev = x.device
    with torch.cuda.device(dev):
        st = _stream(x)
        hh = torch.empty(B * N, C, device=dev, dtype=torch.float16)
        hl = torch.empty(B * N, C, device=dev, dtype=torch.float16)
        nb = torch.empty(B * N, 64, device=dev, dtype=torch.float16)
        xxs = torch.empty(B * N, device=dev, dtype=torch.float32)
        cmax = torch.empty(B * ((N + 31) // 32), device=dev, dtype=torch.float32)
        hi = lo = xx = None
        if with_tf32:
            hi = torch.empty(B * N, C, device=dev, dtype=torch.float32)
            lo = torch.empty(B * N, C, device=dev, dtype=torch.float32)
            xx = torch.empty(B * N, device=dev, dtype=torch.float32)
        if amax is None:
            amax = torch.empty(AMAX_SLOTS, device=dev, dtype=torch.float32)
            _lib.call("ecb200_absmax", _ptr(x), x.numel(), _ptr(amax), st)
        _lib.call("ecb200_split_f16", _ptr(x), B, C, N, _ptr(amax), _ptr(hh), _ptr(hl), _ptr(nb), _ptr(xxs),
                  _ptr(cmax), _ptr(hi), _ptr(lo), _ptr(xx), st)
    if not with_tf32:
        hi, lo, xx = xxs.new_empty(0), xxs.new_empty(0), xxs.new_empty(0)
    return [hh, hl, nb, xxs, cmax, hi, lo, xx]


@split_f16_op.register_fake
def _(x, with_tf32, amax=None):
    B, C, N = x.shape
    h = x.new_empty((B * N, C), dtype=torch.float16)
    f = x.new_empty((B * N, C)) if with_tf32 else x.new_empty(0)
    v = x.new_empty((B * N,)) if with_tf32 else x.new_empty(0)
    return [h, x.new_empty((B * N, C), dtype=torch.float16), x.new_empty((B * N, 64), dtype=torch.float16),
            x.new_empty((B * N,)), x.new_empty((B * ((N + 31) // 32),)), f,
            x.new_empty((B * N, C)) if with_tf32 else x.new_empty(0), v]


@torch.library.custom_op("edgeconv_b200::knn_tc_f16", mutates_args=(), device_types="cuda")
def knn_tc_f16_op(hh: Tensor, hl: Tensor, nb: Tensor, xxs: Tensor, cmax: Tensor, B: int, N: int, k: int) -> Tensor:
    """kNN graph from the packed fp16 operands: int32 [B,N,k], nearest first."""
    C = hh.shape[1]
    if k > N or k < 1:
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")
    with torch.cuda.device(hh.device):
        idx = torch.empty(B, N, k, device=hh.device, dtype=torch.int32)
        _lib.call("ecb200_knn_tc_f16", _ptr(hh), _ptr(hl), _ptr(nb), _ptr(xxs), _ptr(cmax), B, C, N, k, _ptr(idx),
                  None, _stream(hh))
    return idx


@knn_tc_f16_op.register_fake
def _(hh, hl, nb, xxs, cmax, B, N, k):
    return hh.new_empty((B, N, k), dtype=torch.int32)


@torch.library.custom_op("edgeconv_b200::knn_tc_xyz", mutates_args=(), device_types="cuda")
def knn_tc_xyz_op(x: Tensor, k: int) -> Tensor:
    """kNN of a low-dimensional cloud (C <= 5, the xyz layer) on the tensor-core pipeline: the
    three terms of the compensated product of a point sit in ONE 16-deep K step."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    if k > N or k < 1:
        raise RuntimeError(f"selected index k out of range (k={k}, N={N})")
    x = x.contiguous()
    dev = x.device
    with torch.cuda.device(dev):
        st = _stream(x)
        rows = torch.empty(2, B * N, 64, device=dev, dtype=torch.float16)
        xxs = torch.empty(B * N, device=dev, dtype=torch.float32)
        cmax = torch.empty(B * ((N + 31) // 32), device=dev, dtype=torch.float32)
        idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
        _lib.call("ecb200_pack_xyz_f16", _ptr(x), B, C, N, _ptr(rows[0]), _ptr(rows[1]), _ptr(xxs), _ptr(cmax), st)
        _lib.call("ecb200_knn_tc_xyz", _ptr(rows[0]), _ptr(rows[1]), _ptr(xxs), _ptr(cmax), B, N, k, _ptr(idx),
                  None, st)
    return idx


@knn_tc_xyz_op.register_fake
def _(x, k):
    B, C, N = x.shape
    return x.new_empty((B, N, k), dtype=torch.int32)


def knn_tc_kind(C: int, N: int, k: int) -> str:
    """Which tensor-core kernel knn() uses: "f16" (packed fp16 halves, C = 64 or 128), "tf32"
    (the other multiples of 32 up to 128), "xyz" (C <= 4, N >= 64) or "" (FP32-FMA kernel).  ECB200_KNN=fma|tf32 and
    ECB200_KNN_XYZ=fma override (A/B tests)."""
    mode = os.environ.get("ECB200_KNN", "auto")
    if mode == "fma" or k > 40:
        return ""
    if C <= 4:
        return "xyz" if (N >= 64 and os.environ.get("ECB200_KNN_XYZ", "tc") == "tc") else ""
    if C in (64, 128) and mode != "tf32":
        return "f16"
    if C % 32 == 0 and 32 <= C <= 128:
        return "tf32"
    return ""


def knn_uses_tensor_cores(C: int, N: int, k: int) -> bool:
    """Kernel choice for knn(): tensor cores where the contraction is a real GEMM
    (C a multiple of 32 in [32, 128], k <= 40), FP32 FMA otherwise (any C, k <= 64).
    ECB200_KNN=fma forces the FMA kernel (A/B tests)."""
    return knn_tc_kind(C, N, k) in ("f16", "tf32")


def debug_tc_scores_f16(x: Tensor) -> Tensor:
    """Diagnostic: scores x_i.x_j - 0.5|x_j|^2 [B,N,N] with the dot products from the packed-fp16
    three-term pipeline (unscaled; the column term is added here in fp64-free torch arithmetic)."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    x = x.contiguous()
    with torch.cuda.device(x.device):
        hh, hl, nb, xxs, cmax = split_f16_op(x, False)[:5]
        out = torch.full((B, N, N), float("nan"), device=x.device, dtype=torch.float32)
        _lib.call("ecb200_debug_tc_scores_f16", _ptr(hh), _ptr(hl), B, C, N, _ptr(out), _stream(x))
        xx = (x.double() ** 2).sum(1)
        s2 = (xxs.view(B, N).double().sum() / xx.sum().clamp_min(1e-300)).item() if float(xx.sum()) > 0 else 1.0
    return (out.double() / s2 - 0.5 * xx[:, None, :]).float()


def debug_tc_scores(x: Tensor) -> Tensor:
    """Diagnostic: scores x_i.x_j - 0.5|x_j|^2 [B,N,N] from the tensor-core pipeline."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    x = x.contiguous()
    with torch.cuda.device(x.device):
        hi = torch.empty(B * N, C, device=x.device, dtype=torch.float32)
        lo = torch.empty_like(hi)
        xx = torch.empty(B * N, device=x.device, dtype=torch.float32)
        out = torch.full((B, N, N), float("nan"), device=x.device, dtype=torch.float32)
        st = _stream(x)
        _lib.call("ecb200_split_tf32", _ptr(x), B, C, N, _ptr(hi), _ptr(lo), _ptr(xx), st)
        _lib.call("ecb200_debug_tc_scores", _ptr(hi), _ptr(lo), _ptr(xx), B, C, N, _ptr(out), st)
    return out


@knn_op.register_fake
def _(x, k, sorted=True):
    B, C, N = x.shape
    return x.new_empty((B, N, k), dtype=torch.int32)


# ------------------------------------------------------------------- graph feature
def _gf_shape(B, C, N, k, mode):
    if mode == GF_KNN_ONLY:
        return (B, N, k, C)
    if mode == GF_DISP_ONLY:
        return (B, C, N, k)
    return (B, 2 * C, N, k)


@torch.library.custom_op("edgeconv_b200::graph_feature", mutates_args=(), device_types="cuda")
def graph_feature_op(x: Tensor, idx: Tensor, mode: int) -> Tensor:
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    k = idx.shape[-1]
    x = x.contiguous()
    idx = idx.contiguous()
    assert idx.dtype == torch.int32 and tuple(idx.shape) == (B, N, k)
    with torch.cuda.device(x.device):
        out = torch.empty(_gf_shape(B, C, N, k, mode), device=x.device, dtype=torch.float32)
        _lib.call("ecb200_graph_feature", _ptr(x), _ptr(idx), B, C, N, k, mode, _ptr(out), _stream(x))
    return out


@graph_feature_op.register_fake
def _(x, idx, mode):
    B, C, N = x.shape
    return x.new_empty(_gf_shape(B, C, N, idx.shape[-1], mode))


@torch.library.custom_op("edgeconv_b200::graph_feature_bwd", mutates_args=(), device_types="cuda")
def graph_feature_bwd_op(gout: Tensor, idx: Tensor, C: int, mode: int) -> Tensor:
    B, N, k = idx.shape
    gout = gout.contiguous().float()
    with torch.cuda.device(gout.device):
        dx = torch.empty(B, C, N, device=gout.device, dtype=torch.float32)
        _lib.call("ecb200_graph_feature_bwd", _ptr(gout), _ptr(idx.contiguous()), B, C, N, k, mode,
                  _ptr(dx), _stream(gout))
    return dx


@graph_feature_bwd_op.register_fake
def _(gout, idx, C, mode):
    B, N, k = idx.shape
    return gout.new_empty((B, C, N))


def _gf_setup(ctx, inputs, output):
    x, idx, mode = inputs
    ctx.save_for_backward(idx)
    ctx.C = x.shape[1]
    ctx.mode = mode


def _gf_backward(ctx, gout):
    (idx,) = ctx.saved_tensors
    return graph_feature_bwd_op(gout, idx, ctx.C, ctx.mode), None, None


graph_feature_op.register_autograd(_gf_backward, setup_context=_gf_setup)


# ------------------------------------------------------------------ fused EdgeConv
@torch.library.custom_op("edgeconv_b200::edgeconv_fwd", mutates_args=(), device_types="cuda")
def edgeconv_fwd_op(x: Tensor, idx: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor,
                    running_mean: Optional[Tensor], running_var: Optional[Tensor],
                    use_batch_stats: bool, eps: float, slope: float, subtract_center: bool,
                    group: int, save_for_bwd: bool, xhi: Optional[Tensor],
                    xlo: Optional[Tensor], emit_pm: bool) -> List[Tensor]:
    """Returns [out, out_pm, sel, arg, esum, Y, Wcat, affine, stats, wsplit, amax]; ``out_pm`` [B*N, Co] is
    the same output point-major (rows of channels; empty unless ``emit_pm``); ``affine`` is [4,Co] =
    (mean, invstd, a, b); ``amax`` [AMAX_SLOTS] = max |out| (the next layer's operand scale).  Everything
    between ``out_pm`` and ``amax`` exists for the backward pass."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    k = idx.shape[-1]
    Co = weight.shape[0]
    if weight.numel() != Co * 2 * C:
        raise RuntimeError(f"edgeconv_b200: weight {tuple(weight.shape)} does not match 2*C = {2 * C}")
    if Co % 4 != 0:
        raise RuntimeError(f"edgeconv_b200: output channels must be a multiple of 4, got {Co}")
    x = x.contiguous()
    idx = idx.contiguous()
    w = weight.detach().reshape(Co, 2 * C).contiguous().float()
    gamma_c, beta_c = gamma.detach().contiguous().float(), beta.detach().contiguous().float()
    dev = x.device
    M = B * N
    with torch.cuda.device(dev):
        st = _stream(x)
        f32 = dict(device=dev, dtype=torch.float32)
        Wcat = torch.empty(2 * Co, C, **f32)
        Y = torch.empty(M, 2 * Co, **f32)
        sel = torch.empty(M, Co, **f32)
        arg = torch.empty(M, Co, device=dev, dtype=torch.uint8)
        esum = torch.empty(M, Co, **f32) if save_for_bwd else None
        stats = torch.zeros(2 * Co + 1, device=dev, dtype=torch.float64)
        affine = torch.empty(4, Co, **f32)
        mean, invstd, a, b = (c_void_p(affine.data_ptr() + 4 * Co * r) for r in range(4))
        out = torch.empty(B, Co, N, **f32)
        out_pm = torch.empty(M, Co, **f32) if emit_pm else None
        if xhi is not None and xlo is not None and point_gemm_uses_tensor_cores(C):
            # the per-point GEMM on the tensor cores, from the operands the kNN already made; one
            # kernel packs Wcat and produces its tf32 halves for this GEMM and (transposed) for
            # the backward dx GEMM
            wsplit = torch.empty(4, 2 * Co * C, **f32)     # Wcat hi, lo [2Co,C]; Wcat^T hi, lo [C,2Co]
            _lib.call("ecb200_prepare_weights", _ptr(w), Co, C, int(subtract_center), _ptr(Wcat),
                      _ptr(wsplit[0]), _ptr(wsplit[1]), _ptr(wsplit[2]) if save_for_bwd else None,
                      _ptr(wsplit[3]) if save_for_bwd else None, st)
            _lib.call("ecb200_point_gemm_tc", _ptr(xhi), _ptr(xlo), _ptr(wsplit[0]), _ptr(wsplit[1]), M, C,
                      2 * Co, _ptr(Y), st)
        else:
            wsplit = None
            _lib.call("ecb200_pack_weight", _ptr(w), Co, C, int(subtract_center), _ptr(Wcat), st)
            _lib.call("ecb200_point_gemm", _ptr(x), _ptr(Wcat), B, C, N, 2 * Co, _ptr(Y), st)
        _lib.call("ecb200_edge_gather", _ptr(Y), _ptr(idx), _ptr(gamma_c), B, N, k, Co, _ptr(sel),
                  _ptr(arg), _ptr(esum), _ptr(stats) if use_batch_stats else None, st)
        if use_batch_stats:
            # the one exchange step of the path: [sum e, sum e^2, count] over the ranks (group != 0),
            # fused with the finalize when it runs over peer memory
            _stats_to_affine(stats, group, gamma_c, beta_c, eps, Co, mean, invstd, a, b, st)
        else:
            _lib.call("ecb200_bn_finalize", None, _ptr(gamma_c), _ptr(beta_c), _ptr(running_mean),
                      _ptr(running_var), 0, float(eps), Co, mean, invstd, a, b, st)
        amax = torch.empty(AMAX_SLOTS, **f32)
        _lib.call("ecb200_edge_apply_amax", _ptr(sel), a, b, float(slope), B, N, Co, _ptr(out), _ptr(out_pm),
                  Co, _ptr(amax), st)
    if esum is None:
        esum = sel.new_empty(0)
    if out_pm is None:
        out_pm = sel.new_empty(0)
    if wsplit is None:
        wsplit = sel.new_empty(0)
    return [out, out_pm, sel, arg, esum, Y, Wcat, affine, stats, wsplit, amax]


@edgeconv_fwd_op.register_fake
def _(x, idx, weight, gamma, beta, running_mean, running_var, use_batch_stats, eps, slope,
      subtract_center, group, save_for_bwd, xhi, xlo, emit_pm):
    B, C, N = x.shape
    Co = weight.shape[0]
    M = B * N
    f = x.new_empty
    return [f((B, Co, N)), f((M, Co)) if emit_pm else f((0,)), f((M, Co)), f((M, Co), dtype=torch.uint8),
            f((M, Co)) if save_for_bwd else f((0,)), f((M, 2 * Co)), f((2 * Co, C)),
            f((4, Co)), f((2 * Co + 1,), dtype=torch.float64),
            f((4, 2 * Co * C)) if (xhi is not None and xlo is not None and point_gemm_uses_tensor_cores(C))
            else f((0,)), f((AMAX_SLOTS,))]


@torch.library.custom_op("edgeconv_b200::edgeconv_bwd", mutates_args=(), device_types="cuda")
def edgeconv_bwd_op(gout: Optional[Tensor], gout_pm: Optional[Tensor], x: Tensor, idx: Tensor, sel: Tensor,
                    arg: Tensor, esum: Tensor, wsplit: Tensor,
                    Y: Tensor, Wcat: Tensor, affine: Tensor, stats: Tensor, use_batch_stats: bool,
                    slope: float, subtract_center: bool, group: int, xhi: Optional[Tensor],
                    xlo: Optional[Tensor], need_dx: bool) -> List[Tensor]:
    """-> [dx [B,C,N] (empty unless need_dx), dW [Co,2C], dgamma [Co], dbeta [Co]]"""
    B, C, N = x.shape
    k = idx.shape[-1]
    Co = sel.shape[1]
    M = B * N
    dev = x.device
    gout = None if gout is None else gout.contiguous().float()
    ld_pm = Co
    if gout_pm is not None:
        gout_pm = gout_pm.float()
        if gout_pm.dim() == 2 and gout_pm.stride(1) == 1 and gout_pm.stride(0) >= Co:
            ld_pm = gout_pm.stride(0)      # e.g. a column slice of the concat's gradient: read in place
        else:
            gout_pm = gout_pm.contiguous()
    with torch.cuda.device(dev):
        st = _stream(x)
        f32 = dict(device=dev, dtype=torch.float32)
        g = torch.empty(M, Co, **f32)
        bstats = torch.zeros(2 * Co, device=dev, dtype=torch.float64)
        mean, invstd, a, b = (c_void_p(affine.data_ptr() + 4 * Co * r) for r in range(4))
        _lib.call("ecb200_bwd_prep", _ptr(gout), _ptr(gout_pm), ld_pm, _ptr(sel), a, b, mean, invstd,
                  float(slope), B, N, Co, _ptr(g), _ptr(bstats), st)
        dgamma = torch.empty(Co, **f32)
        dbeta = torch.empty(Co, **f32)
        cc = torch.empty(2, Co, **f32)
        c1, c2 = c_void_p(cc.data_ptr()), c_void_p(cc.data_ptr() + 4 * Co)
        if use_batch_stats:
            _bstats_to_coeffs(bstats, group, c_void_p(stats.data_ptr() + 8 * 2 * Co), a, invstd, Co, dgamma,
                              dbeta, c1, c2, st)
        else:
            _lib.call("ecb200_bwd_finalize", _ptr(bstats), _ptr(bstats), None, a, invstd, 0, Co, _ptr(dgamma),
                      _ptr(dbeta), c1, c2, st)
        rowptr = src = None
        if use_batch_stats:
            rowptr = torch.empty(M + 1, device=dev, dtype=torch.int32)
            src = torch.empty(M * k, device=dev, dtype=torch.int32)
            cursor = torch.empty(M, device=dev, dtype=torch.int32)
            _lib.call("ecb200_reverse_graph", _ptr(idx), B, N, k, _ptr(rowptr), _ptr(src), _ptr(cursor), st)
        dx = torch.empty(B, C, N, **f32) if need_dx else torch.empty(0, **f32)
        dWcat = torch.empty(2 * Co, C, **f32)
        dW = torch.empty(Co, 2 * C, **f32)
        if xhi is not None and xlo is not None and bwd_gemm_uses_tensor_cores(C, Co):
            # tensor-core GEMMs (3xTF32): the two graph kernels write dY = [dU | dV] directly as
            # tf32 hi/lo halves -- sparse scatter first (into a zeroed [M,Co] accumulator), then
            # the dense BatchNorm terms are added and the sum is split
            dYs = torch.empty(2, M, 2 * Co, **f32)
            dU = torch.zeros(M, Co, **f32)
            if wsplit.numel() == 4 * 2 * Co * C:
                wT = wsplit.view(4, C, 2 * Co)[2:]           # Wcat^T hi/lo from the forward
            else:
                wT = torch.empty(2, C, 2 * Co, **f32)
                _lib.call("ecb200_transpose_split_tf32", _ptr(Wcat), 2 * Co, C, _ptr(wT[0]), _ptr(wT[1]), st)
            _lib.call("ecb200_bwd_scatter", _ptr(g), _ptr(esum), _ptr(arg), _ptr(idx), a, mean, c1, c2,
                      B, N, k, Co, None, _ptr(dU), _ptr(dYs[0]), _ptr(dYs[1]), st)
            _lib.call("ecb200_bwd_dense", _ptr(Y), _ptr(rowptr), _ptr(src), mean, c1, c2,
                      int(use_batch_stats), B, N, Co, None, _ptr(dU), _ptr(dYs[0]), _ptr(dYs[1]), st)
            if need_dx:
                _lib.call("ecb200_gemm_dx_tc", _ptr(dYs[0]), _ptr(dYs[1]), _ptr(wT[0]), _ptr(wT[1]), B, C, N,
                          2 * Co, _ptr(dx), st)
            _lib.call("ecb200_gemm_dw_tc", _ptr(dYs[0]), _ptr(dYs[1]), _ptr(xhi), _ptr(xlo), M, C, 2 * Co,
                      _ptr(dWcat), st)
        else:
            dY = torch.empty(M, 2 * Co, **f32)
            _lib.call("ecb200_bwd_dense", _ptr(Y), _ptr(rowptr), _ptr(src), mean, c1, c2,
                      int(use_batch_stats), B, N, Co, _ptr(dY), None, None, None, st)
            _lib.call("ecb200_bwd_scatter", _ptr(g), _ptr(esum), _ptr(arg), _ptr(idx), a, mean,
                      c1, c2, B, N, k, Co, _ptr(dY), None, None, None, st)
            if need_dx:
                _lib.call("ecb200_gemm_dx", _ptr(dY), _ptr(Wcat), B, C, N, 2 * Co, _ptr(dx), st)
            _lib.call("ecb200_gemm_dw", _ptr(dY), _ptr(x), B, C, N, 2 * Co, _ptr(dWcat), st)
        _lib.call("ecb200_unpack_weight_grad", _ptr(dWcat), Co, C, int(subtract_center), _ptr(dW), st)
    return [dx, dW, dgamma, dbeta]


@edgeconv_bwd_op.register_fake
def _(gout, gout_pm, x, idx, sel, arg, esum, wsplit, Y, Wcat, affine, stats, use_batch_stats, slope,
      subtract_center, group, xhi, xlo, need_dx):
    B, C, N = x.shape
    Co = sel.shape[1]
    f = x.new_empty
    return [f((B, C, N)) if need_dx else f((0,)), f((Co, 2 * C)), f((Co,)), f((Co,))]


def _ec_setup(ctx, inputs, output):
    (x, idx, weight, gamma, beta, _rm, _rv, use_batch_stats, _eps, slope, subtract_center, group,
     save_for_bwd, _xhi, _xlo, _emit_pm) = inputs
    out, out_pm, sel, arg, esum, Y, Wcat, affine, stats, wsplit = output[:10]
    if not save_for_bwd:
        raise RuntimeError("edgeconv_b200: forward ran with save_for_bwd=False but a gradient "
                           "is required")
    ctx.save_for_backward(x, idx, sel, arg, esum, Y, Wcat, affine, stats, _xhi, _xlo, wsplit)
    ctx.cfg = (use_batch_stats, slope, subtract_center, group)
    ctx.need_dx = bool(x.requires_grad)
    ctx.wshape = tuple(weight.shape)
    ctx.set_materialize_grads(False)


def _ec_backward(ctx, grads):
    gout, gout_pm = grads[0], grads[1]
    n_in = 16
    if gout is None and gout_pm is None:
        return (None,) * n_in
    x, idx, sel, arg, esum, Y, Wcat, affine, stats, xhi, xlo, wsplit = ctx.saved_tensors
    use_batch_stats, slope, subtract_center, group = ctx.cfg
    dx, dW, dgamma, dbeta = edgeconv_bwd_op(gout, gout_pm, x, idx, sel, arg, esum, wsplit, Y, Wcat, affine, stats,
                                            use_batch_stats, slope, subtract_center, group, xhi, xlo,
                                            ctx.need_dx)
    return (dx if ctx.need_dx else None, None, dW.view(ctx.wshape), dgamma, dbeta) + (None,) * (n_in - 5)


edgeconv_fwd_op.register_autograd(_ec_backward, setup_context=_ec_setup)


@torch.library.custom_op("edgeconv_b200::bn_update_running",
                         mutates_args=("running_mean", "running_var", "num_batches_tracked"),
                         device_types="cuda")
def bn_update_running_op(stats: Tensor, running_mean: Optional[Tensor], running_var: Optional[Tensor],
                         num_batches_tracked: Optional[Tensor], momentum: float) -> None:
    """The in-place side effect of a training-mode BatchNorm2d forward (running statistics,
    num_batches_tracked), from the global [sum e, sum e^2, count] buffer."""
    Co = (stats.numel() - 1) // 2
    with torch.cuda.device(stats.device):
        _lib.call("ecb200_bn_update_running", _ptr(stats), Co, float(momentum), _ptr(running_mean),
                  _ptr(running_var), _ptr(num_batches_tracked), _stream(stats))


def point_gemm_uses_tensor_cores(C: int) -> bool:
    """ECB200_GEMM=fma forces the FP32-FMA GEMM (A/B tests)."""
    return os.environ.get("ECB200_GEMM", "auto") != "fma" and C % 32 == 0 and 32 <= C <= 128


def bwd_gemm_uses_tensor_cores(C: int, Co: int) -> bool:
    return (os.environ.get("ECB200_GEMM", "auto") != "fma" and C in (32, 64, 128)
            and (2 * Co) % 32 == 0)


def edgeconv(x: Tensor, idx: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor,
             running_mean: Optional[Tensor], running_var: Optional[Tensor],
             num_batches_tracked: Optional[Tensor], training: bool, momentum: Optional[float] = 0.1,
             eps: float = 1e-5, slope: float = 0.2, subtract_center: bool = False,
             group: int = 0, xhi: Optional[Tensor] = None, xlo: Optional[Tensor] = None,
             return_point_major: bool = False):
    """Fused EdgeConv block on a given kNN graph:
    max_k LeakyReLU(BatchNorm2d(Conv2d_1x1([x_j (- x_i) ; x_i])))  ->  [B, Co, N]
    (models/dgcnn.py:84-86 with :54-58).  BatchNorm semantics follow nn.BatchNorm2d:
    batch statistics when ``training`` or when no running statistics exist."""
    use_batch_stats = bool(training or running_mean is None or running_var is None)
    update_running = bool(training and running_mean is not None)
    need_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (x, weight, gamma, beta))
    mom = -1.0 if momentum is None else float(momentum)
    res = edgeconv_fwd_op(x, idx, weight, gamma, beta, running_mean, running_var, use_batch_stats,
                          float(eps), float(slope), bool(subtract_center), int(group), bool(need_grad),
                          xhi, xlo, bool(return_point_major))
    if update_running:
        bn_update_running_op(res[8].detach(), running_mean, running_var, num_batches_tracked, mom)
    # side channel to the next layer's ecb200_split_f16: valid while `out` is not written in place
    res[0]._ecb200_amax = (res[10].detach(), res[0]._version)
    if return_point_major:
        return res[0], res[1]          # [B,Co,N] and the same values as [B*N, Co]
    return res[0]


def known_amax(x: Tensor) -> Optional[Tensor]:
    """max |x| [1] if x is the untouched output of edgeconv() (which measured it while writing x)."""
    tag = getattr(x, "_ecb200_amax", None)
    if tag is not None and tag[1] == x._version and tag[0].device == x.device:
        return tag[0]
    return None


# ------------------------------------------- conv5's BN + LeakyReLU + global max|avg pooling
@torch.library.custom_op("edgeconv_b200::embed_pool_fwd", mutates_args=(), device_types="cuda")
def embed_pool_fwd_op(z: Tensor, B: int, N: int, gamma: Tensor, beta: Tensor,
                      running_mean: Optional[Tensor], running_var: Optional[Tensor],
                      use_batch_stats: bool, eps: float, slope: float, group: int,
                      stats_in: Optional[Tensor] = None) -> List[Tensor]:
    """z [B*N, E] (conv5's raw output, point-major) -> [pooled [B,2E], arg [B,E] i32,
    affine [4,E] = (mean, invstd, a, b), stats [2E+1] f64]:  pooled = (max_n | mean_n) of
    LeakyReLU(BatchNorm(z)) -- dgcnn.py:75-78,:102 followed by upstream DGCNN_cls's pooling."""
    _check_cuda_f32("z", z, 2)
    M, E = z.shape
    if M != B * N:
        raise RuntimeError(f"edgeconv_b200: z has {M} rows, expected B*N = {B * N}")
    if E % 4 != 0:
        raise RuntimeError(f"edgeconv_b200: embedding width must be a multiple of 4, got {E}")
    z = z.contiguous()
    gamma_c, beta_c = gamma.detach().contiguous().float(), beta.detach().contiguous().float()
    dev = z.device
    with torch.cuda.device(dev):
        st = _stream(z)
        f32 = dict(device=dev, dtype=torch.float32)
        # [sum z | sum z^2 | count]: from the GEMM's epilogue (ecb200_embed_gemm) when it produced z
        stats = (stats_in.clone() if stats_in is not None
                 else torch.zeros(2 * E + 1, device=dev, dtype=torch.float64))
        affine = torch.empty(4, E, **f32)
        mean, invstd, a, b = (c_void_p(affine.data_ptr() + 4 * E * r) for r in range(4))
        pooled = torch.empty(B, 2 * E, **f32)
        arg = torch.empty(B, E, device=dev, dtype=torch.int32)
        if use_batch_stats:
            if stats_in is None:
                _lib.call("ecb200_colstats", _ptr(z), M, E, _ptr(stats), st)
            _stats_to_affine(stats, group, gamma_c, beta_c, eps, E, mean, invstd, a, b, st)
        else:
            _lib.call("ecb200_bn_finalize", None, _ptr(gamma_c), _ptr(beta_c), _ptr(running_mean),
                      _ptr(running_var), 0, float(eps), E, mean, invstd, a, b, st)
        _lib.call("ecb200_embed_pool", _ptr(z), a, b, float(slope), B, N, E, _ptr(pooled), _ptr(arg), st)
    return [pooled, arg, affine, stats]


@embed_pool_fwd_op.register_fake
def _(z, B, N, gamma, beta, running_mean, running_var, use_batch_stats, eps, slope, group, stats_in=None):
    E = z.shape[1]
    f = z.new_empty
    return [f((B, 2 * E)), f((B, E), dtype=torch.int32), f((4, E)), f((2 * E + 1,), dtype=torch.float64)]


@torch.library.custom_op("edgeconv_b200::embed_pool_bwd", mutates_args=(), device_types="cuda")
def embed_pool_bwd_op(gpool: Tensor, z: Tensor, arg: Tensor, affine: Tensor, stats: Tensor, B: int,
                      N: int, use_batch_stats: bool, slope: float, group: int) -> List[Tensor]:
    """-> [dz [B*N,E], dgamma [E], dbeta [E]]"""
    M, E = z.shape
    dev = z.device
    gpool = gpool.contiguous().float()
    with torch.cuda.device(dev):
        st = _stream(z)
        f32 = dict(device=dev, dtype=torch.float32)
        mean, invstd, a, b = (c_void_p(affine.data_ptr() + 4 * E * r) for r in range(4))
        bstats = torch.zeros(2 * E, device=dev, dtype=torch.float64)
        _lib.call("ecb200_embed_pool_bwd_stats", _ptr(z), _ptr(gpool), _ptr(arg), a, b, mean, invstd,
                  float(slope), B, N, E, _ptr(bstats), st)
        dgamma = torch.empty(E, **f32)
        dbeta = torch.empty(E, **f32)
        cc = torch.empty(2, E, **f32)
        c1, c2 = c_void_p(cc.data_ptr()), c_void_p(cc.data_ptr() + 4 * E)
        if use_batch_stats:
            _bstats_to_coeffs(bstats, group, c_void_p(stats.data_ptr() + 8 * 2 * E), a, invstd, E, dgamma, dbeta,
                              c1, c2, st)
        else:
            _lib.call("ecb200_bwd_finalize", _ptr(bstats), _ptr(bstats), None, a, invstd, 0, E, _ptr(dgamma),
                      _ptr(dbeta), c1, c2, st)
        dz = torch.empty(M, E, **f32)
        _lib.call("ecb200_embed_pool_bwd_dz", _ptr(z), _ptr(gpool), _ptr(arg), a, b, mean, c1, c2,
                  float(slope), B, N, E, _ptr(dz), st)
    return [dz, dgamma, dbeta]


@embed_pool_bwd_op.register_fake
def _(gpool, z, arg, affine, stats, B, N, use_batch_stats, slope, group):
    E = z.shape[1]
    return [z.new_empty(z.shape), z.new_empty((E,)), z.new_empty((E,))]


def _ep_setup(ctx, inputs, output):
    z, B, N, _g, _b, _rm, _rv, use_batch_stats, _eps, slope, group = inputs[:11]
    _pooled, arg, affine, stats = output
    ctx.save_for_backward(z, arg, affine, stats)
    ctx.cfg = (B, N, use_batch_stats, slope, group)
    ctx.set_materialize_grads(False)


def _ep_backward(ctx, grads):
    gpool = grads[0]
    n_in = len(ctx.needs_input_grad)      # 11 or 12: a defaulted trailing stats_in may not be part of the call
    if gpool is None:
        return (None,) * n_in
    z, arg, affine, stats = ctx.saved_tensors
    B, N, use_batch_stats, slope, group = ctx.cfg
    dz, dgamma, dbeta = embed_pool_bwd_op(gpool, z, arg, affine, stats, B, N, use_batch_stats, slope, group)
    return (dz, None, None, dgamma, dbeta) + (None,) * (n_in - 5)


embed_pool_fwd_op.register_autograd(_ep_backward, setup_context=_ep_setup)


def embed_pool(z: Tensor, B: int, N: int, gamma: Tensor, beta: Tensor, running_mean: Optional[Tensor],
               running_var: Optional[Tensor], num_batches_tracked: Optional[Tensor], training: bool,
               momentum: Optional[float] = 0.1, eps: float = 1e-5, slope: float = 0.2,
               group: int = 0, stats: Optional[Tensor] = None) -> Tensor:
    """cat(max_n, mean_n) of LeakyReLU(BatchNorm(z)) over the N points of each cloud -> [B, 2E].
    z [B*N, E] point-major; BatchNorm semantics as nn.BatchNorm2d / SyncBatchNorm (see edgeconv).
    ``stats``: this rank's [sum z | sum z^2 | count] if the producer of z already has them."""
    use_batch_stats = bool(training or running_mean is None or running_var is None)
    update_running = bool(training and running_mean is not None)
    mom = -1.0 if momentum is None else float(momentum)
    res = embed_pool_fwd_op(z, int(B), int(N), gamma, beta, running_mean, running_var, use_batch_stats,
                            float(eps), float(slope), int(group), stats.detach() if stats is not None else None)
    if update_running:
        bn_update_running_op(res[3].detach(), running_mean, running_var, num_batches_tracked, mom)
    return res[0]


# ------------------------------------------------ conv5 as a per-point GEMM with BN statistics
def embed_gemm_mode(K: int, E: int) -> str:
    """How conv5's GEMM runs: "cudnn" (library convolution), "tf32" (own tcgen05 kernel, plain TF32
    on the raw fp32 operands, statistics in the epilogue) or "3xtf32" (own kernel, fp32-equivalent).
    auto (default): the library convolution while torch.backends.cudnn.allow_tf32 is on (PyTorch's
    default: its 2-SM 256x256 TF32 kernel runs at the L2-bandwidth bound of this fp32-operand GEMM,
    67 us at M=32768, vs 240 us for the own 128x128-tile kernel -- profiles/r2_embed_gemm.txt), the
    own 3xTF32 kernel when it is off (fp32 semantics asked for: 359 us incl. the operand split, where
    the library falls back to an FP32 SIMT convolution).  ECB200_CONV5=cudnn|tf32|3xtf32 overrides."""
    mode = os.environ.get("ECB200_CONV5", "auto")
    if mode == "cudnn" or K % 32 != 0 or E % 128 != 0:
        return "cudnn"
    if mode in ("tf32", "3xtf32"):
        return mode
    return "cudnn" if torch.backends.cudnn.allow_tf32 else "3xtf32"


@torch.library.custom_op("edgeconv_b200::embed_gemm", mutates_args=(), device_types="cuda")
def embed_gemm_op(x_pm: Tensor, weight: Tensor, three: bool) -> List[Tensor]:
    """x_pm [M,K] (the channels-last concat of dgcnn.py:100), weight [E,K,1,1] (conv5) ->
    [z [M,E] = x_pm . W^T, stats [2E+1] f64 = [sum z | sum z^2 | M]]  (ecb200_embed_gemm)."""
    _check_cuda_f32("x_pm", x_pm, 2)
    M, K = x_pm.shape
    E = weight.shape[0]
    x_pm = x_pm.contiguous()
    w = weight.detach().reshape(E, K).contiguous().float()
    dev = x_pm.device
    with torch.cuda.device(dev):
        st = _stream(x_pm)
        z = torch.empty(M, E, device=dev, dtype=torch.float32)
        stats = torch.zeros(2 * E + 1, device=dev, dtype=torch.float64)
        if three:
            xs = torch.empty(2, M, K, device=dev, dtype=torch.float32)
            ws = torch.empty(2, E, K, device=dev, dtype=torch.float32)
            _lib.call("ecb200_split_rows_tf32", _ptr(x_pm), M * K, _ptr(xs[0]), _ptr(xs[1]), st)
            _lib.call("ecb200_split_rows_tf32", _ptr(w), E * K, _ptr(ws[0]), _ptr(ws[1]), st)
            _lib.call("ecb200_embed_gemm", _ptr(xs[0]), _ptr(xs[1]), _ptr(ws[0]), _ptr(ws[1]), M, K, E, _ptr(z),
                      _ptr(stats), st)
        else:
            _lib.call("ecb200_embed_gemm", _ptr(x_pm), None, _ptr(w), None, M, K, E, _ptr(z), _ptr(stats), st)
    return [z, stats]


@embed_gemm_op.register_fake
def _(x_pm, weight, three):
    M, E = x_pm.shape[0], weight.shape[0]
    return [x_pm.new_empty((M, E)), x_pm.new_empty((2 * E + 1,), dtype=torch.float64)]


def _eg_setup(ctx, inputs, output):
    x_pm, weight, _three = inputs
    ctx.save_for_backward(x_pm, weight)
    ctx.set_materialize_grads(False)


def _eg_backward(ctx, grads):
    gz = grads[0]
    if gz is None:
        return None, None, None
    x_pm, weight = ctx.saved_tensors
    M, K = x_pm.shape
    E = weight.shape[0]
    # the two backward GEMMs stay plain library calls -- the convolution backward on channels-last
    # views, exactly what autograd ran before this op existed (same TF32 policy as the reference)
    x4 = x_pm.view(1, M, 1, K).permute(0, 3, 1, 2)
    g4 = gz.contiguous().view(1, M, 1, E).permute(0, 3, 1, 2)
    dx4, dw4, _ = torch.ops.aten.convolution_backward(
        g4, x4, weight.reshape(E, K, 1, 1), None, [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
        [bool(ctx.needs_input_grad[0]), bool(ctx.needs_input_grad[1]), False])
    dx = dx4.permute(0, 2, 3, 1).reshape(M, K) if dx4 is not None else None
    dw = dw4.reshape(weight.shape) if dw4 is not None else None
    return dx, dw, None


embed_gemm_op.register_autograd(_eg_backward, setup_context=_eg_setup)


# ---------------------------------------------------- two-conv edge block, fused forward (row f-1)
def two_conv_block_supported(C: int, block1, block2, k: int) -> bool:
    """Shapes ecb200_two_conv_fwd takes: bias-free 1x1 convs, C1 in {32,64}, C2 in {32,64,128}."""
    import torch.nn as nn
    try:
        c1, bn1, c2, bn2 = block1[0], block1[1], block2[0], block2[1]
    except (IndexError, TypeError):
        return False
    ok = (isinstance(c1, nn.Conv2d) and isinstance(c2, nn.Conv2d) and c1.bias is None and c2.bias is None
          and tuple(c1.kernel_size) == (1, 1) and tuple(c2.kernel_size) == (1, 1)
          and isinstance(bn1, nn.modules.batchnorm._BatchNorm) and isinstance(bn2, nn.modules.batchnorm._BatchNorm)
          and bn1.affine and bn2.affine)
    return bool(ok and c1.in_channels == 2 * C and c1.out_channels in (32, 64) and c2.in_channels == c1.out_channels
                and c2.out_channels in (32, 64, 128) and 1 <= k <= MAX_K
                and os.environ.get("ECB200_TWO_CONV", "fused") == "fused")


def _bn_affine(bn, stats: Optional[Tensor], group: int, Co: int, dev) -> Tensor:
    """[4, Co] = (mean, invstd, a, b) of a BatchNorm layer: from the batch statistics buffer (training
    mode or no running statistics; all-reduced under SyncBatchNorm, running statistics updated) or
    from the running statistics."""
    use_batch = stats is not None
    affine = torch.empty(4, Co, device=dev, dtype=torch.float32)
    mean, invstd, a, b = (c_void_p(affine.data_ptr() + 4 * Co * r) for r in range(4))
    g32, b32 = bn.weight.detach().contiguous().float(), bn.bias.detach().contiguous().float()
    st = c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    if use_batch and group:
        _allreduce_stats(stats, group)
    _lib.call("ecb200_bn_finalize", _ptr(stats), _ptr(g32), _ptr(b32),
              None if use_batch else _ptr(bn.running_mean), None if use_batch else _ptr(bn.running_var),
              int(use_batch), float(bn.eps), Co, mean, invstd, a, b, st)
    if use_batch and bn.training and bn.running_mean is not None:
        mom = -1.0 if bn.momentum is None else float(bn.momentum)
        _lib.call("ecb200_bn_update_running", _ptr(stats), Co, mom, _ptr(bn.running_mean), _ptr(bn.running_var),
                  _ptr(bn.num_batches_tracked), st)
    return affine


@torch.no_grad()
def two_conv_block(x: Tensor, idx: Tensor, block1, block2, subtract_center: bool = False,
                   group1: int = 0, group2: int = 0) -> Tensor:
    """max_k LeakyReLU(BN2(W2 . LeakyReLU(BN1(W1 . [x_j (- x_i) ; x_i]))))  ->  [B, C2, N], fused
    (csrc/two_conv.cu); forward only -- see dgcnn.two_conv_edge_block for the differentiable path."""
    _check_cuda_f32("x", x, 3)
    B, C, N = x.shape
    k = idx.shape[-1]
    conv1, bn1, act1 = block1[0], block1[1], block1[2]
    conv2, bn2, act2 = block2[0], block2[1], block2[2]
    C1, C2 = conv1.out_channels, conv2.out_channels
    M = B * N
    dev = x.device
    x = x.contiguous()
    idx = idx.contiguous()
    f32 = dict(device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        st = _stream(x)
        w1 = conv1.weight.detach().reshape(C1, 2 * C).contiguous().float()
        Wcat = torch.empty(2 * C1, C, **f32)
        Y = torch.empty(M, 2 * C1, **f32)
        _lib.call("ecb200_pack_weight", _ptr(w1), C1, C, int(subtract_center), _ptr(Wcat), st)
        _lib.call("ecb200_point_gemm", _ptr(x), _ptr(Wcat), B, C, N, 2 * C1, _ptr(Y), st)
        batch1 = bool(bn1.training or bn1.running_mean is None)
        stats1 = None
        if batch1:
            # BatchNorm1 statistics over all edges: the gather kernel of the single-conv path
            stats1 = torch.zeros(2 * C1 + 1, device=dev, dtype=torch.float64)
            g1 = bn1.weight.detach().contiguous().float()
            sel1 = torch.empty(M, C1, **f32)
            arg1 = torch.empty(M, C1, device=dev, dtype=torch.uint8)
            _lib.call("ecb200_edge_gather", _ptr(Y), _ptr(idx), _ptr(g1), B, N, k, C1, _ptr(sel1), _ptr(arg1),
                      None, _ptr(stats1), st)
        aff1 = _bn_affine(bn1, stats1, group1, C1, dev)
        w2 = conv2.weight.detach().reshape(C2, C1).contiguous().float()
        w2s = torch.empty(2, C2, C1, **f32)
        _lib.call("ecb200_split_rows_tf32", _ptr(w2), C2 * C1, _ptr(w2s[0]), _ptr(w2s[1]), st)
        batch2 = bool(bn2.training or bn2.running_mean is None)
        stats2 = torch.zeros(2 * C2 + 1, device=dev, dtype=torch.float64) if batch2 else None
        sel2 = torch.empty(M, C2, **f32)
        g2 = bn2.weight.detach().contiguous().float()
        _lib.call("ecb200_two_conv_fwd", _ptr(Y), _ptr(idx), _ptr(aff1[2]), _ptr(aff1[3]),
                  float(getattr(act1, "negative_slope", 0.0)), _ptr(w2s[0]), _ptr(w2s[1]), _ptr(g2),
                  B, N, k, C1, C2, _ptr(sel2), _ptr(stats2), st)
        aff2 = _bn_affine(bn2, stats2, group2, C2, dev)
        out = torch.empty(B, C2, N, **f32)
        _lib.call("ecb200_edge_apply", _ptr(sel2), _ptr(aff2[2]), _ptr(aff2[3]),
                  float(getattr(act2, "negative_slope", 0.0)), B, N, C2, _ptr(out), None, C2, st)
    return out


# ------------------------------------------------------------- kNN result reuse (row f-3, part 1)
# The reference's part-seg Net computes knn(src, k) on the SAME xyz tensor three times per forward
# (models/dgcnn.py:84 via DGCNN.forward, models/model_partseg.py:26 in compute_hog_1x1, and
# models/layers.py:45 via PositionEmbedding).  One entry per device: the last xyz-sized query and
# its graph.  The key holds the tensor's storage (so its address cannot be recycled while cached)
# and its version counter (bumped by any in-place write); never used while a CUDA graph is captured.
_KNN_CACHE = {}
knn_cache_hits = 0


def knn_cached(x: Tensor, k: int, sorted: bool = True) -> Tensor:
    """knn_op with reuse of the previous result for an identical (storage, view, version, k) query.
    Only small-C inputs (the xyz layer) are cached; ECB200_KNN_CACHE=0 disables it."""
    global knn_cache_hits
    if (x.shape[1] > 16 or os.environ.get("ECB200_KNN_CACHE", "1") == "0" or not x.is_cuda
            or torch.cuda.is_current_stream_capturing()):
        return knn_op(x, k, sorted)
    st = x.untyped_storage()
    key = (st.data_ptr(), x.storage_offset(), tuple(x.shape), tuple(x.stride()), x._version, int(k), x.dtype)
    ent = _KNN_CACHE.get(x.device.index)
    if ent is not None and ent[0] == key and (ent[3] or not sorted):
        knn_cache_hits += 1
        return ent[2]
    idx = knn_op(x, k, sorted)
    _KNN_CACHE[x.device.index] = (key, st, idx, bool(sorted))
    return idx


# ------------------------------------------------------------------------- autocast
# Under torch.autocast (main_partseg_dist.py:253) the reference's bmm / conv2d would run in fp16
# while its pow/sum and BatchNorm statistics stay fp32.  The fused path always
# computes in fp32 (3xTF32 on the tensor cores): every op casts its floating-point inputs to fp32
# and runs with autocast disabled.  Gradients are linear in the incoming gradient, so GradScaler's
# loss scaling and its inf/nan detection pass through unchanged.
# (knn_tc_f16_op is left out on purpose: its operands ARE fp16 bit patterns made by split_f16_op)
for _op in (knn_op, split_tf32_op, knn_tc_op, split_f16_op, knn_tc_xyz_op, graph_feature_op, graph_feature_bwd_op, edgeconv_fwd_op,
            edgeconv_bwd_op, embed_pool_fwd_op, embed_pool_bwd_op, embed_gemm_op):
    _op.register_autocast("cuda", torch.float32)


def check_neighbour_indices(idx: Tensor, B: int, N: int) -> Tensor:
    """Validate caller-supplied neighbour indices (shape [B,N,k], 0 <= idx < N) and return them as
    contiguous int32.  The kernels index device memory with them unchecked, where the reference's
    advanced indexing (dgcnn.py:33) would raise; the range check is asynchronous (device-side
    assert) and skipped while a CUDA graph is being captured."""
    if idx.dim() != 3 or idx.shape[0] != B or idx.shape[1] != N:
        raise ValueError(f"edgeconv_b200: idx must have shape [{B}, {N}, k], got {tuple(idx.shape)}")
    if idx.dtype not in (torch.int32, torch.int64):
        raise TypeError(f"edgeconv_b200: idx must be int32 or int64, got {idx.dtype}")
    if idx.is_cuda and not torch.cuda.is_current_stream_capturing():
        lo, hi = torch.aminmax(idx)
        torch._assert_async((lo >= 0) & (hi < N), "edgeconv_b200: neighbour index out of range [0, N)")
    return idx.to(torch.int32).contiguous()
