"""Drop-in for the reference's ``models/dgcnn.py`` (knn, get_graph_feature, DGCNN).

Same names, call signatures, return layouts, parameter names and state_dict keys
as /root/reference/models/dgcnn.py, so ``models/model_partseg.py`` (Net),
``models/layers.py`` (PositionEmbedding) and the training scripts run unchanged
on top of it -- but every tensor operation of the EdgeConv path is a hand-written
sm_100a kernel reached through the C ABI (see ops.py / include/edgeconv_b200.h).
CUDA tensors only; CPU tensors raise (the CPU path is the reference itself).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.nn as nn

from . import ops


def knn(x: torch.Tensor, k: int) -> torch.Tensor:
    """models/dgcnn.py:6-12.  x [B,C,N] -> int64 [B,N,k]: indices of the k nearest
    points (self included) by the reference's -|xi|^2 + 2 xi.xj - |xj|^2 score,
    nearest first; equal scores resolve to the smaller index."""
    return ops.knn_cached(_as_f32(x), int(k)).long()


def _as_f32(x: torch.Tensor) -> torch.Tensor:
    # under autocast the reference's matmul would run in fp16; the fused
    # path always computes in fp32, so half inputs are widened here
    if x.dim() != 3:
        raise ValueError(f"expected x of shape [B, C, N], got {tuple(x.shape)}")
    if not x.is_cuda:
        raise RuntimeError("edgeconv_b200 runs on CUDA tensors only: there is no CPU fallback "
                           f"(got a tensor on {x.device}); the CPU path is the reference itself")
    return x if x.dtype == torch.float32 else x.float()


def get_graph_feature(x: torch.Tensor, k: int = 20, knn_only: bool = False,
                      disp_only: bool = False, idx: Optional[torch.Tensor] = None,
                      dim9: bool = False, subtract_center: bool = False) -> torch.Tensor:
    """models/dgcnn.py:15-44.  Default: [B,2C,N,k] = (x_j, x_i) along channels;
    ``knn_only`` -> [B,N,k,C] of x_j; ``disp_only`` -> [B,C,N,k] of x_j - x_i.
    Differentiable w.r.t. x (gather + centre copy), not w.r.t. the indices.

    Extras kept from upstream DGCNN, which main_semseg.py expects: ``idx`` (reuse a
    graph), ``dim9`` (build the graph on channels 6: of a 9-channel input) and
    ``subtract_center`` for the canonical (x_j - x_i, x_i) feature."""
    x = _as_f32(x)
    if idx is None:
        src = x[:, 6:] if dim9 else x
        idx32 = ops.knn_cached(src.contiguous(), int(k))
    else:
        idx32 = ops.check_neighbour_indices(idx, x.shape[0], x.shape[2])
    if knn_only:
        mode = ops.GF_KNN_ONLY
    elif disp_only:
        mode = ops.GF_DISP_ONLY
    else:
        mode = ops.GF_CONCAT_CENTERED if subtract_center else ops.GF_CONCAT
    return ops.graph_feature_op(x, idx32, mode)


def _sync_group(bn: nn.Module) -> int:
    """Handle of the process group whose ranks share this BatchNorm's statistics (SyncBatchNorm in
    training mode, main_partseg_dist.py:189); 0 = statistics stay local."""
    if isinstance(bn, nn.SyncBatchNorm) and bn.training and torch.distributed.is_available() \
            and torch.distributed.is_initialized():
        pg = bn.process_group if bn.process_group is not None else torch.distributed.group.WORLD
        if torch.distributed.get_world_size(pg) > 1:
            return ops.register_group(pg)
    return 0


def _edge_block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False),
                         nn.BatchNorm2d(cout),
                         nn.LeakyReLU(negative_slope=0.2, inplace=True))


def edgeconv_block(x: torch.Tensor, block: nn.Sequential, k: int,
                   idx: Optional[torch.Tensor] = None, subtract_center: bool = False,
                   return_point_major: bool = False):
    """One EdgeConv layer = ``block(get_graph_feature(x, k)).max(-1)[0]`` of the
    reference (models/dgcnn.py:84-86), fused.  ``block`` is the reference's own
    ``nn.Sequential(Conv2d(2C,Co,1,bias=False), BatchNorm2d | SyncBatchNorm,
    LeakyReLU)``; its parameters and buffers are read at call time, so
    ``SyncBatchNorm.convert_sync_batchnorm`` and DDP keep working.
    Returns (out [B,Co,N], idx int32 [B,N,k]); with ``return_point_major`` out is the pair
    (out [B,Co,N], out_pm [B*N,Co]) of the same values in both layouts."""
    conv, bn, act = block[0], block[1], block[2]
    if conv.bias is not None or tuple(conv.kernel_size) != (1, 1):
        raise RuntimeError("edgeconv_block expects a bias-free 1x1 Conv2d")
    x = _as_f32(x)
    B, C, N = x.shape
    if idx is not None:
        idx = ops.check_neighbour_indices(idx, B, N)
    xhi = xlo = None
    if ops.knn_uses_tensor_cores(C, N, int(k)) or (idx is not None and ops.point_gemm_uses_tensor_cores(C)):
        # feature-space layer: one split into tf32 hi/lo operands feeds both the tensor-core
        # kNN and the tensor-core per-point GEMM
        if idx is None and ops.knn_tc_kind(C, N, int(k)) == "f16":
            # one pass over x makes the packed fp16 halves of the kNN and the tf32 halves of the GEMMs
            gemm_tc = ops.point_gemm_uses_tensor_cores(C)
            hh, hl, nb, xxs, cmax, xhi, xlo, _ = ops.split_f16_op(x.detach().contiguous(), gemm_tc, ops.known_amax(x))
            idx = ops.knn_tc_f16_op(hh, hl, nb, xxs, cmax, B, N, int(k))
            if not gemm_tc:
                xhi = xlo = None
        else:
            xhi, xlo, xx = ops.split_tf32_op(x.detach().contiguous())
            if idx is None:
                idx = ops.knn_tc_op(xhi, xlo, xx, B, N, int(k))
    if idx is None:
        # the order over k is irrelevant here, but the xyz graph is asked for again by the callers of
        # knn() / get_graph_feature() on the same tensor (model_partseg.py:26, layers.py:45), which
        # need it nearest-first: compute it sorted once and let them reuse it
        idx = ops.knn_cached(x.detach().contiguous(), int(k), C <= 16)
    group = _sync_group(bn)
    slope = float(getattr(act, "negative_slope", 0.0))
    out = ops.edgeconv(x, idx, conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                       bn.num_batches_tracked, bn.training, bn.momentum, bn.eps, slope,
                       subtract_center, group, xhi, xlo, return_point_major)
    return out, idx


def two_conv_edge_block(x: torch.Tensor, block1: nn.Sequential, block2: nn.Sequential, k: int,
                        idx: Optional[torch.Tensor] = None, subtract_center: bool = False,
                        dim9: bool = False) -> torch.Tensor:
    """``block2(block1(get_graph_feature(x, k))).max(-1)[0]``: the two-conv edge block of the
    reference's PositionEmbedding (models/layers.py:45-52) and of upstream's part-seg / sem-seg
    EdgeConv blocks.  x [B,C,N] -> [B,Co2,N].

    In inference (no gradient required) the block runs FUSED: the first conv splits as
    W1.[x_j ; x_i] = U_j + V_i (one per-point GEMM), and one kernel gathers U rows, applies BN1 +
    LeakyReLU in registers, multiplies the 128-edge tile with W2 on the tensor cores (3xTF32,
    accumulators in tensor memory) and reduces max / min over k with the BN2 statistics in its
    epilogue (ops.two_conv_block) -- neither [B,2C,N,k] nor [B,Co1,N,k] ever exists.  When a
    gradient is required the block runs on the materialising path below (graph-feature kernel +
    the library convolutions under autograd), which is the reference's own arithmetic."""
    x = _as_f32(x)
    need_grad = torch.is_grad_enabled() and (x.requires_grad or any(
        p.requires_grad for blk in (block1, block2) for p in blk.parameters()))
    if not need_grad and ops.two_conv_block_supported(x.shape[1], block1, block2, int(k)):
        src = x[:, 6:].contiguous() if dim9 else x
        if idx is None:
            idx32 = ops.knn_op(src, int(k), False)
        else:
            idx32 = ops.check_neighbour_indices(idx, x.shape[0], x.shape[2])
        return ops.two_conv_block(x, idx32, block1, block2, subtract_center,
                                  _sync_group(block1[1]), _sync_group(block2[1]))
    gf = get_graph_feature(x, k=k, idx=idx, dim9=dim9, subtract_center=subtract_center)
    return block2(block1(gf)).max(dim=-1, keepdim=False)[0]


class DGCNN(nn.Module):
    """models/dgcnn.py:47-103: four EdgeConv layers on dynamic (feature-space) kNN
    graphs, concatenated and embedded by conv5.  ``args.k`` neighbours;
    ``args.emb_dim`` (or upstream's ``args.emb_dims``) embedding width.
    forward: x [B,3,N] -> [B, emb_dim, N]."""

    def __init__(self, args):
        super().__init__()
        emb = getattr(args, "emb_dim", None)
        if emb is None:
            emb = getattr(args, "emb_dims")
        self.emb_dims = emb
        self.k = args.k
        self.subtract_center = bool(getattr(args, "subtract_center", False))
        self.conv1 = _edge_block(3 * 2, 64)
        self.conv2 = _edge_block(64 * 2, 64)
        self.conv3 = _edge_block(64 * 2, 128)
        self.conv4 = _edge_block(128 * 2, 256)
        self.conv5 = _edge_block(512, self.emb_dims)
        self.record_idx = False
        self.last_idx: List[torch.Tensor] = []
        # the reference returns a contiguous [B, emb, N] tensor; internally the embedding is
        # produced point-major (channels-last).  Consumers that only reduce over the points
        # (DGCNN_cls) may set this to skip the final transposing copy.
        self.strided_output = False

    def _edge_features(self, x: torch.Tensor, idx_list=None) -> torch.Tensor:
        """The four EdgeConv layers (dgcnn.py:84-98) -> their concatenated outputs (dgcnn.py:100) as a
        channels-last [B,512,N,1] tensor ([B*N, 512] in memory)."""
        batch_size, _, num_points = x.size()
        feats = []
        if self.record_idx:
            self.last_idx = []
        h = x
        for layer, block in enumerate((self.conv1, self.conv2, self.conv3, self.conv4)):
            forced = None if idx_list is None else idx_list[layer].to(torch.int32)
            (h, h_pm), idx = edgeconv_block(h, block, self.k, idx=forced,
                                            subtract_center=self.subtract_center,
                                            return_point_major=True)
            if self.record_idx:
                self.last_idx.append(idx)
            feats.append(h_pm)
        # dgcnn.py:100: cat(x1..x4, dim=1) -> [B,512,N,1].  The fused layers also emit their output
        # as rows of channels, so the concat is built channels-last ([B*N, 512] in memory): conv5
        # (cuDNN) and its BatchNorm then run without NCHW<->NHWC transposes of 64-128 MiB tensors.
        return torch.cat(feats, dim=1).view(batch_size, num_points, 1, -1).permute(0, 3, 1, 2)   # NHWC strides

    def forward(self, x: torch.Tensor, idx_list=None) -> torch.Tensor:
        batch_size, _, num_points = x.size()
        out = self.conv5(self._edge_features(x, idx_list)).view(batch_size, -1, num_points)   # dgcnn.py:102
        return out if self.strided_output else out.contiguous()

    def forward_pooled(self, x: torch.Tensor, idx_list=None) -> torch.Tensor:
        """cat(max over the points, mean over the points) of ``forward(x)`` -> [B, 2*emb]: what
        upstream DGCNN_cls computes from the embedding.  conv5's convolution runs in cuDNN; its
        BatchNorm + LeakyReLU are fused with the pooling, so the [B, emb, N] tensor never exists."""
        batch_size, _, num_points = x.size()
        conv, bn, act = self.conv5[0], self.conv5[1], self.conv5[2]
        feats = self._edge_features(x, idx_list)                          # [B,512,N,1], channels-last
        mode = ops.embed_gemm_mode(conv.in_channels, conv.out_channels)
        stats = None
        if mode != "cudnn" and conv.bias is None and feats.dtype == torch.float32:
            # conv5 as a per-point GEMM on the tensor cores, BatchNorm statistics from its epilogue
            x_pm = feats.permute(0, 2, 3, 1).reshape(batch_size * num_points, -1)   # [B*N, 512] (a view)
            z, stats = ops.embed_gemm_op(x_pm, conv.weight, mode == "3xtf32")
            if not (bn.training or bn.running_mean is None):
                stats = None
        else:
            z = conv(feats)                                                # library convolution
            z = z.permute(0, 2, 3, 1).reshape(batch_size * num_points, -1)    # [B*N, emb] (a view)
        return ops.embed_pool(z, batch_size, num_points, bn.weight, bn.bias, bn.running_mean,
                              bn.running_var, bn.num_batches_tracked, bn.training, bn.momentum, bn.eps,
                              float(getattr(act, "negative_slope", 0.0)), _sync_group(bn), stats)
