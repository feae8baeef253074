"""Classification model used by the headline benchmark ("DGCNN-cls").

The reference fork ships only the backbone (models/dgcnn.py:47-103); its
main_cls.py:25,56 imports a ``DGCNN_cls`` from a ``model.py`` that does not exist
there (SURVEY.md §0 trap 2).  Following BASELINE.md §3.4, DGCNN-cls here is the
reference backbone plus upstream DGCNN's classification head -- global max + average
pooling over the points, then 2*emb -> 512 -> 256 -> classes -- so that the CPU
reference arm and the B200 arm time the same network.  The head is plain torch
(library GEMMs on tiny [B, *] matrices); it is not part of the EdgeConv path.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .dgcnn import DGCNN, _edge_block, edgeconv_block, two_conv_edge_block
from .syncbn import batch_norm_rows


class ClsHead(nn.Module):
    """[B, emb, N] -> logits [B, classes] (upstream DGCNN_cls tail)."""

    def __init__(self, emb_dims: int, output_channels: int = 40, dropout: float = 0.5):
        super().__init__()
        self.linear1 = nn.Linear(emb_dims * 2, 512, bias=False)
        self.bn6 = nn.BatchNorm1d(512)
        self.dp1 = nn.Dropout(p=dropout)
        self.linear2 = nn.Linear(512, 256)
        self.bn7 = nn.BatchNorm1d(256)
        self.dp2 = nn.Dropout(p=dropout)
        self.linear3 = nn.Linear(256, output_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() == 3:
            b = x.size(0)
            # == adaptive_max_pool1d / adaptive_avg_pool1d over the points (upstream head), as
            # plain reductions (torch's adaptive max pool kernel is ~20x slower at N=1024)
            x = torch.cat((x.max(dim=-1)[0].view(b, -1), x.mean(dim=-1).view(b, -1)), 1)
        # else: already pooled, [B, 2*emb] (DGCNN.forward_pooled)
        # bn6 / bn7: plain BatchNorm1d, or -- after SyncBatchNorm conversion -- the statistics exchange
        # of the EdgeConv layers instead of torch's NCCL collectives (syncbn.py)
        x = self.dp1(F.leaky_relu(batch_norm_rows(self.linear1(x), self.bn6), negative_slope=0.2))
        x = self.dp2(F.leaky_relu(batch_norm_rows(self.linear2(x), self.bn7), negative_slope=0.2))
        return self.linear3(x)


class DGCNN_cls(nn.Module):
    """``DGCNN_cls(args, output_channels=40)``: the name main_cls.py:56 asks for.
    args: k, emb_dims (or emb_dim), dropout."""

    def __init__(self, args, output_channels: int = 40):
        super().__init__()
        self.backbone = DGCNN(args)
        self.backbone.strided_output = True      # the head only reduces over the points
        self.head = ClsHead(self.backbone.emb_dims, output_channels,
                            float(getattr(args, "dropout", 0.5)))

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.backbone.forward_pooled(x))


def cal_loss(pred: torch.Tensor, gold: torch.Tensor, smoothing: bool = True) -> torch.Tensor:
    """Label-smoothed cross entropy, eps = 0.2 (the reference's loss.py:4-21, under the
    name main_cls.py:28 imports from util)."""
    gold = gold.contiguous().view(-1)
    if not smoothing:
        return F.cross_entropy(pred, gold, reduction="mean")
    eps = 0.2
    n_class = pred.size(1)
    target = torch.full_like(pred, eps / (n_class - 1)).scatter_(1, gold.view(-1, 1), 1.0 - eps)
    return -(target * F.log_softmax(pred, dim=1)).sum(dim=1).mean()


class PointNet(nn.Module):
    """``PointNet(args, output_channels=40)``: upstream's baseline classifier, the other name
    main_cls.py:25,54 imports from ``model``.  Plain per-point 1x1 convolutions + global max pool:
    no neighbourhood graph, so nothing of the EdgeConv path is involved (library ops throughout)."""

    def __init__(self, args, output_channels: int = 40):
        super().__init__()
        emb = getattr(args, "emb_dims", None) or getattr(args, "emb_dim")
        self.conv1 = nn.Conv1d(3, 64, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(64, 64, kernel_size=1, bias=False)
        self.conv3 = nn.Conv1d(64, 64, kernel_size=1, bias=False)
        self.conv4 = nn.Conv1d(64, 128, kernel_size=1, bias=False)
        self.conv5 = nn.Conv1d(128, emb, kernel_size=1, bias=False)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(64), nn.BatchNorm1d(64)
        self.bn4, self.bn5 = nn.BatchNorm1d(128), nn.BatchNorm1d(emb)
        self.linear1 = nn.Linear(emb, 512, bias=False)
        self.bn6 = nn.BatchNorm1d(512)
        self.dp1 = nn.Dropout()
        self.linear2 = nn.Linear(512, output_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        for conv, bn in ((self.conv1, self.bn1), (self.conv2, self.bn2), (self.conv3, self.bn3),
                         (self.conv4, self.bn4), (self.conv5, self.bn5)):
            x = F.relu(bn(conv(x)))
        x = F.adaptive_max_pool1d(x, 1).squeeze(-1)
        x = self.dp1(F.relu(self.bn6(self.linear1(x))))
        return self.linear2(x)


class DGCNN_semseg(nn.Module):
    """``DGCNN_semseg(args)``: upstream's S3DIS network, the name main_semseg.py:20,160 imports.
    x [B,9,N] (xyz, rgb, room-normalised xyz) -> per-point logits [B,13,N].  Three EdgeConv stages on
    the canonical (x_j - x_i, x_i) feature -- two-conv blocks (conv1+conv2, conv3+conv4) and a
    single-conv block (conv5); the first graph is built on channels 6: (``dim9``) -- then conv6 on
    the 192-channel concat, a global max, and the per-point head conv7..conv9.
    args: k, emb_dims (or emb_dim), dropout."""

    def __init__(self, args, num_classes: int = 13):
        super().__init__()
        emb = getattr(args, "emb_dims", None) or getattr(args, "emb_dim")
        self.k = args.k
        self.conv1 = _edge_block(18, 64)
        self.conv2 = _edge_block(64, 64)
        self.conv3 = _edge_block(64 * 2, 64)
        self.conv4 = _edge_block(64, 64)
        self.conv5 = _edge_block(64 * 2, 64)
        self.conv6 = nn.Sequential(nn.Conv1d(192, emb, kernel_size=1, bias=False), nn.BatchNorm1d(emb),
                                   nn.LeakyReLU(negative_slope=0.2))
        self.conv7 = nn.Sequential(nn.Conv1d(emb + 192, 512, kernel_size=1, bias=False), nn.BatchNorm1d(512),
                                   nn.LeakyReLU(negative_slope=0.2))
        self.conv8 = nn.Sequential(nn.Conv1d(512, 256, kernel_size=1, bias=False), nn.BatchNorm1d(256),
                                   nn.LeakyReLU(negative_slope=0.2))
        self.dp1 = nn.Dropout(p=float(getattr(args, "dropout", 0.5)))
        self.conv9 = nn.Conv1d(256, num_classes, kernel_size=1, bias=False)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = x.size(2)
        x1 = two_conv_edge_block(x, self.conv1, self.conv2, self.k, subtract_center=True, dim9=True)
        x2 = two_conv_edge_block(x1, self.conv3, self.conv4, self.k, subtract_center=True)
        x3, _ = edgeconv_block(x2, self.conv5, self.k, subtract_center=True)
        feats = torch.cat((x1, x2, x3), dim=1)                       # [B,192,N]
        g = self.conv6(feats).max(dim=-1, keepdim=True)[0]           # [B,emb,1]
        x = torch.cat((g.repeat(1, 1, n), feats), dim=1)             # [B,emb+192,N]
        return self.conv9(self.dp1(self.conv8(self.conv7(x))))


class IOStream:
    """The reference's tee-to-file logger (util.py:10-20), under the name main_*.py import."""

    def __init__(self, path):
        self.f = open(path, "a")

    def cprint(self, text):
        print(text)
        self.f.write(text + "\n")
        self.f.flush()

    def close(self):
        self.f.close()
