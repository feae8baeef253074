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

from .dgcnn import DGCNN
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
