"""CPU oracle for the DGCNN EdgeConv hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under
``dgcnn.pytorch_b200/`` imports or falls back to it.

It restates, in plain torch ops on whatever device/dtype the inputs carry
(CPU fp32 for parity, CPU fp64 for gradchecks), the algorithm of the
reference file ``/root/reference/models/dgcnn.py``:

* ``neg_sqdist`` / ``knn_oracle``      <- ``knn``               dgcnn.py:6-12
* ``graph_feature_oracle``             <- ``get_graph_feature`` dgcnn.py:15-44
* ``edgeconv_block_oracle``            <- ``conv{n}`` Sequential + max over k
                                          dgcnn.py:54-73 applied at :84-98
* ``DGCNNOracle``                      <- ``class DGCNN``       dgcnn.py:47-103
* ``two_conv_edge_block_oracle``       <- ``PositionEmbedding`` conv1 -> conv2 -> max
                                          models/layers.py:45-52 (row f-1, next round)
* ``embed_pool_oracle``                <- conv5's BatchNorm2d + LeakyReLU (dgcnn.py:75-78,
                                          :102) + the cls head's max | avg pooling

Parity pin: the reference ships no tests or golden vectors (SURVEY.md §4,
§8c), so the pins are generated HERE by ``oracle/make_golden.py``, which
imports the unmodified reference from ``/root/reference``, asserts this
restatement reproduces it bit-for-bit on CPU, and commits the reference's
outputs as fixtures under ``tests/golden/``.  ``tests/test_oracle.py`` re-checks
the oracle against those fixtures on every run.

Extras the reference does not have (needed to test a fused implementation):
an ``idx=`` override so the EdgeConv arithmetic can be compared on an
identical neighbour graph, and ``subtract_center=True`` for the canonical
``(x_j - x_i, x_i)`` edge feature (reference notebook test.ipynb cell 7; the
``.py`` fork uses ``(x_j, x_i)``, see SURVEY.md §0 trap 1).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

EDGE_WIDTHS = ((3, 64), (64, 64), (64, 128), (128, 256))  # dgcnn.py:54-73


# --------------------------------------------------------------------------- kNN
def neg_sqdist(x: torch.Tensor) -> torch.Tensor:
    """[B,C,N] -> [B,N,N], D[i,j] = -|xi|^2 + 2 xi.xj - |xj|^2 in the
    reference's *expanded* evaluation order (dgcnn.py:7-9)."""
    gram = torch.matmul(x.transpose(2, 1).contiguous(), x)      # :7  x^T x
    inner = -2 * gram                                           # :7
    sq = (x ** 2).sum(dim=1, keepdim=True)                      # :8  [B,1,N]
    return -sq - inner - sq.transpose(2, 1).contiguous()       # :9


def knn_oracle(x: torch.Tensor, k: int) -> torch.Tensor:
    """int64 [B,N,k]: the k largest D per row, nearest first (dgcnn.py:11)."""
    return neg_sqdist(x).topk(k=k, dim=-1)[1]


def exact_sqdist64(x: torch.Tensor) -> torch.Tensor:
    """fp64 direct-form squared distances [B,N,N]; the tie arbiter of the kNN
    parity tests (not part of the reference)."""
    p = x.detach().double().transpose(1, 2)                     # [B,N,C]
    return ((p[:, :, None, :] - p[:, None, :, :]) ** 2).sum(-1)


# ------------------------------------------------------------- neighbour gather
def gather_rows(pts: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """pts [B,N,C] point-major, idx [B,N,k] (per-cloud indices) -> x_j rows
    [B,N,k,C] (dgcnn.py:22-34: flat index = idx + b*N into the [B*N, C] view)."""
    B, N, C = pts.shape
    k = idx.shape[-1]
    base = torch.arange(B, device=pts.device).view(B, 1, 1) * N  # :22-23
    flat = (idx + base).reshape(-1)                              # :25-27
    return pts.reshape(B * N, C)[flat, :].view(B, N, k, C)       # :33-34


def graph_feature_oracle(x: torch.Tensor, k: int = 20, knn_only: bool = False,
                         disp_only: bool = False, idx: Optional[torch.Tensor] = None,
                         subtract_center: bool = False) -> torch.Tensor:
    """dgcnn.py:15-44.  Default: [B,2C,N,k] with channels 0..C-1 = x_j and
    C..2C-1 = x_i.  knn_only: [B,N,k,C] of x_j.  disp_only: [B,C,N,k] of
    x_j - x_i.  ``subtract_center`` gives the canonical (x_j - x_i, x_i)."""
    if x.dim() != 3:
        raise ValueError("expected x of shape [B, C, N]")
    B, C, N = x.shape
    if idx is None:
        idx = knn_oracle(x, k)                                  # :17
    pts = x.transpose(2, 1).contiguous()                        # :31  [B,N,C]
    nbr = gather_rows(pts, idx)                                 # [B,N,k,C]
    ctr = pts.view(B, N, 1, C).repeat(1, 1, k, 1)               # :35
    if knn_only:
        return nbr                                              # :37-38
    if disp_only:
        return (nbr - ctr).permute(0, 3, 1, 2).contiguous()     # :39-40
    first = nbr - ctr if subtract_center else nbr
    return torch.cat((first, ctr), dim=3).permute(0, 3, 1, 2).contiguous()  # :42


# ------------------------------------------------------------- one EdgeConv block
def edgeconv_block_oracle(x: torch.Tensor, weight: torch.Tensor, gamma: torch.Tensor,
                          beta: torch.Tensor, running_mean: Optional[torch.Tensor],
                          running_var: Optional[torch.Tensor], k: int, training: bool,
                          momentum: float = 0.1, eps: float = 1e-5, slope: float = 0.2,
                          idx: Optional[torch.Tensor] = None,
                          subtract_center: bool = False) -> torch.Tensor:
    """get_graph_feature -> Conv2d(2C,Co,1,bias=False) -> BatchNorm2d ->
    LeakyReLU(slope) -> max over k (dgcnn.py:84-86 with :54-58).
    ``weight`` is the Conv2d weight [Co,2C,1,1] (or [Co,2C]).  In training mode
    running_mean/var are updated in place exactly as nn.BatchNorm2d does."""
    gf = graph_feature_oracle(x, k=k, idx=idx, subtract_center=subtract_center)
    z = F.conv2d(gf, weight.reshape(weight.shape[0], -1, 1, 1))
    z = F.batch_norm(z, running_mean, running_var, gamma, beta, training, momentum, eps)
    z = F.leaky_relu(z, negative_slope=slope)
    return z.max(dim=-1, keepdim=False)[0]


# ----------------------------------------------------------------- the backbone
def _block(cin: int, cout: int) -> nn.Sequential:
    return nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False),
                         nn.BatchNorm2d(cout),
                         nn.LeakyReLU(negative_slope=0.2))


class DGCNNOracle(nn.Module):
    """Same parameters / state_dict keys as the reference ``DGCNN``
    (conv{1..5}.0.weight, conv{n}.1.*; dgcnn.py:48-78), forward restated from
    dgcnn.py:80-103.  ``forward(x, idx_list=...)`` runs each EdgeConv layer on a
    caller-supplied graph; ``last_idx`` records the graphs actually used."""

    def __init__(self, args):
        super().__init__()
        self.emb_dims = getattr(args, "emb_dim", None) or getattr(args, "emb_dims")
        self.k = args.k
        for n, (c, co) in enumerate(EDGE_WIDTHS, start=1):
            setattr(self, f"conv{n}", _block(2 * c, co))
        self.conv5 = _block(sum(co for _, co in EDGE_WIDTHS), self.emb_dims)
        self.last_idx: List[torch.Tensor] = []

    def edge_layers(self) -> Sequence[nn.Sequential]:
        return [self.conv1, self.conv2, self.conv3, self.conv4]

    def forward(self, x: torch.Tensor, idx_list: Optional[Sequence[torch.Tensor]] = None,
                subtract_center: bool = False) -> torch.Tensor:
        B, _, N = x.shape
        feats = []
        self.last_idx = []
        h = x
        for layer, seq in enumerate(self.edge_layers()):
            idx = idx_list[layer] if idx_list is not None else knn_oracle(h, self.k)
            self.last_idx.append(idx)
            gf = graph_feature_oracle(h, k=self.k, idx=idx, subtract_center=subtract_center)
            h = seq(gf).max(dim=-1, keepdim=False)[0]           # :85-86 etc.
            feats.append(h)
        cat = torch.cat(feats, dim=1).unsqueeze(-1)             # :100
        return self.conv5(cat).view(B, -1, N)                   # :102


def make_dgcnn_oracle(emb_dim: int = 1024, k: int = 20) -> DGCNNOracle:
    return DGCNNOracle(SimpleNamespace(emb_dim=emb_dim, k=k))


class DGCNNClsOracle(nn.Module):
    """DGCNN-cls for the CPU reference arm of bench.py: the reference backbone
    (restated above) + upstream DGCNN's classification head (global max+avg pool ->
    2*emb -> 512 -> 256 -> classes), as BASELINE.md §3.4 defines the headline model.
    Same submodule names as the product's ``DGCNN_cls`` so state_dicts interchange."""

    def __init__(self, args, output_channels: int = 40):
        super().__init__()
        self.backbone = DGCNNOracle(args)
        emb = self.backbone.emb_dims
        head = nn.Module()
        head.linear1 = nn.Linear(emb * 2, 512, bias=False)
        head.bn6 = nn.BatchNorm1d(512)
        head.dp1 = nn.Dropout(p=float(getattr(args, "dropout", 0.5)))
        head.linear2 = nn.Linear(512, 256)
        head.bn7 = nn.BatchNorm1d(256)
        head.dp2 = nn.Dropout(p=float(getattr(args, "dropout", 0.5)))
        head.linear3 = nn.Linear(256, output_channels)
        self.head = head

    def forward(self, x: torch.Tensor, idx_list=None) -> torch.Tensor:
        h = self.head
        f = self.backbone(x, idx_list=idx_list)
        b = f.size(0)
        f = torch.cat((F.adaptive_max_pool1d(f, 1).view(b, -1),
                       F.adaptive_avg_pool1d(f, 1).view(b, -1)), 1)
        f = h.dp1(F.leaky_relu(h.bn6(h.linear1(f)), negative_slope=0.2))
        f = h.dp2(F.leaky_relu(h.bn7(h.linear2(f)), negative_slope=0.2))
        return h.linear3(f)


def two_conv_edge_block_oracle(x: torch.Tensor, k: int, block1: nn.Sequential, block2: nn.Sequential,
                               idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The two-conv edge block of the reference's ``PositionEmbedding`` (models/layers.py:45-52;
    also the shape of upstream part-seg / sem-seg EdgeConv blocks), SURVEY.md §8 row f-1:
        get_graph_feature(x, k) -> conv1 (Conv2d 1x1 + BN2d + LeakyReLU) -> conv2 (same) -> max over k
    x [B,C,N] -> [B,Co2,N].  ``block1`` / ``block2`` are the reference's own Sequentials.  The first
    conv splits as W.[x_j ; x_i] = U_j + V_i like every EdgeConv layer; the second is a genuine
    per-edge GEMM on a tensor a fused kernel would never materialise.  Not built on the GPU in
    round 1: this restatement and its fixture pin the row for the next round."""
    gf = graph_feature_oracle(x, k=k, idx=idx)          # [B,2C,N,k]   layers.py:45
    t = block1(gf)                                      # layers.py:48
    t = block2(t)                                       # layers.py:50
    return t.max(dim=-1, keepdim=False)[0]              # layers.py:52


def embed_pool_oracle(z: torch.Tensor, B: int, N: int, gamma: torch.Tensor, beta: torch.Tensor,
                      running_mean: Optional[torch.Tensor], running_var: Optional[torch.Tensor],
                      training: bool, slope: float = 0.2, eps: float = 1e-5,
                      momentum: float = 0.1) -> torch.Tensor:
    """conv5's BatchNorm2d + LeakyReLU (dgcnn.py:75-78 applied at :102) on conv5's raw output
    ``z`` [B*N, E] (point-major rows), followed by the pooling upstream DGCNN_cls applies to the
    embedding (adaptive_max_pool1d | adaptive_avg_pool1d over the N points, as in
    ``DGCNNClsOracle.forward``) -> [B, 2E].  Checker for the fused embed_pool kernels."""
    E = z.shape[1]
    zz = z.view(B, N, E).permute(0, 2, 1).unsqueeze(-1)                 # [B,E,N,1] as conv5 emits it
    y = F.batch_norm(zz, running_mean, running_var, gamma, beta, training, momentum, eps)
    y = F.leaky_relu(y, slope).squeeze(-1)                               # [B,E,N]
    return torch.cat((F.adaptive_max_pool1d(y, 1).view(B, -1), F.adaptive_avg_pool1d(y, 1).view(B, -1)), 1)


def smoothed_ce_oracle(pred: torch.Tensor, gold: torch.Tensor, eps: float = 0.2) -> torch.Tensor:
    """Label-smoothed cross entropy of the reference (loss.py:4-21)."""
    gold = gold.contiguous().view(-1)
    n_class = pred.size(1)
    one_hot = torch.zeros_like(pred).scatter(1, gold.view(-1, 1), 1)
    one_hot = one_hot * (1 - eps) + (1 - one_hot) * eps / (n_class - 1)
    return -(one_hot * F.log_softmax(pred, dim=1)).sum(dim=1).mean()


# ------------------------------------------------- synthetic inputs
def hog_oracle(x: torch.Tensor, idx: torch.Tensor, canonical_sign: bool = False) -> torch.Tensor:
    """compute_hog_1x1 (models/model_partseg.py:15-92) restated on the CPU, line by line, for a GIVEN
    neighbour list ``idx`` [B,N,k] (the reference calls knn(x, k) itself, :26).  x [B,3,N] -> [B,N,18].

    Kept bug-compatible: the gathers at :28-30 and :51-54 index a [B*N, 3] VIEW of the channel-major
    memory with indices that carry no per-cloud offset, so every cloud reads rows 0..N-1 of that view
    (the first 3N floats of the batch) and the gradients of cloud 0.

    ``canonical_sign``: np.linalg.svd (:36) fixes the sign of each right singular vector by LAPACK's
    internals, and the zenith angle acos(v_z) (:56) is NOT invariant under v -> -v.  With
    canonical_sign the leading vector is flipped to v_z >= 0 (ties: v_y, then v_x): the convention the
    GPU kernel (ecb200_hog_1x1) uses; without it this function is the reference bit for bit."""
    import numpy as np
    import torch.nn.functional as F
    batch_size, num_pts, k = idx.shape
    nn_idx = idx.reshape(-1)                                                       # :26
    x_nn = x.contiguous().view(batch_size * num_pts, -1)[nn_idx, :].view(batch_size, num_pts, k, 3)   # :28-30
    mean = x_nn.mean(dim=2, keepdim=True)                                          # :32
    centered = x_nn - mean                                                         # :33
    _, s, v = np.linalg.svd(centered.detach().cpu().numpy(), full_matrices=False)  # :36-37
    if canonical_sign:
        v0 = v[:, :, 0, :]
        neg = (v0[..., 2] < 0) | ((v0[..., 2] == 0) & ((v0[..., 1] < 0) | ((v0[..., 1] == 0) & (v0[..., 0] < 0))))
        v[:, :, 0, :] = np.where(neg[..., None], -v0, v0)
    v = torch.from_numpy(v)                                                        # :39
    s = torch.from_numpy(np.sqrt(s))                                               # :40
    gradients = v[:, :, 0]                                                         # :49
    magnitudes = s[:, :, 0].unsqueeze(-1)                                          # :50
    gradients_nn = gradients.view(batch_size * num_pts, -1)[nn_idx, :].view(batch_size, num_pts, k, 3)    # :53-54
    magnitudes_nn = magnitudes.view(batch_size * num_pts, -1)[nn_idx, :].view(batch_size, num_pts, k, 1)  # :55-56
    zenith = torch.acos(gradients_nn[:, :, :, 2]).unsqueeze(-1) * 180 / np.pi      # :58
    azimuth = torch.atan(gradients_nn[:, :, :, 1] / gradients_nn[:, :, :, 0]).unsqueeze(-1) * 180 / np.pi   # :59-60
    cells = torch.cat((zenith.int(), azimuth.int(), magnitudes_nn), dim=-1)        # :62
    cells[cells < 0] += 180                                                        # :64
    histogram = torch.zeros((batch_size, num_pts, 9, 2))                           # :66-75
    bins = torch.floor(cells[:, :, :, :2] / 20.0 - 0.5) % 9                        # :77
    width = 20.0
    num_bins = 9
    first_centers = width * ((bins + 1) % num_bins + 0.5)                          # :81
    first_votes = cells[:, :, :, 2].unsqueeze(-1) * ((first_centers - cells[:, :, :, :2]) % 180) / width   # :82-83
    second_centers = width * (bins + 0.5)                                          # :85
    second_votes = cells[:, :, :, 2].unsqueeze(-1) * ((cells[:, :, :, :2] - second_centers) % 180) / width  # :86-87
    for c in range(9):                                                             # :88-90
        histogram[:, :, c] += (first_votes * (bins == c)).sum(dim=2)
        histogram[:, :, (c + 1) % 9] += (second_votes * (bins == c)).sum(dim=2)
    histogram = F.normalize(histogram, p=2.0, dim=2)                               # :91
    return histogram.view(batch_size, num_pts, -1)                                 # :92-93


def synthetic_xyz(B: int, N: int, seed: int = 1, device="cpu") -> torch.Tensor:
    """ModelNet40-shape clouds: centred, scaled into the unit ball, [B,3,N]."""
    g = torch.Generator().manual_seed(seed)
    p = torch.randn(B, N, 3, generator=g)
    p = p - p.mean(dim=1, keepdim=True)
    p = p / p.norm(dim=2).amax(dim=1).view(B, 1, 1)
    return p.permute(0, 2, 1).contiguous().to(device)


def synthetic_features(B: int, C: int, N: int, seed: int = 1, device="cpu") -> torch.Tensor:
    """Post-activation-like feature clouds [B,C,N]."""
    g = torch.Generator().manual_seed(seed)
    return F.leaky_relu(torch.randn(B, C, N, generator=g), 0.2).to(device)


# ------------------------------------------------- kNN comparison with the tie rule
def knn_mismatch_report(x: torch.Tensor, idx_test: torch.Tensor, idx_ref: torch.Tensor,
                        rel_eps: float = 1e-6) -> dict:
    """Compare two kNN results as per-row SETS.  A row may differ only where the
    members that differ are tied: for every j in test\\ref and j' in ref\\test,
    |d64(i,j) - d64(i,j')| <= rel_eps * (|x_i|^2 + max(|x_j|^2, |x_j'|^2)),
    with d64 the exact fp64 squared distance (the scale is the magnitude that
    enters the reference's expanded-form cancellation, dgcnn.py:9).
    Returns counts; ``bad_rows`` must be 0 for parity."""
    B, C, N = x.shape
    d = exact_sqdist64(x.cpu())
    sq = (x.detach().cpu().double() ** 2).sum(1)                 # [B,N]
    a = idx_test.cpu().long().sort(dim=-1)[0]
    r = idx_ref.cpu().long().sort(dim=-1)[0]
    differing = (a != r).any(-1)                                # [B,N]
    rows = differing.nonzero(as_tuple=False)
    bad = 0
    worst = 0.0
    for b, i in rows.tolist():
        sa, sr = set(a[b, i].tolist()), set(r[b, i].tolist())
        if len(sa) != a.shape[-1]:
            bad += 1                                            # duplicate neighbour
            continue
        only_a, only_r = sorted(sa - sr), sorted(sr - sa)
        ok = True
        for j in only_a:
            for jr in only_r:
                gap = abs(float(d[b, i, j] - d[b, i, jr]))
                scale = float(sq[b, i] + max(sq[b, j], sq[b, jr]))
                worst = max(worst, gap / max(scale, 1e-30))
                if gap > rel_eps * scale:
                    ok = False
        bad += (not ok)
    return {"rows": B * N, "differing_rows": int(differing.sum()), "bad_rows": bad,
            "worst_rel_gap": worst}
