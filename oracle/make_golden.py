"""Generate tests/golden/*.npz from the UNMODIFIED reference and pin the oracle.

Run in the build container only (needs /root/reference, which does not exist
on the GPU box):

    python oracle/make_golden.py

For every case it (1) runs the reference's own ``models/dgcnn.py`` on a seeded
CPU input, (2) asserts that ``oracle/edgeconv_oracle.py`` reproduces the result
bit-for-bit (same torch ops, same order), and (3) stores the reference's
inputs/outputs as a small fixture.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.environ.get("DGCNN_REFERENCE", "/root/reference")

sys.path.insert(0, HERE)
import edgeconv_oracle as orc  # noqa: E402


def load_reference():
    sys.path.insert(0, REF)
    import models.dgcnn as ref  # the reference's models/dgcnn.py, unmodified
    sys.path.pop(0)
    return ref


def npz(name, **arrs):
    out = {}
    for key, v in arrs.items():
        out[key] = v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)
    path = os.path.join(GOLD, name)
    np.savez_compressed(path, **out)
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.1f} KiB)")


def case_knn(ref):
    """knn(): dgcnn.py:6-12 on xyz and feature-space clouds."""
    for tag, x, k in (
        ("xyz_B2_N64_k5", orc.synthetic_xyz(2, 64, seed=1), 5),
        ("xyz_B3_N200_k20", orc.synthetic_xyz(3, 200, seed=2), 20),
        ("feat64_B2_N256_k20", orc.synthetic_features(2, 64, 256, seed=3), 20),
        ("feat128_B1_N320_k40", orc.synthetic_features(1, 128, 320, seed=4), 40),
        ("c9_B2_N96_k8", orc.synthetic_features(2, 9, 96, seed=5), 8),
    ):
        idx_ref = ref.knn(x, k)
        idx_orc = orc.knn_oracle(x, k)
        assert torch.equal(idx_ref, idx_orc), tag
        assert idx_ref.dtype == torch.int64 and tuple(idx_ref.shape) == (x.shape[0], x.shape[2], k)
        npz(f"knn_{tag}.npz", x=x, k=k, idx=idx_ref.to(torch.int32))


def case_graph_feature(ref):
    """get_graph_feature(): dgcnn.py:15-44, all three output layouts."""
    x = orc.synthetic_features(2, 5, 48, seed=7)
    k = 6
    full = ref.get_graph_feature(x, k=k)
    nbr = ref.get_graph_feature(x, k=k, knn_only=True)
    disp = ref.get_graph_feature(x, k=k, disp_only=True)
    assert torch.equal(full, orc.graph_feature_oracle(x, k=k))
    assert torch.equal(nbr, orc.graph_feature_oracle(x, k=k, knn_only=True))
    assert torch.equal(disp, orc.graph_feature_oracle(x, k=k, disp_only=True))
    assert tuple(full.shape) == (2, 10, 48, k) and tuple(nbr.shape) == (2, 48, k, 5)
    npz("graph_feature_B2_C5_N48_k6.npz", x=x, k=k, idx=ref.knn(x, k).to(torch.int32),
        full=full, knn_only=nbr, disp_only=disp)


def case_block(ref):
    """One EdgeConv block = get_graph_feature -> conv{n} -> max (dgcnn.py:84-86),
    training and eval BatchNorm, with negative BN scales (min path), values and
    gradients."""
    torch.manual_seed(11)
    B, C, N, k, Co = 3, 6, 80, 7, 16
    x = orc.synthetic_features(B, C, N, seed=11)
    seq = torch.nn.Sequential(torch.nn.Conv2d(2 * C, Co, 1, bias=False),
                              torch.nn.BatchNorm2d(Co),
                              torch.nn.LeakyReLU(0.2, inplace=True))
    with torch.no_grad():
        seq[1].weight.normal_(1.0, 0.5)
        seq[1].weight[::3] *= -1.0          # negative scales exercise the min path
        seq[1].bias.normal_(0.0, 0.3)
        seq[1].running_mean.normal_(0.0, 0.2)
        seq[1].running_var.uniform_(0.5, 1.5)
    w = seq[0].weight.detach().clone()
    gamma, beta = seq[1].weight.detach().clone(), seq[1].bias.detach().clone()
    rm0, rv0 = seq[1].running_mean.clone(), seq[1].running_var.clone()
    gout = torch.randn(B, Co, N, generator=torch.Generator().manual_seed(12))
    out = {}
    for mode in ("train", "eval"):
        seq.train(mode == "train")
        seq[1].running_mean.copy_(rm0)
        seq[1].running_var.copy_(rv0)
        xr = x.clone().requires_grad_(True)
        seq.zero_grad()
        y = seq(ref.get_graph_feature(xr, k=k)).max(dim=-1, keepdim=False)[0]
        (y * gout).sum().backward()
        # the oracle must agree with the reference path
        rm, rv = rm0.clone(), rv0.clone()
        xo = x.clone().requires_grad_(True)
        wo, go, bo = (t.clone().requires_grad_(True) for t in (w, gamma, beta))
        yo = orc.edgeconv_block_oracle(xo, wo, go, bo, rm, rv, k, training=(mode == "train"))
        (yo * gout).sum().backward()
        assert torch.equal(y, yo), mode
        assert torch.allclose(xr.grad, xo.grad, rtol=0, atol=0), mode
        assert torch.equal(seq[0].weight.grad, wo.grad.view_as(seq[0].weight.grad)), mode
        assert torch.equal(seq[1].running_mean, rm) and torch.equal(seq[1].running_var, rv)
        out.update({f"{mode}_out": y, f"{mode}_dx": xr.grad, f"{mode}_dw": seq[0].weight.grad.clone(),
                    f"{mode}_dgamma": seq[1].weight.grad.clone(), f"{mode}_dbeta": seq[1].bias.grad.clone(),
                    f"{mode}_running_mean": seq[1].running_mean.clone(),
                    f"{mode}_running_var": seq[1].running_var.clone()})
    npz("block_B3_C6_N80_k7_Co16.npz", x=x, k=k, idx=ref.knn(x, k).to(torch.int32), weight=w,
        gamma=gamma, beta=beta, running_mean0=rm0, running_var0=rv0, gout=gout, **out)


def case_dgcnn(ref):
    """class DGCNN (dgcnn.py:47-103): forward + backward in training mode, and an
    eval forward, with the per-layer kNN graphs the reference used."""
    torch.manual_seed(21)
    args = SimpleNamespace(emb_dim=64, k=8)
    net = ref.DGCNN(args)
    with torch.no_grad():
        for n in range(1, 6):
            bn = getattr(net, f"conv{n}")[1]
            bn.weight.normal_(1.0, 0.4)
            bn.bias.normal_(0.0, 0.2)
    sd = {k_: v.clone() for k_, v in net.state_dict().items()}
    o = orc.DGCNNOracle(args)
    o.load_state_dict(sd)                       # identical state_dict keys
    assert list(o.state_dict().keys()) == list(sd.keys())
    x = orc.synthetic_xyz(2, 96, seed=21)
    gout = torch.randn(2, 64, 96, generator=torch.Generator().manual_seed(22))

    net.train(); o.train()
    xr = x.clone().requires_grad_(True)
    y = net(xr)
    (y * gout).sum().backward()
    xo = x.clone().requires_grad_(True)
    yo = o(xo)
    (yo * gout).sum().backward()
    assert torch.equal(y, yo)
    assert torch.equal(xr.grad, xo.grad)
    for (n1, p1), (n2, p2) in zip(net.named_parameters(), o.named_parameters()):
        assert n1 == n2 and torch.equal(p1.grad, p2.grad), n1
    for (n1, b1), (n2, b2) in zip(net.named_buffers(), o.named_buffers()):
        assert n1 == n2 and torch.equal(b1, b2), n1
    grads = {f"grad.{n}": p.grad.clone() for n, p in net.named_parameters()}
    sd_after = {f"after.{k_}": v.clone() for k_, v in net.named_buffers()}  # BN running stats
    idx_train = [i.to(torch.int32) for i in o.last_idx]

    net.eval(); o.eval()
    with torch.no_grad():
        ye = net(x)
        assert torch.equal(ye, o(x))
    idx_eval = [i.to(torch.int32) for i in o.last_idx]
    npz("dgcnn_emb64_k8_B2_N96.npz", x=x, gout=gout, emb_dim=64, k=8,
        train_out=y, train_dx=xr.grad, eval_out=ye,
        **{f"sd.{k_}": v for k_, v in sd.items()}, **grads, **sd_after,
        **{f"idx_train{n}": t for n, t in enumerate(idx_train)},
        **{f"idx_eval{n}": t for n, t in enumerate(idx_eval)})


def case_two_conv_block():
    """Row f-1: the two-conv edge block of the reference's PositionEmbedding
    (models/layers.py:45-52), captured from the unmodified module by a forward hook on its
    third conv (whose input is the block's output after the max over k)."""
    sys.path.insert(0, REF)
    from models.layers import PositionEmbedding  # unmodified reference module
    sys.path.pop(0)
    torch.manual_seed(31)
    pe = PositionEmbedding(SimpleNamespace(k=6)).train()
    with torch.no_grad():
        for bn in (pe.bn1, pe.bn2):
            bn.weight.normal_(1.0, 0.4)
            bn.weight[::4] *= -1.0
            bn.bias.normal_(0.0, 0.2)
    sd = {k_: v.clone() for k_, v in pe.state_dict().items() if k_.startswith(("conv1.", "conv2."))}
    x = orc.synthetic_xyz(2, 64, seed=31)
    got = {}
    h = pe.conv3.register_forward_hook(lambda m, i, o: got.__setitem__("t", i[0]))
    xr = x.clone().requires_grad_(True)
    pe(xr)
    h.remove()
    t = got["t"]                                          # [B,128,N]
    gout = torch.randn(t.shape, generator=torch.Generator().manual_seed(32))
    (t * gout).sum().backward()
    # the oracle, on fresh modules carrying the same parameters and the initial running statistics
    b1 = torch.nn.Sequential(torch.nn.Conv2d(6, 64, 1, bias=False), torch.nn.BatchNorm2d(64), torch.nn.LeakyReLU(0.2))
    b2 = torch.nn.Sequential(torch.nn.Conv2d(64, 128, 1, bias=False), torch.nn.BatchNorm2d(128), torch.nn.LeakyReLU(0.2))
    b1.load_state_dict({k_[len("conv1."):]: v for k_, v in sd.items() if k_.startswith("conv1.")})
    b2.load_state_dict({k_[len("conv2."):]: v for k_, v in sd.items() if k_.startswith("conv2.")})
    b1.train(); b2.train()
    xo = x.clone().requires_grad_(True)
    to = orc.two_conv_edge_block_oracle(xo, 6, b1, b2)
    (to * gout).sum().backward()
    assert torch.equal(t, to)
    assert torch.equal(xr.grad, xo.grad)
    assert torch.equal(pe.conv1[0].weight.grad, b1[0].weight.grad) and torch.equal(pe.conv2[0].weight.grad, b2[0].weight.grad)
    assert torch.equal(pe.bn2.running_var, b2[1].running_var)
    npz("two_conv_block_B2_N64_k6.npz", x=x, k=6, gout=gout, out=t, dx=xr.grad,
        dw1=pe.conv1[0].weight.grad, dw2=pe.conv2[0].weight.grad,
        dgamma2=pe.bn2.weight.grad, dbeta2=pe.bn2.bias.grad,
        **{f"sd.{k_}": v for k_, v in sd.items()})


def case_hog():
    """compute_hog_1x1 (models/model_partseg.py:15-92, row f-3): the restatement must equal the
    unmodified reference bit for bit (use_cpu=True keeps the reference on the CPU); the fixture also
    stores the canonical-sign variant the GPU kernel is compared with."""
    sys.path.insert(0, REF)
    import models.model_partseg as ps
    sys.path.pop(0)
    for tag, B, N, k, seed in (("B2_N64_k8", 2, 64, 8, 21), ("B3_N200_k20", 3, 200, 20, 22)):
        x = orc.synthetic_xyz(B, N, seed=seed)
        ref_h = ps.compute_hog_1x1(x, k, use_cpu=True)
        idx = orc.knn_oracle(x, k)
        mine = orc.hog_oracle(x, idx, canonical_sign=False)
        assert torch.equal(ref_h, mine), f"hog {tag}: restatement differs from the reference"
        canon = orc.hog_oracle(x, idx, canonical_sign=True)
        flipped = float(((ref_h - canon).abs().amax(-1) > 1e-6).float().mean())
        print(f"hog {tag}: reference == restatement; canonical sign changes {100 * flipped:.1f} % of the points' histograms")
        npz(f"hog_{tag}.npz", x=x, k=k, idx=idx.to(torch.int32), hog_reference=ref_h, hog_canonical=canon)


def main():
    torch.set_num_threads(1)        # deterministic reduction order
    os.makedirs(GOLD, exist_ok=True)
    ref = load_reference()
    case_knn(ref)
    case_graph_feature(ref)
    case_block(ref)
    case_dgcnn(ref)
    case_two_conv_block()
    case_hog()
    print("oracle == reference on every case; fixtures written")


if __name__ == "__main__":
    main()
