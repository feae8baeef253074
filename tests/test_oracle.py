"""The oracle (oracle/edgeconv_oracle.py) against the golden vectors that
oracle/make_golden.py recorded from the unmodified reference
(/root/reference/models/dgcnn.py).  CPU only."""
import glob
import os
from types import SimpleNamespace

import pytest
import torch

import edgeconv_oracle as orc
from conftest import GOLDEN, load_golden


@pytest.fixture(autouse=True)
def _one_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)        # fixtures were recorded with one thread
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "knn_*.npz"))),
                         ids=os.path.basename)
def test_knn_matches_reference(path):
    g = load_golden(os.path.basename(path))
    idx = orc.knn_oracle(g["x"], g["k"])
    assert idx.dtype == torch.int64
    assert torch.equal(idx, g["idx"].long())
    # self is the nearest neighbour of every point
    n = g["x"].shape[2]
    assert torch.equal(idx[..., 0], torch.arange(n).expand_as(idx[..., 0]))
    rep = orc.knn_mismatch_report(g["x"], idx, g["idx"])
    assert rep["differing_rows"] == 0 and rep["bad_rows"] == 0


def test_graph_feature_layouts():
    g = load_golden("graph_feature_B2_C5_N48_k6.npz")
    x, k = g["x"], g["k"]
    assert torch.equal(orc.graph_feature_oracle(x, k), g["full"])
    assert torch.equal(orc.graph_feature_oracle(x, k, knn_only=True), g["knn_only"])
    assert torch.equal(orc.graph_feature_oracle(x, k, disp_only=True), g["disp_only"])
    # idx override reproduces the same tensor
    assert torch.equal(orc.graph_feature_oracle(x, k, idx=g["idx"].long()), g["full"])
    # canonical form differs from the fork form only by the centre subtraction
    can = orc.graph_feature_oracle(x, k, subtract_center=True)
    assert torch.equal(can[:, 5:], g["full"][:, 5:])
    assert torch.equal(can[:, :5], g["disp_only"])


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_block_matches_reference(mode):
    g = load_golden("block_B3_C6_N80_k7_Co16.npz")
    x = g["x"].clone().requires_grad_(True)
    w, ga, be = (g[n].clone().requires_grad_(True) for n in ("weight", "gamma", "beta"))
    rm, rv = g["running_mean0"].clone(), g["running_var0"].clone()
    y = orc.edgeconv_block_oracle(x, w, ga, be, rm, rv, g["k"], training=(mode == "train"))
    (y * g["gout"]).sum().backward()
    assert torch.equal(y, g[f"{mode}_out"])
    assert torch.equal(x.grad, g[f"{mode}_dx"])
    assert torch.equal(w.grad, g[f"{mode}_dw"])
    assert torch.equal(ga.grad, g[f"{mode}_dgamma"])
    assert torch.equal(be.grad, g[f"{mode}_dbeta"])
    assert torch.equal(rm, g[f"{mode}_running_mean"])
    assert torch.equal(rv, g[f"{mode}_running_var"])


def _load_dgcnn():
    g = load_golden("dgcnn_emb64_k8_B2_N96.npz")
    net = orc.DGCNNOracle(SimpleNamespace(emb_dim=g["emb_dim"], k=g["k"]))
    net.load_state_dict({k[3:]: v for k, v in g.items() if k.startswith("sd.")})
    return g, net


def test_dgcnn_train_step_matches_reference():
    g, net = _load_dgcnn()
    net.train()
    x = g["x"].clone().requires_grad_(True)
    y = net(x)
    (y * g["gout"]).sum().backward()
    assert torch.equal(y, g["train_out"])
    assert torch.equal(x.grad, g["train_dx"])
    for name, p in net.named_parameters():
        assert torch.equal(p.grad, g[f"grad.{name}"]), name
    for name, b in net.named_buffers():
        assert torch.equal(b, g[f"after.{name}"]), name
    for n, idx in enumerate(net.last_idx):
        assert torch.equal(idx, g[f"idx_train{n}"].long())


def test_dgcnn_eval_and_idx_override():
    g, net = _load_dgcnn()
    # the reference's eval forward was recorded after its one training step
    net.load_state_dict({k[6:]: v for k, v in g.items() if k.startswith("after.")}, strict=False)
    net.eval()
    with torch.no_grad():
        assert torch.equal(net(g["x"]), g["eval_out"])
        forced = [g[f"idx_eval{n}"].long() for n in range(4)]
        assert torch.equal(net(g["x"], idx_list=forced), g["eval_out"])


def test_dgcnn_accepts_emb_dims_alias():
    # main_cls.py:228 defines --emb_dims, models/dgcnn.py:51 reads emb_dim (SURVEY §0 trap 3)
    net = orc.DGCNNOracle(SimpleNamespace(emb_dims=32, k=4))
    assert net.conv5[0].weight.shape == (32, 512, 1, 1)


def test_split_weight_identity_fp64():
    """The algebra the fused design rests on: W.[x_j; x_i] =
    W1.x_j + W2.x_i, and max/min by sign(gamma) commutes with BN+LeakyReLU."""
    torch.manual_seed(0)
    B, C, N, k, Co = 2, 5, 40, 6, 7
    x = torch.randn(B, C, N, dtype=torch.float64)
    w = torch.randn(Co, 2 * C, dtype=torch.float64)
    gamma = torch.randn(Co, dtype=torch.float64)
    beta = torch.randn(Co, dtype=torch.float64)
    idx = orc.knn_oracle(x, k)
    ref = orc.edgeconv_block_oracle(x, w, gamma, beta, None, None, k, training=True, idx=idx)
    pts = x.transpose(1, 2)                                   # [B,N,C]
    U, V = pts @ w[:, :C].T, pts @ w[:, C:].T                 # [B,N,Co]
    e = orc.gather_rows(U.contiguous(), idx) + V[:, :, None, :]   # [B,N,k,Co]
    mean = e.mean(dim=(0, 1, 2))
    var = e.var(dim=(0, 1, 2), unbiased=False)
    a = gamma / torch.sqrt(var + 1e-5)
    b = beta - a * mean
    sel = torch.where(gamma >= 0, e.max(dim=2)[0], e.min(dim=2)[0])
    out = torch.nn.functional.leaky_relu(a * sel + b, 0.2).transpose(1, 2)
    assert torch.allclose(out, ref, rtol=1e-10, atol=1e-12)


def test_embed_pool_oracle_is_the_tail_of_the_cls_oracle():
    """embed_pool_oracle(conv5's raw output) == what DGCNNClsOracle feeds its head (conv5's
    BatchNorm + LeakyReLU, dgcnn.py:75-78,:102, then max | avg pooling over the points)."""
    from types import SimpleNamespace
    torch.manual_seed(0)
    args = SimpleNamespace(emb_dim=24, k=5, dropout=0.0)
    net = orc.DGCNNClsOracle(args).train()
    x = orc.synthetic_xyz(3, 40, seed=2)
    B, N = 3, 40
    captured = {}
    h1 = net.backbone.conv5[0].register_forward_hook(lambda m, i, o: captured.__setitem__("z", o.detach()))
    h2 = net.head.linear1.register_forward_hook(lambda m, i, o: captured.__setitem__("pooled", i[0].detach()))
    bn = net.backbone.conv5[1]
    rm0, rv0 = bn.running_mean.clone(), bn.running_var.clone()
    net(x)
    h1.remove(); h2.remove()
    z = captured["z"].squeeze(-1).permute(0, 2, 1).reshape(B * N, -1)      # [B,E,N,1] -> [B*N,E]
    pooled = orc.embed_pool_oracle(z, B, N, bn.weight.detach(), bn.bias.detach(), rm0, rv0, True)
    assert torch.allclose(pooled, captured["pooled"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(rm0, bn.running_mean) and torch.allclose(rv0, bn.running_var)


def test_two_conv_edge_block_matches_reference():
    """Row f-1 (next round): the reference PositionEmbedding's conv1 -> conv2 -> max over k
    (models/layers.py:45-52), recorded from the unmodified module."""
    g = load_golden("two_conv_block_B2_N64_k6.npz")
    b1 = torch.nn.Sequential(torch.nn.Conv2d(6, 64, 1, bias=False), torch.nn.BatchNorm2d(64), torch.nn.LeakyReLU(0.2))
    b2 = torch.nn.Sequential(torch.nn.Conv2d(64, 128, 1, bias=False), torch.nn.BatchNorm2d(128), torch.nn.LeakyReLU(0.2))
    b1.load_state_dict({k[len("sd.conv1."):]: v for k, v in g.items() if k.startswith("sd.conv1.")})
    b2.load_state_dict({k[len("sd.conv2."):]: v for k, v in g.items() if k.startswith("sd.conv2.")})
    b1.train(); b2.train()
    x = g["x"].clone().requires_grad_(True)
    t = orc.two_conv_edge_block_oracle(x, g["k"], b1, b2)
    (t * g["gout"]).sum().backward()
    assert torch.equal(t, g["out"])
    assert torch.equal(x.grad, g["dx"])
    assert torch.equal(b1[0].weight.grad, g["dw1"]) and torch.equal(b2[0].weight.grad, g["dw2"])
    assert torch.equal(b2[1].weight.grad, g["dgamma2"]) and torch.equal(b2[1].bias.grad, g["dbeta2"])


@pytest.mark.parametrize("name", ["hog_B2_N64_k8.npz", "hog_B3_N200_k20.npz"])
def test_hog_oracle_equals_the_recorded_reference(name):
    """compute_hog_1x1 (models/model_partseg.py:15-92, row f-3): the restatement reproduces what the unmodified
    reference returned (recorded by oracle/make_golden.py with use_cpu=True) bit for bit, and the
    canonical-sign variant -- the GPU kernel's convention -- is reproducible too."""
    g = load_golden(name)
    h = orc.hog_oracle(g["x"], g["idx"].long(), canonical_sign=False)
    assert torch.equal(h, g["hog_reference"])
    hc = orc.hog_oracle(g["x"], g["idx"].long(), canonical_sign=True)
    assert torch.equal(hc, g["hog_canonical"])
    assert tuple(h.shape) == (g["x"].shape[0], g["x"].shape[2], 18)
    # every histogram is L2-normalised per angle type (or all zero)
    n = hc.view(*hc.shape[:2], 9, 2).norm(dim=2)
    assert bool(((n - 1).abs() < 1e-5).logical_or(n == 0).all())
