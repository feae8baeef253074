"""GPU diagnostic: per-block and per-layer error of ours / oracle-fp32 against the fp64 oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from types import SimpleNamespace
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def rel(a, b):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-300)).item()


def block(B, C, N, k, Co, training, sub=False, seed=0):
    gen = torch.Generator().manual_seed(seed)
    x = orc.synthetic_xyz(B, N, seed=N) if C == 3 else orc.synthetic_features(B, C, N, seed=C + N)
    w = torch.randn(Co, 2 * C, generator=gen) / (2 * C) ** 0.5
    gamma = torch.randn(Co, generator=gen) * 0.5 + 1.0
    beta = torch.randn(Co, generator=gen) * 0.3
    rm = torch.randn(Co, generator=gen) * 0.2
    rv = torch.rand(Co, generator=gen) + 0.5
    gout = torch.randn(B, Co, N, generator=gen)
    idx = orc.knn_oracle(x, k)
    res = {}
    for name, dt in (("r32", torch.float32), ("r64", torch.float64)):
        xr = x.to(dt).clone().requires_grad_(True)
        wr, gr, br = (t.to(dt).clone().requires_grad_(True) for t in (w, gamma, beta))
        yr = orc.edgeconv_block_oracle(xr, wr, gr, br, rm.to(dt).clone(), rv.to(dt).clone(), k, training,
                                       idx=idx, subtract_center=sub)
        (yr * gout.to(dt)).sum().backward()
        res[name] = (yr.detach(), xr.grad, wr.grad, gr.grad, br.grad)
    xg = x.to(dev).requires_grad_(True)
    wg, gg, bg = (t.to(dev).requires_grad_(True) for t in (w, gamma, beta))
    y = ec.edgeconv(xg, idx.to(dev).int(), wg, gg, bg, rm.to(dev), rv.to(dev), None, training, 0.1, 1e-5,
                    0.2, sub)
    (y * gout.to(dev)).sum().backward()
    ours = (y, xg.grad, wg.grad, gg.grad, bg.grad)
    names = ("out", "dx", "dW", "dgamma", "dbeta")
    print(f"block B={B} C={C} N={N} k={k} Co={Co} train={training} sub={sub}")
    for n, o, a, b in zip(names, ours, res["r32"], res["r64"]):
        print(f"   {n:7s} ours-vs-64 {rel(o, b):.2e}   ref32-vs-64 {rel(a, b):.2e}")


def model():
    torch.manual_seed(5)
    args = SimpleNamespace(emb_dim=1024, k=20)
    net = ec.DGCNN(args).to(dev).train()
    net.record_idx = True
    sd = {k: v.cpu() for k, v in net.state_dict().items()}
    x = orc.synthetic_xyz(2, 1024, seed=1)
    xg = x.to(dev).requires_grad_(True)
    y = net(xg)
    y.square().mean().backward()
    idx_list = [i.long().cpu() for i in net.last_idx]
    out = {}
    for name, dt in (("r32", torch.float32), ("r64", torch.float64)):
        ref = orc.DGCNNOracle(args).to(dt)
        ref.load_state_dict(sd)
        ref.train()
        xr = x.detach().clone().to(dt).requires_grad_(True)
        yr = ref(xr, idx_list=idx_list)
        yr.square().mean().backward()
        out[name] = (yr.detach(), xr.grad, {n: p.grad for n, p in ref.named_parameters()})
    print("model: out ours-vs-64 %.2e ref32-vs-64 %.2e" % (rel(y, out["r64"][0]), rel(out["r32"][0], out["r64"][0])))
    print("model: dx  ours-vs-64 %.2e ref32-vs-64 %.2e" % (rel(xg.grad, out["r64"][1]), rel(out["r32"][1], out["r64"][1])))
    for n, p in net.named_parameters():
        print(f"   grad {n:16s} ours-vs-64 {rel(p.grad, out['r64'][2][n]):.2e}  ref32-vs-64 "
              f"{rel(out['r32'][2][n], out['r64'][2][n]):.2e}")


if __name__ == "__main__":
    for tr in (True, False):
        block(2, 3, 1024, 20, 64, tr)
        block(2, 64, 1024, 20, 64, tr)
        block(2, 64, 1024, 20, 128, tr)
        block(1, 128, 1024, 20, 256, tr)
    model()
