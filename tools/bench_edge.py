"""Isolated timings of the graph kernels of one EdgeConv layer (forward gather, backward prep /
reverse graph / scatter / dense) at the four DGCNN layer widths; CUDA events, L2 not flushed
(the rows these kernels gather are meant to be L2-resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from ctypes import c_void_p
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc

L = ec._lib
dev = torch.device("cuda:0")
P = lambda t: None if t is None else c_void_p(t.data_ptr())
B, N, k = 32, 1024, 20
if len(sys.argv) > 1:
    B, N, k = (int(v) for v in sys.argv[1].split(","))
M = B * N


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n


x = orc.synthetic_xyz(B, N, seed=1).to(dev)
idx = ec.ops.knn_op(x, k, False)
st = c_void_p(torch.cuda.current_stream().cuda_stream)
tot = {}
for Co in (64, 64, 128, 256):
    Y = torch.randn(M, 2 * Co, device=dev)
    gamma = torch.randn(Co, device=dev)
    sel = torch.empty(M, Co, device=dev); arg = torch.empty(M, Co, device=dev, dtype=torch.uint8)
    esum = torch.empty(M, Co, device=dev)
    stats = torch.zeros(2 * Co + 1, device=dev, dtype=torch.float64)
    aff = torch.randn(4, Co, device=dev)
    mean, invstd, a, b = (c_void_p(aff.data_ptr() + 4 * Co * r) for r in range(4))
    cc = torch.randn(2, Co, device=dev) * 1e-3
    c1, c2 = c_void_p(cc.data_ptr()), c_void_p(cc.data_ptr() + 4 * Co)
    gout = torch.randn(B, Co, N, device=dev); gpm = torch.randn(M, Co, device=dev)
    g = torch.empty(M, Co, device=dev); bst = torch.zeros(2 * Co, device=dev, dtype=torch.float64)
    rowptr = torch.empty(M + 1, device=dev, dtype=torch.int32); src = torch.empty(M * k, device=dev, dtype=torch.int32)
    cur = torch.empty(M, device=dev, dtype=torch.int32)
    dYs = torch.empty(2, M, 2 * Co, device=dev); dU = torch.zeros(M, Co, device=dev)
    out = torch.empty(B, Co, N, device=dev); opm = torch.empty(M, Co, device=dev)
    fns = {
        "gather": lambda: L.call("ecb200_edge_gather", P(Y), P(idx), P(gamma), B, N, k, Co, P(sel), P(arg), P(esum), P(stats), st),
        "apply": lambda: L.call("ecb200_edge_apply", P(sel), a, b, 0.2, B, N, Co, P(out), P(opm), Co, st),
        "prep": lambda: L.call("ecb200_bwd_prep", P(gout), P(gpm), Co, P(sel), a, b, mean, invstd, 0.2, B, N, Co, P(g), P(bst), st),
        "reverse": lambda: L.call("ecb200_reverse_graph", P(idx), B, N, k, P(rowptr), P(src), P(cur), st),
        "scatter": lambda: L.call("ecb200_bwd_scatter", P(g), P(esum), P(arg), P(idx), a, mean, c1, c2, B, N, k, Co, None, P(dU), P(dYs[0]), P(dYs[1]), st),
        "dense": lambda: L.call("ecb200_bwd_dense", P(Y), P(rowptr), P(src), mean, c1, c2, 1, B, N, Co, None, P(dU), P(dYs[0]), P(dYs[1]), st),
    }
    line = f"Co={Co:3d}:"
    for name, fn in fns.items():
        t = timeit(fn)
        tot[name] = tot.get(name, 0.0) + t
        line += f"  {name} {t:6.1f}"
    gb = (4 * M * k + 4 * M * k * Co + 4 * M * Co * 3 + M * Co) / 1e3
    print(line + f"   | gather {gb / timeit(fns['gather']):6.0f} GB/s (SpMM convention)", flush=True)
print("sum over the 4 layers (us): " + "  ".join(f"{n} {t:.1f}" for n, t in tot.items()), flush=True)
