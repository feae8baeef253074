"""get_graph_feature (models/dgcnn.py:15-44, row a2) materialising kernels: time and effective GB/s.
Bytes = the [B,2C,N,k] (or [B,C,N,k] / [B,N,k,C]) tensor written once + idx and x read once (forward);
the same tensor read once + dx written (backward)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc

dev = torch.device("cuda:0")


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


for B, C, N, k in [(32, 3, 1024, 20), (32, 64, 1024, 20), (8, 64, 2048, 40), (4, 128, 4096, 20)]:
    x = (orc.synthetic_xyz(B, N, seed=1) if C == 3 else orc.synthetic_features(B, C, N, seed=1)).to(dev)
    idx = ec.ops.knn_op(x, k, True)
    for mode, name, width in ((ec.ops.GF_CONCAT, "concat [B,2C,N,k]", 2 * C), (ec.ops.GF_DISP_ONLY, "disp_only [B,C,N,k]", C),
                              (ec.ops.GF_KNN_ONLY, "knn_only [B,N,k,C]", C)):
        out = ec.ops.graph_feature_op(x, idx, mode)
        nbytes = out.numel() * 4 + idx.numel() * 4 + x.numel() * 4
        t = timeit(lambda: ec.ops.graph_feature_op(x, idx, mode))
        g = torch.randn_like(out)
        tb = timeit(lambda: ec.ops.graph_feature_bwd_op(g, idx, C, mode))
        print(f"B={B} C={C} N={N} k={k} {name:22s}: fwd {t:8.1f} us = {nbytes / t / 1e3:7.1f} GB/s   "
              f"bwd {tb:8.1f} us = {nbytes / tb / 1e3:7.1f} GB/s   ({out.numel() * 4 / 2**20:.0f} MiB tensor)", flush=True)
