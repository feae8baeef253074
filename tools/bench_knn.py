"""Isolated kNN timing (CUDA events) per (B, C, N, k); used for ncu captures too."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc

cfgs = [(32, 3, 1024, 20), (32, 64, 1024, 20), (32, 128, 1024, 20), (32, 3, 2048, 40), (32, 64, 2048, 40)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
dev = torch.device("cuda:0")
for B, C, N, k in cfgs:
    x = (orc.synthetic_xyz(B, N, seed=1) if C == 3 else orc.synthetic_features(B, C, N, seed=1)).to(dev)
    for srt in (False, True):
        for _ in range(3):
            ec.ops.knn_op(x, k, srt)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(10):
            ec.ops.knn_op(x, k, srt)
        e.record()
        torch.cuda.synchronize()
        us = s.elapsed_time(e) * 100
        print(f"knn B={B} C={C} N={N} k={k} sorted={srt}: {us:8.1f} us  {B*N*N/us/1e6:7.3f} Tpairs/s "
              f"{2*B*N*N*C/us/1e6:7.2f} TFLOP/s")
