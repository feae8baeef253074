"""A few launches of one tensor-core kNN kernel for ncu: python tools/profile_knn.py B,C,N,k [tf32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc
B, C, N, k = (int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "32,64,1024,20").split(","))
if len(sys.argv) > 2:
    os.environ["ECB200_KNN"] = sys.argv[2]
dev = torch.device("cuda:0")
x = (orc.synthetic_xyz(B, N, seed=1) if C <= 5 else orc.synthetic_features(B, C, N, seed=1)).to(dev)
for _ in range(6):
    idx = ec.ops.knn_op(x, k, True)
torch.cuda.synchronize()
print("ok", ec.ops.knn_tc_kind(C, N, k), tuple(idx.shape))
