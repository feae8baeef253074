"""GPU diagnostic for the tensor-core kNN pipeline: raw scores vs fp64, then idx vs the FMA kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc

dev = torch.device("cuda:0")
for B, C, N in [(1, 32, 128), (2, 64, 256), (2, 64, 200), (1, 128, 384), (3, 96, 130)]:
    x = orc.synthetic_features(B, C, N, seed=C + N)
    s = ec.ops.debug_tc_scores(x.to(dev)).cpu().double()
    torch.cuda.synchronize()
    p = x.double().transpose(1, 2)
    ref = p @ p.transpose(1, 2) - 0.5 * (p ** 2).sum(-1)[:, None, :]
    err = (s - ref).abs()
    scale = (p ** 2).sum(-1).max().item()
    print(f"scores B={B} C={C} N={N}: max|err| {err.max().item():.3e} (|x|^2 max {scale:.2f}) "
          f"nan {int(torch.isnan(s).sum())}  rel {err.max().item() / scale:.2e}", flush=True)
    if err.max().item() > 1e-3 * scale:
        bad = (err > 1e-3 * scale).nonzero()
        print("   first bad entries (b,i,j):", bad[:8].tolist())
        print("   bad rows:", sorted(set(bad[:, 1].tolist()))[:20], " bad cols:", sorted(set(bad[:, 2].tolist()))[:20])

for B, C, N, k in [(2, 64, 256, 20), (4, 64, 1024, 20), (2, 128, 1024, 20), (1, 128, 2048, 40), (2, 64, 333, 16)]:
    x = orc.synthetic_features(B, C, N, seed=1 + C + N).to(dev)
    os.environ["ECB200_KNN"] = "fma"
    a = ec.knn(x, k)
    os.environ["ECB200_KNN"] = "auto"
    b = ec.knn(x, k)
    torch.cuda.synchronize()
    same_rows = (a.sort(-1)[0] == b.sort(-1)[0]).all(-1).float().mean().item()
    print(f"knn tc-vs-fma B={B} C={C} N={N} k={k}: identical sets in {100 * same_rows:.3f}% of rows, "
          f"identical order {100 * (a == b).all(-1).float().mean().item():.3f}%", flush=True)
