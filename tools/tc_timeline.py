"""Per-role clock64() timeline of CTA (0,0) of the tensor-core kNN kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from ctypes import c_void_p
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc
L = ec._lib
B, C, N, k = (int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "32,64,1024,20".split(",")))
dev = torch.device("cuda:0")
x = orc.synthetic_features(B, C, N, seed=1).to(dev)
hi = torch.empty(B * N, C, device=dev); lo = torch.empty_like(hi); xx = torch.empty(B * N, device=dev)
idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
nb = L.load().ecb200_knn_tc_workspace_bytes(B, N, k)
ws = torch.empty(nb, device=dev, dtype=torch.uint8)
ntile = B * ((N + 127) // 128)
tl = torch.zeros(6 * 256 + 3 * ntile, device=dev, dtype=torch.int64)
P = lambda t: c_void_p(t.data_ptr())
st = c_void_p(torch.cuda.current_stream().cuda_stream)
L.call("ecb200_split_tf32", P(x), B, C, N, P(hi), P(lo), P(xx), st)
for _ in range(3):
    L.call("ecb200_debug_tc_timeline", P(hi), P(lo), P(xx), B, C, N, k, P(idx), P(ws), P(tl), st)
torch.cuda.synchronize()
allc = tl.cpu()[6 * 256:].view(ntile, 3)
t = tl.cpu()[:6 * 256].view(6, 256)
t0 = int(t[5, 0])
nct = (N + 127) // 128; nkb = C // 32
rel = lambda v: (int(v) - t0) if int(v) else None
print("producer stage-acquire times (first 20, last 4):", [rel(v) for v in t[0, :20]], [rel(v) for v in t[0, 2 * nct * nkb - 4:2 * nct * nkb]])
print("mma per tile (start, commit):")
for i in range(2 * nct):
    print("   tile", i, rel(t[1, 2 * i]), rel(t[1, 2 * i + 1]))
for g in (0, 1):
    print(f"epilogue group {g}: per use (begin, hx-barrier done, t_full acquired, done)")
    for u in range(nct):
        print("   use", u, [rel(v) for v in t[2 + g, 4 * u:4 * u + 4]])
    print("   pass A end / pass B end / - / keys converted / ranked:", [rel(v) for v in t[4, 8 * g:8 * g + 5]])
    print("   tau stage: start / sorted / done:", [rel(v) for v in t[4, 8 * g + 5:8 * g + 8]])
    print("   rank stage: fast ranks done / barrier / slow path + barrier:", [rel(v) for v in t[5, 8 + 4 * g:11 + 4 * g]])

sv = t[5, 128:256]
c0, c1 = (sv >> 32).float(), (sv & 0xffffffff).float()
tot = c0 + c1
print(f"survivors per row (CTA 0,0): mean {tot.mean():.1f} max {tot.max():.0f}; per thread mean {c0.mean():.1f}/{c1.mean():.1f} max {c0.max():.0f}/{c1.max():.0f}; "
      f"per-warp max of total: {[int(tot[w*32:(w+1)*32].max()) for w in range(4)]}")
# every CTA: start / duration (us) relative to the first start, by SM
st0 = int(allc[:, 0].min())
dur = (allc[:, 1] - allc[:, 0]).double() / 1e3
beg = (allc[:, 0] - st0).double() / 1e3
print(f"all CTAs: {ntile} on {len(set(allc[:, 2].tolist()))} SMs; kernel span {(int(allc[:, 1].max()) - st0) / 1e3:.1f} us; "
      f"CTA duration us: min {dur.min():.1f} median {dur.median():.1f} max {dur.max():.1f}")
first = beg < 1.0
print(f"  first wave: {int(first.sum())} CTAs, duration median {dur[first].median():.1f} max {dur[first].max():.1f}; "
      f"later CTAs: start median {beg[~first].median() if (~first).any() else 0:.1f} duration median "
      f"{dur[~first].median() if (~first).any() else 0:.1f} max {dur[~first].max() if (~first).any() else 0:.1f}")
