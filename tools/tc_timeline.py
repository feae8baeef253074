"""Per-role clock64() timeline of CTA (0,0) of the tensor-core kNN kernel."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from ctypes import c_void_p
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc
L = ec._lib
dev = torch.device("cuda:0")
if len(sys.argv) > 1 and sys.argv[1].startswith("model"):
    # the input of EdgeConv layer `layer` (2..4) of a DGCNN-cls forward on synthetic clouds: real activations
    from types import SimpleNamespace
    layer = int(sys.argv[1].split(":")[1])
    B, N, k = 32, 1024, 20
    torch.manual_seed(1)
    net = ec.DGCNN_cls(SimpleNamespace(emb_dims=1024, k=k, dropout=0.5)).to(dev).train()
    seen = []
    orig = ec.ops.split_f16_op
    def spy(x, *rest):
        seen.append(x.detach().clone())
        return orig(x, *rest)
    ec.ops.split_f16_op = spy
    import dgcnn_pytorch_b200.dgcnn as dg
    dg.ops.split_f16_op = spy
    with torch.no_grad():
        net(orc.synthetic_xyz(B, N, seed=100).to(dev))
    x = seen[layer - 2].contiguous()
    C = x.shape[1]
    print(f"layer {layer} input: C={C}, mean {x.mean():.3f} std {x.std():.3f}")
else:
    B, C, N, k = (int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "32,64,1024,20".split(",")))
    x = (orc.synthetic_xyz(B, N, seed=1) if C == 3 else orc.synthetic_features(B, C, N, seed=1)).to(dev)
KIND = os.environ.get("TL_KIND", "xyz" if C <= 5 else "f16")   # f16 | tf32 | xyz
hi = torch.empty(B * N, C, device=dev); lo = torch.empty_like(hi); xx = torch.empty(B * N, device=dev)
idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
nb = L.load().ecb200_knn_tc_workspace_bytes(B, N, k)
ws = torch.empty(nb, device=dev, dtype=torch.uint8)
ntile = B * ((N + 127) // 128)
tl = torch.zeros(6 * 256 + 3 * ntile, device=dev, dtype=torch.int64)
P = lambda t: None if t is None else c_void_p(t.data_ptr())
st = c_void_p(torch.cuda.current_stream().cuda_stream)
L.call("ecb200_split_tf32", P(x), B, C, N, P(hi), P(lo), P(xx), st)
if KIND == "f16":
    hh, hl, nb, xxs, cmax = ec.ops.split_f16_op(x, False)[:5]
    launch = lambda t: L.call("ecb200_knn_tc_f16", P(hh), P(hl), P(nb), P(xxs), P(cmax), B, C, N, k, P(idx), t, st)
elif KIND == "xyz":
    rows = torch.empty(2, B * N, 64, device=dev, dtype=torch.float16); xxs = torch.empty(B * N, device=dev)
    cmax = torch.empty(B * ((N + 31) // 32), device=dev)
    L.call("ecb200_pack_xyz_f16", P(x), B, C, N, P(rows[0]), P(rows[1]), P(xxs), P(cmax), st)
    launch = lambda t: L.call("ecb200_knn_tc_xyz", P(rows[0]), P(rows[1]), P(xxs), P(cmax), B, N, k, P(idx), t, st)
else:
    launch = lambda t: (L.call("ecb200_debug_tc_timeline", P(hi), P(lo), P(xx), B, C, N, k, P(idx), P(ws), t, st) if t is not None
                        else L.call("ecb200_knn_tc", P(hi), P(lo), P(xx), B, C, N, k, 1, P(idx), P(ws), nb, st))
print("kernel kind:", KIND)
for _ in range(3):
    launch(P(tl))
torch.cuda.synchronize()
allc = tl.cpu()[6 * 256:].view(ntile, 3)
t = tl.cpu()[:6 * 256].view(6, 256)
t0 = int(t[5, 0])
nct = (N + 127) // 128; nkb = max(1, C // (64 if KIND != "tf32" else 32))
rel = lambda v: (int(v) - t0) if int(v) else None
print("producer stage-acquire times (first 20, last 4):", [rel(v) for v in t[0, :20]], [rel(v) for v in t[0, 2 * nct * nkb - 4:2 * nct * nkb]])
print("mma per tile (start, commit):")
for i in range(2 * nct):
    print("   tile", i, rel(t[1, 2 * i]), rel(t[1, 2 * i + 1]))
for g in (0, 1):
    print(f"epilogue group {g}: per use (begin, hx-barrier done, t_full acquired, done)")
    for u in range(nct):
        print("   use", u, [rel(v) for v in t[2 + g, 4 * u:4 * u + 4]])
    print("   pass A end / pass B end / union copied / lists complete / ranked:", [rel(v) for v in t[4, 8 * g:8 * g + 5]])
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

# warm, back-to-back launches timed with CUDA events (sustained clocks), and the SM clock implied by
# CTA (0,0): cycles between its first and last stamp vs its wall-clock duration
for _ in range(5):
    launch(None)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(50):
    launch(None)
e.record()
torch.cuda.synchronize()
us = s.elapsed_time(e) * 1e3 / 50
print(f"50 back-to-back launches: {us:.1f} us each = {2.0 * B * N * N * C / us / 1e6:.1f} TFLOP/s "
      f"({2.0 * B * N * N * C / us / 1e6 / 268.2:.3f} of the 3xTF32 roofline 268.2, {2.0 * B * N * N * C / us / 1e6 / 536.4:.3f} of the 3xFP16 roofline 536.4)")
import numpy as np
d = dur.numpy(); bg = beg.numpy(); sm = allc[:, 2].numpy()
nrt = (N + 127) // 128
print("CTA duration percentiles (us) 10/50/90/99/max:", [round(float(np.percentile(d, p)), 1) for p in (10, 50, 90, 99, 100)])
print("  by row tile (blockIdx.x):", [round(float(d[i::nrt].mean()), 1) for i in range(nrt)])
order = np.argsort(-d)[:12]
print("  slowest CTAs (tile, cloud, sm, start, dur):", [(int(i % nrt), int(i // nrt), int(sm[i]), round(float(bg[i]), 1), round(float(d[i]), 1)) for i in order])
per_sm = {}
for i in range(len(d)):
    per_sm.setdefault(int(sm[i]), []).append((float(bg[i]), float(d[i])))
busy = sorted(((sum(x[1] for x in v), len(v), s) for s, v in per_sm.items()), reverse=True)
print("  busiest SMs (sum of CTA durations, #CTAs, sm):", [(round(b, 1), n, s) for b, n, s in busy[:8]], " least:", [(round(b, 1), n, s) for b, n, s in busy[-4:]])
