"""Packed-FP16 tensor-core kNN: score accuracy vs fp64, graph parity vs the oracle, kernel timing.

    python tools/check_f16.py            # everything
    python tools/check_f16.py time       # timing only
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from ctypes import c_void_p
import torch
import dgcnn_pytorch_b200 as ec
import edgeconv_oracle as orc

L = ec._lib
ops = ec.ops
dev = torch.device("cuda:0")
P = lambda t: None if t is None else c_void_p(t.data_ptr())
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def exact_scores(x):
    xd = x.double()
    return torch.einsum("bci,bcj->bij", xd, xd) - 0.5 * (xd * xd).sum(1)[:, None, :]


if what in ("all", "scores"):
    for (B, C, N, scale) in [(2, 64, 256, 1.0), (1, 128, 384, 1.0), (2, 64, 200, 1e-3), (1, 128, 130, 3e4)]:
        x = (orc.synthetic_features(B, C, N, seed=C + N) * scale).to(dev)
        ref = exact_scores(x)
        s16 = ops.debug_tc_scores_f16(x).double()
        den = float((x.double() ** 2).sum(1).max())
        e16 = float((s16 - ref).abs().max()) / den
        msg = f"scores B={B} C={C} N={N} scale={scale:g}: f16x3 err/max|x|^2 = {e16:.2e}"
        if C <= 128:
            s32 = ops.debug_tc_scores(x).double()
            msg += f"   tf32x3 = {float((s32 - ref).abs().max()) / den:.2e}"
        print(msg, flush=True)
        assert e16 < 4e-6, e16

if what in ("all", "parity"):
    cases = [(4, 64, 1024, 20, "feat"), (2, 128, 1024, 20, "feat"), (2, 64, 2048, 40, "feat"), (1, 128, 4096, 20, "feat"),
             (3, 64, 200, 12, "feat"), (1, 128, 300, 33, "feat"), (32, 3, 1024, 20, "xyz"), (4, 3, 2048, 40, "xyz"),
             (2, 3, 4096, 20, "xyz"), (3, 3, 333, 16, "xyz"), (5, 3, 64, 40, "xyz"), (2, 4, 100, 7, "feat4")]
    for B, C, N, k, kind in cases:
        if kind == "xyz":
            x = orc.synthetic_xyz(B, N, seed=B + N)
        else:
            x = orc.synthetic_features(B, C, N, seed=C + N)
        xg = x.to(dev)
        tk = ops.knn_tc_kind(C, N, k)
        idx = ops.knn_op(xg, k, True).long().cpu()
        rep = orc.knn_mismatch_report(x, idx, orc.knn_oracle(x, k), rel_eps=1e-6)
        print(f"parity B={B} C={C} N={N} k={k} kernel={tk}: {rep}", flush=True)
        assert rep["bad_rows"] == 0, rep
    # ties: lattice + duplicates through the xyz tensor-core kernel
    g = torch.Generator().manual_seed(5)
    pts = torch.randint(0, 6, (2, 3, 512), generator=g).float()
    pts[:, :, 256:] = pts[:, :, :256]
    idx = ops.knn_op(pts.to(dev), 20, True).long().cpu()
    rep = orc.knn_mismatch_report(pts, idx, orc.knn_oracle(pts, 20), rel_eps=1e-6)
    print("parity lattice+duplicates xyz:", rep, flush=True)
    assert rep["bad_rows"] == 0, rep
    same = (idx == orc.knn_oracle(pts, 20)).float().mean().item()
    print(f"  identical index tensor fraction (tie order = smaller j first): {same:.4f}")


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n


if what in ("all", "time"):
    for rows in ("128", "256"):
        os.environ["ECB200_KNN_ROWS"] = rows
        for (B, C, N, k) in [(32, 64, 1024, 20), (32, 3, 1024, 20)]:
            x = (orc.synthetic_xyz(B, N, seed=1) if C == 3 else orc.synthetic_features(B, C, N, seed=1)).to(dev)
            t = timeit(lambda: ops.knn_op(x, k, True))
            print(f"time knn() incl. operand prep, rows/CTA {rows}: B={B} C={C} N={N} k={k}: {t:.1f} us", flush=True)
    os.environ.pop("ECB200_KNN_ROWS")
    for (B, C, N, k) in [(32, 64, 1024, 20), (32, 128, 1024, 20), (32, 64, 2048, 40), (8, 128, 4096, 20)]:
        x = orc.synthetic_features(B, C, N, seed=1).to(dev)
        st = stream()
        hh, hl, nb, xxs, cmax, hi, lo, xx = ops.split_f16_op(x, True)
        idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
        ws = torch.empty(256, device=dev, dtype=torch.uint8)
        amax = torch.empty(32, device=dev)
        fl = 2.0 * B * N * N * C
        t16 = timeit(lambda: L.call("ecb200_knn_tc_f16", P(hh), P(hl), P(nb), P(xxs), P(cmax), B, C, N, k, P(idx), None, st))
        t32 = timeit(lambda: L.call("ecb200_knn_tc", P(hi), P(lo), P(xx), B, C, N, k, 1, P(idx), P(ws), 256, st))
        tam = timeit(lambda: L.call("ecb200_absmax", P(x), x.numel(), P(amax), st))
        tsp = timeit(lambda: L.call("ecb200_split_f16", P(x), B, C, N, P(amax), P(hh), P(hl), P(nb), P(xxs), P(cmax), P(hi), P(lo), P(xx), st))
        ts0 = timeit(lambda: L.call("ecb200_split_tf32", P(x), B, C, N, P(hi), P(lo), P(xx), st))
        print(f"time B={B} C={C} N={N} k={k}: knn f16 {t16:.1f} us ({fl / t16 / 1e6:.1f} TF/s)  tf32 {t32:.1f} us "
              f"({fl / t32 / 1e6:.1f} TF/s)   absmax {tam:.1f}  split_f16(+tf32) {tsp:.1f}  split_tf32 {ts0:.1f}", flush=True)
    for (B, N, k) in [(32, 1024, 20), (32, 2048, 40), (16, 4096, 20)]:
        x = orc.synthetic_xyz(B, N, seed=1).to(dev)
        st = stream()
        rows = torch.empty(2, B * N, 64, device=dev, dtype=torch.float16)
        xxs = torch.empty(B * N, device=dev)
        cmax = torch.empty(B * ((N + 31) // 32), device=dev)
        idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
        xx = torch.empty(B * N, device=dev)
        tpk = timeit(lambda: L.call("ecb200_pack_xyz_f16", P(x), B, 3, N, P(rows[0]), P(rows[1]), P(xxs), P(cmax), st))
        ttc = timeit(lambda: L.call("ecb200_knn_tc_xyz", P(rows[0]), P(rows[1]), P(xxs), P(cmax), B, N, k, P(idx), None, st))
        L.call("ecb200_sqnorms", P(x), B, 3, N, P(xx), st)
        tfm = timeit(lambda: L.call("ecb200_knn", P(x), P(xx), B, 3, N, k, 1, P(idx), st))
        print(f"time xyz B={B} N={N} k={k}: tensor-core {ttc:.1f} us (+ pack {tpk:.1f})   fma kernel {tfm:.1f} us", flush=True)
print("OK")
