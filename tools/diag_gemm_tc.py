"""Tensor-core GEMMs (point GEMM, dx, dW) against fp64 torch, plus their timings.

    python tools/diag_gemm_tc.py [B,C,N,Co ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from ctypes import c_void_p
import torch
import dgcnn_pytorch_b200 as ec

L = ec._lib
dev = torch.device("cuda:0")
P = lambda t: c_void_p(t.data_ptr())
cfgs = [(32, 64, 1024, 64), (32, 64, 1024, 128), (32, 128, 1024, 256), (2, 32, 200, 16), (3, 128, 333, 48)]
if len(sys.argv) > 1:
    cfgs = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]


def timeit(fn, n=10):
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


for B, C, N, Co in cfgs:
    torch.manual_seed(B + C + N + Co)
    M, Co2 = B * N, 2 * Co
    st = c_void_p(torch.cuda.current_stream().cuda_stream)
    x = torch.randn(B, C, N, device=dev)
    dY = torch.randn(M, Co2, device=dev)
    Wcat = torch.randn(Co2, C, device=dev) / C ** 0.5
    xhi = torch.empty(M, C, device=dev); xlo = torch.empty_like(xhi); xx = torch.empty(M, device=dev)
    L.call("ecb200_split_tf32", P(x), B, C, N, P(xhi), P(xlo), P(xx), st)
    dYs = torch.empty(2, M, Co2, device=dev)
    wT = torch.empty(2, C, Co2, device=dev)
    ws = torch.empty(2, Co2, C, device=dev)
    dx = torch.empty(B, C, N, device=dev)
    dW = torch.empty(Co2, C, device=dev)
    Y = torch.empty(M, Co2, device=dev)
    f_split = lambda: L.call("ecb200_split_rows_tf32", P(dY), M * Co2, P(dYs[0]), P(dYs[1]), st)
    f_wt = lambda: L.call("ecb200_transpose_split_tf32", P(Wcat), Co2, C, P(wT[0]), P(wT[1]), st)
    f_ws = lambda: L.call("ecb200_split_rows_tf32", P(Wcat), Co2 * C, P(ws[0]), P(ws[1]), st)
    f_dx = lambda: L.call("ecb200_gemm_dx_tc", P(dYs[0]), P(dYs[1]), P(wT[0]), P(wT[1]), B, C, N, Co2, P(dx), st)
    f_dw = lambda: L.call("ecb200_gemm_dw_tc", P(dYs[0]), P(dYs[1]), P(xhi), P(xlo), M, C, Co2, P(dW), st)
    f_y = lambda: L.call("ecb200_point_gemm_tc", P(xhi), P(xlo), P(ws[0]), P(ws[1]), M, C, Co2, P(Y), st)
    f_split(); f_wt(); f_ws(); f_dx(); f_dw(); f_y()
    torch.cuda.synchronize()
    X = x.permute(0, 2, 1).reshape(M, C).double()
    ref_dx = (dY.double() @ Wcat.double()).view(B, N, C).permute(0, 2, 1)
    ref_dw = dY.double().t() @ X
    ref_y = X @ Wcat.double().t()
    rel = lambda a, b: ((a.double() - b).abs().max() / b.abs().max()).item()
    print(f"B={B} C={C} N={N} Co={Co}: rel err  Y {rel(Y, ref_y):.2e}  dx {rel(dx, ref_dx):.2e}  "
          f"dW {rel(dW, ref_dw):.2e}   us: split_dY {timeit(f_split):.1f} Y {timeit(f_y):.1f} "
          f"dx {timeit(f_dx):.1f} dW {timeit(f_dw):.1f}", flush=True)

# ---- stress: repeated dx / dW on fresh data, error pattern of any bad result
if os.environ.get("STRESS"):
    for B, C, N, Co in [(1, 128, 1024, 256), (2, 64, 1024, 64), (2, 64, 300, 128)]:
        M, Co2 = B * N, 2 * Co
        st = c_void_p(torch.cuda.current_stream().cuda_stream)
        bad = 0
        for it in range(int(os.environ["STRESS"])):
            junk = torch.randn(1 << 20, device=dev) * 1e3      # dirty the allocator's free blocks
            del junk
            dY = torch.randn(M, Co2, device=dev) * (1 + 10 * (torch.rand(M, 1, device=dev) > 0.9))
            Wcat = (torch.rand(Co2, C, device=dev) - 0.5) / C ** 0.5
            x = torch.randn(B, C, N, device=dev)
            xhi = torch.empty(M, C, device=dev); xlo = torch.empty_like(xhi); xx = torch.empty(M, device=dev)
            L.call("ecb200_split_tf32", P(x), B, C, N, P(xhi), P(xlo), P(xx), st)
            dYs = torch.empty(2, M, Co2, device=dev)
            wT = torch.empty(2, C, Co2, device=dev)
            dx = torch.empty(B, C, N, device=dev)
            dW = torch.empty(Co2, C, device=dev)
            L.call("ecb200_split_rows_tf32", P(dY), M * Co2, P(dYs[0]), P(dYs[1]), st)
            L.call("ecb200_transpose_split_tf32", P(Wcat), Co2, C, P(wT[0]), P(wT[1]), st)
            L.call("ecb200_gemm_dx_tc", P(dYs[0]), P(dYs[1]), P(wT[0]), P(wT[1]), B, C, N, Co2, P(dx), st)
            L.call("ecb200_gemm_dw_tc", P(dYs[0]), P(dYs[1]), P(xhi), P(xlo), M, C, Co2, P(dW), st)
            ref_dx = (dY.double() @ Wcat.double()).view(B, N, C).permute(0, 2, 1)
            ref_dw = dY.double().t() @ x.permute(0, 2, 1).reshape(M, C).double()
            for name, ours, ref in (("dx", dx, ref_dx), ("dW", dW, ref_dw)):
                err = (ours.double() - ref).abs()
                r = (err.max() / ref.abs().max()).item()
                if r > 2e-5:
                    bad += 1
                    w = (err > 2e-5 * ref.abs().max()).nonzero()
                    print(f"  BAD {name} it={it} shape={(B, C, N, Co)} rel={r:.2e} n_bad={len(w)} "
                          f"first={w[:4].tolist()} last={w[-2:].tolist()}", flush=True)
        print(f"stress {(B, C, N, Co)}: {bad} bad results", flush=True)
