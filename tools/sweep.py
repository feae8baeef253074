"""BASELINE config 5: isolated kNN / EdgeConv sweep against the tensor-pipe and HBM rooflines.

    python tools/sweep.py [--out profiles/r1_sweep.json]

For N in {1024..16384}, C in {3, 64, 128}, k in {20, 40} (B = 32*1024/N: constant 32 Ki points):
  knn       : ecb200 kNN (FP32-FMA kernel for C=3, tcgen05 3xTF32 kernel for C=64/128), CUDA events,
              TFLOP/s = 2*M*N*C / t  against the 3xTF32 tensor roofline (bf16 peak / 2 / 3)
  edgeconv  : fused EdgeConv block forward (training mode) and forward+backward on a fixed graph,
              GB/s = SpMM-convention bytes / t against the measured HBM peak
              (forward bytes 4M(C + k + k*Co + 3Co) + M*Co)
Inputs are synthetic (oracle generators); nothing here is a parity claim -- see tests/.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import torch  # noqa: E402

import dgcnn_pytorch_b200 as ec  # noqa: E402
import edgeconv_oracle as orc  # noqa: E402


def timeit(fn, n=10, warm=3):
    """us per call of fn, captured once into a CUDA graph and replayed (the kernels are 10-100 us:
    eager Python launches would measure the host)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / n      # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_sweep.json"))
    ap.add_argument("--points", type=int, default=32 * 1024)
    a = ap.parse_args()
    peaks = {"hbm_gbs": 6551.4, "bf16_tflops": 1609.0}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        with open(pp) as f:
            peaks.update({k: float(v) for k, v in json.load(f).items() if k in peaks})
    tf32x3 = peaks["bf16_tflops"] / 2.0 / 3.0
    dev = torch.device("cuda:0")
    rows = []
    for N in (1024, 2048, 4096, 8192, 16384):
        B = max(1, a.points // N)
        M = B * N
        for C, Co in ((3, 64), (64, 64), (128, 256)):
            x = (orc.synthetic_xyz(B, N, seed=1) if C == 3 else orc.synthetic_features(B, C, N, seed=1)).to(dev)
            for k in (20, 40):
                t_knn = timeit(lambda: ec.ops.knn_op(x, k, False), n=5 if N >= 8192 else 10)
                idx = ec.ops.knn_op(x, k, False)
                torch.manual_seed(0)
                block = torch.nn.Sequential(torch.nn.Conv2d(2 * C, Co, 1, bias=False), torch.nn.BatchNorm2d(Co),
                                            torch.nn.LeakyReLU(0.2)).to(dev).train()
                xg = x.clone().requires_grad_(True)

                def fwd():
                    with torch.no_grad():
                        return ec.edgeconv_block(x, block, k, idx=idx)[0]

                gout = torch.ones(B, Co, N, device=dev)

                def fwdbwd():
                    y = ec.edgeconv_block(xg, block, k, idx=idx)[0]
                    return torch.autograd.grad(y, [xg] + list(block.parameters()), gout)

                t_f = timeit(fwd)
                t_fb = timeit(fwdbwd)
                q_fwd = 4 * M * (C + k + k * Co + 3 * Co) + M * Co
                flops = 2.0 * M * N * C
                row = {"N": N, "B": B, "C": C, "Co": Co, "k": k,
                       "knn_us": round(t_knn, 1), "knn_tflops": round(flops / t_knn / 1e6, 2),
                       "knn_kernel": "fp32-fma" if C == 3 else "tcgen05-3xtf32",
                       "knn_frac_of_3xtf32_roofline": None if C == 3 else round(flops / t_knn / 1e6 / tf32x3, 3),
                       "knn_tpairs_s": round(M * N / t_knn / 1e6, 3),
                       "edgeconv_fwd_us": round(t_f, 1), "edgeconv_fwd_gbs": round(q_fwd / t_f / 1e3, 0),
                       "edgeconv_fwd_frac_of_hbm": round(q_fwd / t_f / 1e3 / peaks["hbm_gbs"], 3),
                       "edgeconv_fwdbwd_us": round(t_fb, 1)}
                rows.append(row)
                print(json.dumps(row), flush=True)
    out = {"peaks": peaks, "tf32x3_roofline_tflops": tf32x3,
           "note": "knn_us includes the operand split / norms kernel; edgeconv on a fixed graph "
                   "(block only: per-point GEMM + gather + BN + apply, training mode); L2 not flushed",
           "rows": rows}
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
