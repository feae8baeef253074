"""conv5 forward: own tcgen05 GEMM + statistics epilogue vs the library convolution + colstats."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import torch
import torch.nn.functional as F
import dgcnn_pytorch_b200 as ec
from ctypes import c_void_p

dev = torch.device("cuda:0")
M, K, E = 32768, 512, 1024
x = torch.randn(M, K, device=dev)
w = torch.randn(E, K, 1, 1, device=dev) / K ** 0.5


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


flops = 2.0 * M * K * E
for three in (False, True):
    us = timeit(lambda: ec.ops.embed_gemm_op(x, w, three))
    print(f"own embed GEMM ({'3xTF32 incl. operand split' if three else 'TF32'}): {us:7.1f} us  {flops / us / 1e6:6.1f} TFLOP/s")
x4 = x.view(1, M, 1, K).permute(0, 3, 1, 2)
stats = torch.zeros(2 * E + 1, device=dev, dtype=torch.float64)
P = lambda t: c_void_p(t.data_ptr())
st = c_void_p(torch.cuda.current_stream().cuda_stream)


def lib():
    z = F.conv2d(x4, w)
    zz = z.permute(0, 2, 3, 1).reshape(M, E)
    ec._lib.call("ecb200_colstats", P(zz), M, E, P(stats), st)


torch.backends.cudnn.allow_tf32 = True
print(f"library conv (TF32) + colstats kernel: {timeit(lib):7.1f} us")
print(f"library conv (TF32) alone:             {timeit(lambda: F.conv2d(x4, w)):7.1f} us")
