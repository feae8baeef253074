// Backward of BatchNorm over the rows of a small [M, F] activation (the classification head's bn6 / bn7
// under SyncBatchNorm, main_partseg_dist.py:189 -> dgcnn.pytorch_b200/syncbn.py): two launches around the
// statistics exchange instead of a dozen element-wise / reduction kernels of a few microseconds each.
//   stats:  total[f] = sum_m g, total[F+f] = sum_m g * xhat   (fp64, the vector that is all-reduced),
//           dbeta[f], dgamma[f] = the same sums in fp32 (the LOCAL parameter gradients, as in torch's SyncBatchNorm)
//   dx:     dx = a * (g - total[f]/n - xhat * total[F+f]/n),   xhat = (x - mean) * invstd
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(128)
rows_bn_bwd_stats_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                         const float* __restrict__ invstd, int M, int F, double* __restrict__ total,
                         float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const float mu = mean[f], r = invstd[f];
  double sg = 0.0, sgx = 0.0;
  for (int m = 0; m < M; ++m) {     // consecutive threads read consecutive columns: coalesced rows
    const float gv = g[(size_t)m * F + f];
    const float xh = (x[(size_t)m * F + f] - mu) * r;
    sg += (double)gv;
    sgx += (double)(gv * xh);
  }
  total[f] = sg;
  total[F + f] = sgx;
  dbeta[f] = (float)sg;
  dgamma[f] = (float)sgx;
}

__global__ void __launch_bounds__(256)
rows_bn_bwd_dx_kernel(const float* __restrict__ g, const float* __restrict__ x, const float* __restrict__ mean,
                      const float* __restrict__ invstd, const float* __restrict__ a,
                      const double* __restrict__ total, const double* __restrict__ count, long long n, int F,
                      float* __restrict__ dx) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int f = (int)(i % F);
  const double cnt = *count;
  const float c1 = (float)(total[f] / cnt), c2 = (float)(total[F + f] / cnt);
  const float xh = (x[i] - mean[f]) * invstd[f];
  dx[i] = a[f] * (g[i] - c1 - xh * c2);
}

}  // namespace

extern "C" int ecb200_rows_bn_bwd_stats(const float* g, const float* x, const float* mean, const float* invstd,
                                        int M, int F, double* total, float* dgamma, float* dbeta, void* stream) {
  ECB_REQUIRE(g && x && mean && invstd && total && dgamma && dbeta, "ecb200_rows_bn_bwd_stats: null pointer");
  ECB_REQUIRE(M >= 1 && F >= 1, "ecb200_rows_bn_bwd_stats: bad shape");
  rows_bn_bwd_stats_kernel<<<ecb200::ceil_div(F, 128), 128, 0, (cudaStream_t)stream>>>(g, x, mean, invstd, M, F, total,
                                                                                      dgamma, dbeta);
  ECB_LAUNCH_CHECK("rows_bn_bwd_stats_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_rows_bn_bwd_dx(const float* g, const float* x, const float* mean, const float* invstd,
                                     const float* a, const double* total, const double* count, int M, int F,
                                     float* dx, void* stream) {
  ECB_REQUIRE(g && x && mean && invstd && a && total && count && dx, "ecb200_rows_bn_bwd_dx: null pointer");
  ECB_REQUIRE(M >= 1 && F >= 1, "ecb200_rows_bn_bwd_dx: bad shape");
  const long long n = (long long)M * F;
  rows_bn_bwd_dx_kernel<<<(unsigned)ecb200::ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(
      g, x, mean, invstd, a, total, count, n, F, dx);
  ECB_LAUNCH_CHECK("rows_bn_bwd_dx_kernel");
  return ECB200_OK;
}
