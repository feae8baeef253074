// conv5's BatchNorm2d + LeakyReLU (models/dgcnn.py:75-78, :102) fused with the global max / average
// pooling over the points that the classification head applies right after it (upstream
// DGCNN_cls: adaptive_max_pool1d | adaptive_avg_pool1d of the [B, emb, N] embedding).
//
// z [M, E] is conv5's raw output, point-major (what cuDNN produces for a channels-last input).
// The activated [B, emb, N] tensor (128 MiB at B=32, N=1024, emb=1024) is never written:
//   colstats : per-channel sum z, sum z^2 in fp64 (training-mode statistics)      1 read of z
//   pool     : max_n / mean_n of LeakyReLU(a z + b).  LeakyReLU(a . + b) is monotone, so the
//              max needs only max_n z (a >= 0) or min_n z (a < 0) and its position; the mean
//              is accumulated element-wise                                        1 read of z
//   backward : dy[m,o] = gavg[b,o]/N + [n == arg[b,o]] gmax[b,o], through LeakyReLU' and the
//              BatchNorm backward (sum dact, sum dact*xhat, then dz)              2 reads, 1 write
// All kernels are HBM-bound streaming passes with 128-bit accesses.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int PT = 256;  // threads: 32 column-quads x 8 row lanes

// grid (E/128, row slabs): each block reduces `rows` rows of 128 columns
__global__ void __launch_bounds__(PT)
colstats_kernel(const float* __restrict__ z, long long M, int E, int rows, double* __restrict__ stats) {
  __shared__ float red[2][8][128];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + tx * 4;
  const long long r0 = (long long)blockIdx.y * rows;
  const long long r1 = r0 + rows < M ? r0 + rows : M;
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (c < E)
    for (long long r = r0 + ty; r < r1; r += 8) {
      const float4 v = ecb200::ld_stream4(z + r * E + c);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
      q[0] = fmaf(v.x, v.x, q[0]); q[1] = fmaf(v.y, v.y, q[1]);
      q[2] = fmaf(v.z, v.z, q[2]); q[3] = fmaf(v.w, v.w, q[3]);
    }
#pragma unroll
  for (int u = 0; u < 4; ++u) { red[0][ty][tx * 4 + u] = s[u]; red[1][ty][tx * 4 + u] = q[u]; }
  __syncthreads();
  const int which = threadIdx.x >> 7, col = threadIdx.x & 127;  // 256 threads = 2 x 128 columns
  if (blockIdx.x * 128 + col < E) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc += (double)red[which][r][col];
    atomicAdd(&stats[which * E + blockIdx.x * 128 + col], acc);
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) atomicAdd(&stats[2 * E], (double)M);
}

// grid (E/128, B): one cloud, 128 channels per block
__global__ void __launch_bounds__(PT)
pool_kernel(const float* __restrict__ z, const float* __restrict__ a, const float* __restrict__ b,
            float slope, int N, int E, float* __restrict__ pooled, int32_t* __restrict__ arg) {
  __shared__ float sbest[8][128];
  __shared__ int sarg[8][128];
  __shared__ float ssum[8][128];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + tx * 4;
  const int bb = blockIdx.y;
  float av[4] = {0, 0, 0, 0}, bv[4] = {0, 0, 0, 0}, sg[4] = {1, 1, 1, 1};
  if (c < E) {
    const float4 a4 = *reinterpret_cast<const float4*>(a + c), b4 = *reinterpret_cast<const float4*>(b + c);
    av[0] = a4.x; av[1] = a4.y; av[2] = a4.z; av[3] = a4.w;
    bv[0] = b4.x; bv[1] = b4.y; bv[2] = b4.z; bv[3] = b4.w;
#pragma unroll
    for (int u = 0; u < 4; ++u) sg[u] = av[u] < 0.f ? -1.f : 1.f;
  }
  float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
  int bn[4] = {0, 0, 0, 0};
  float sum[4] = {0, 0, 0, 0};
  if (c < E) {
    const float* zb = z + (size_t)bb * N * E + c;
    for (int n = ty; n < N; n += 8) {
      const float4 v4 = ecb200::ld_stream4(zb + (size_t)n * E);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float t = sg[u] * v[u];
        if (t > best[u]) { best[u] = t; bn[u] = n; }   // first occurrence wins within a thread
        sum[u] += ecb200::leaky(fmaf(av[u], v[u], bv[u]), slope);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    sbest[ty][tx * 4 + u] = best[u];
    sarg[ty][tx * 4 + u] = bn[u];
    ssum[ty][tx * 4 + u] = sum[u];
  }
  __syncthreads();
  if (threadIdx.x < 128 && blockIdx.x * 128 + threadIdx.x < E) {
    const int col = threadIdx.x, o = blockIdx.x * 128 + col;
    float bst = sbest[0][col], tot = ssum[0][col];
    int bi = sarg[0][col];
#pragma unroll
    for (int r = 1; r < 8; ++r) {
      const float t = sbest[r][col];
      const int i = sarg[r][col];
      if (t > bst || (t == bst && i < bi)) { bst = t; bi = i; }   // smallest n among equal maxima
      tot += ssum[r][col];
    }
    const float ao = a[o], bo = b[o];
    const float sel = (ao < 0.f ? -1.f : 1.f) * bst;
    pooled[(size_t)bb * 2 * E + o] = ecb200::leaky(fmaf(ao, sel, bo), slope);
    pooled[(size_t)bb * 2 * E + E + o] = tot / (float)N;
    arg[(size_t)bb * E + o] = bi;
  }
}

// dact[m,o] = (gavg[b,o]/N + [n == arg[b,o]] gmax[b,o]) * LeakyReLU'(a z + b)
__device__ __forceinline__ float dact_of(float zv, float ao, float bo, float slope, float gav, float gmx,
                                         bool is_arg) {
  const float dy = gav + (is_arg ? gmx : 0.f);
  return fmaf(ao, zv, bo) > 0.f ? dy : dy * slope;
}

// backward pass 1: bstats[0..E) += sum dact, bstats[E..2E) += sum dact * (z - mean) * invstd
// grid (E/128, B, N slabs)
__global__ void __launch_bounds__(PT)
pool_bwd_stats_kernel(const float* __restrict__ z, const float* __restrict__ gpool,
                      const int32_t* __restrict__ arg, const float* __restrict__ a,
                      const float* __restrict__ b, const float* __restrict__ mean,
                      const float* __restrict__ invstd, float slope, int N, int E, int rows,
                      double* __restrict__ bstats) {
  __shared__ float red[2][8][128];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + tx * 4;
  const int bb = blockIdx.y;
  const int n0 = blockIdx.z * rows, n1 = min(N, n0 + rows);
  float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
  if (c < E) {
    const float4 a4 = *reinterpret_cast<const float4*>(a + c), b4 = *reinterpret_cast<const float4*>(b + c);
    const float4 m4 = *reinterpret_cast<const float4*>(mean + c), r4 = *reinterpret_cast<const float4*>(invstd + c);
    const float4 gm4 = *reinterpret_cast<const float4*>(gpool + (size_t)bb * 2 * E + c);
    const float4 ga4 = *reinterpret_cast<const float4*>(gpool + (size_t)bb * 2 * E + E + c);
    const int4 ar4 = *reinterpret_cast<const int4*>(arg + (size_t)bb * E + c);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    const float mu[4] = {m4.x, m4.y, m4.z, m4.w}, rs[4] = {r4.x, r4.y, r4.z, r4.w};
    const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w};
    const float inv_n = 1.f / (float)N;
    const float ga[4] = {ga4.x * inv_n, ga4.y * inv_n, ga4.z * inv_n, ga4.w * inv_n};
    const int ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
    const float* zb = z + (size_t)bb * N * E + c;
    for (int n = n0 + ty; n < n1; n += 8) {
      const float4 v4 = ecb200::ld_stream4(zb + (size_t)n * E);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float d = dact_of(v[u], av[u], bv[u], slope, ga[u], gm[u], n == ar[u]);
        s[u] += d;
        q[u] = fmaf(d, (v[u] - mu[u]) * rs[u], q[u]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) { red[0][ty][tx * 4 + u] = s[u]; red[1][ty][tx * 4 + u] = q[u]; }
  __syncthreads();
  const int which = threadIdx.x >> 7, col = threadIdx.x & 127;
  if (blockIdx.x * 128 + col < E) {
    double acc = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) acc += (double)red[which][r][col];
    atomicAdd(&bstats[which * E + blockIdx.x * 128 + col], acc);
  }
}

// backward pass 2: dz = a*dact - c1 - c2*(z - mean)   (c1 = a*sum dact/count,
// c2 = a*sum(dact*xhat)*invstd/count from ecb200_bwd_finalize; both 0 in eval mode)
__global__ void __launch_bounds__(PT)
pool_bwd_dz_kernel(const float* __restrict__ z, const float* __restrict__ gpool,
                   const int32_t* __restrict__ arg, const float* __restrict__ a,
                   const float* __restrict__ b, const float* __restrict__ mean,
                   const float* __restrict__ c1, const float* __restrict__ c2, float slope, int N, int E,
                   int rows, float* __restrict__ dz) {
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int c = blockIdx.x * 128 + tx * 4;
  const int bb = blockIdx.y;
  const int n0 = blockIdx.z * rows, n1 = min(N, n0 + rows);
  if (c >= E) return;
  const float4 a4 = *reinterpret_cast<const float4*>(a + c), b4 = *reinterpret_cast<const float4*>(b + c);
  const float4 m4 = *reinterpret_cast<const float4*>(mean + c);
  const float4 k14 = *reinterpret_cast<const float4*>(c1 + c), k24 = *reinterpret_cast<const float4*>(c2 + c);
  const float4 gm4 = *reinterpret_cast<const float4*>(gpool + (size_t)bb * 2 * E + c);
  const float4 ga4 = *reinterpret_cast<const float4*>(gpool + (size_t)bb * 2 * E + E + c);
  const int4 ar4 = *reinterpret_cast<const int4*>(arg + (size_t)bb * E + c);
  const float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
  const float mu[4] = {m4.x, m4.y, m4.z, m4.w};
  const float k1[4] = {k14.x, k14.y, k14.z, k14.w}, k2[4] = {k24.x, k24.y, k24.z, k24.w};
  const float gm[4] = {gm4.x, gm4.y, gm4.z, gm4.w};
  const float inv_n = 1.f / (float)N;
  const float ga[4] = {ga4.x * inv_n, ga4.y * inv_n, ga4.z * inv_n, ga4.w * inv_n};
  const int ar[4] = {ar4.x, ar4.y, ar4.z, ar4.w};
  const size_t off = (size_t)bb * N * E + c;
  for (int n = n0 + ty; n < n1; n += 8) {
    const float4 v4 = ecb200::ld_stream4(z + off + (size_t)n * E);
    const float v[4] = {v4.x, v4.y, v4.z, v4.w};
    float o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float d = dact_of(v[u], av[u], bv[u], slope, ga[u], gm[u], n == ar[u]);
      o[u] = fmaf(av[u], d, -k1[u]) - k2[u] * (v[u] - mu[u]);
    }
    *reinterpret_cast<float4*>(dz + off + (size_t)n * E) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

}  // namespace

extern "C" int ecb200_colstats(const float* z, long long M, int E, double* stats, void* stream) {
  ECB_REQUIRE(z && stats, "ecb200_colstats: null pointer");
  ECB_REQUIRE(M >= 1 && E >= 4 && E % 4 == 0, "ecb200_colstats: bad shape M=%lld E=%d", M, E);
  const int rows = 256;
  dim3 grid(ecb200::ceil_div(E, 128), (unsigned)ecb200::ceil_div64(M, rows));
  colstats_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(z, M, E, rows, stats);
  ECB_LAUNCH_CHECK("colstats_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_embed_pool(const float* z, const float* a, const float* b, float slope, int B, int N,
                                 int E, float* pooled, int32_t* arg, void* stream) {
  ECB_REQUIRE(z && a && b && pooled && arg, "ecb200_embed_pool: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && E >= 4 && E % 4 == 0, "ecb200_embed_pool: bad shape");
  dim3 grid(ecb200::ceil_div(E, 128), B);
  pool_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(z, a, b, slope, N, E, pooled, arg);
  ECB_LAUNCH_CHECK("pool_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_embed_pool_bwd_stats(const float* z, const float* gpool, const int32_t* arg,
                                           const float* a, const float* b, const float* mean,
                                           const float* invstd, float slope, int B, int N, int E,
                                           double* bstats, void* stream) {
  ECB_REQUIRE(z && gpool && arg && a && b && mean && invstd && bstats, "ecb200_embed_pool_bwd_stats: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && E >= 4 && E % 4 == 0, "ecb200_embed_pool_bwd_stats: bad shape");
  const int rows = 128;
  dim3 grid(ecb200::ceil_div(E, 128), B, ecb200::ceil_div(N, rows));
  pool_bwd_stats_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(z, gpool, arg, a, b, mean, invstd, slope, N, E,
                                                              rows, bstats);
  ECB_LAUNCH_CHECK("pool_bwd_stats_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_embed_pool_bwd_dz(const float* z, const float* gpool, const int32_t* arg,
                                        const float* a, const float* b, const float* mean, const float* c1,
                                        const float* c2, float slope, int B, int N, int E, float* dz,
                                        void* stream) {
  ECB_REQUIRE(z && gpool && arg && a && b && mean && c1 && c2 && dz, "ecb200_embed_pool_bwd_dz: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && E >= 4 && E % 4 == 0, "ecb200_embed_pool_bwd_dz: bad shape");
  const int rows = 128;
  dim3 grid(ecb200::ceil_div(E, 128), B, ecb200::ceil_div(N, rows));
  pool_bwd_dz_kernel<<<grid, PT, 0, (cudaStream_t)stream>>>(z, gpool, arg, a, b, mean, c1, c2, slope, N, E, rows,
                                                           dz);
  ECB_LAUNCH_CHECK("pool_bwd_dz_kernel");
  return ECB200_OK;
}
