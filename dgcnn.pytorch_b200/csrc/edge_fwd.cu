// EdgeConv forward epilogue: neighbour gather + BatchNorm statistics + max/min over k,
// BatchNorm finalize, affine + LeakyReLU apply, and the materialised graph feature.
//
// Replaces, without ever building the [B,2C,N,k] / [B,Co,N,k] edge tensors of the
// reference (models/dgcnn.py:31-42, :54-73, :86):
//   e_ij = U[idx[i,j]] + V[i]           (the 1x1 Conv2d output of edge (i,j), split form)
//   BatchNorm2d statistics over all B*N*k edges
//   max_j LeakyReLU(a*e_ij + b) = LeakyReLU(a*sel_i + b), sel_i = max_j e_ij if a >= 0
//                                                              min_j e_ij otherwise
// and sign(a) = sign(gamma) is known before the statistics exist, so ONE pass over the
// gathered rows yields sel, the arg slot, and the sum / sum-of-squares of e.
#include <math_constants.h>

#include "common.cuh"

namespace {

constexpr int GATHER_THREADS = 256;

// A group of LPP lanes owns one (point, 4*LPP-channel group) work item at a time; each
// lane holds 4 consecutive channels, so a neighbour row is fetched as LPP coalesced
// 128-bit loads.  A lane's channels never change, which lets it keep fp64 partial
// statistics in registers for the whole kernel.
template <int LPP>
__global__ void __launch_bounds__(GATHER_THREADS, 4)  // 4 CTAs/SM: measured best (3 -> 302 us, 4 -> 234 us, 5 -> 260 us per step)
edge_gather_kernel(const float* __restrict__ Y, const int32_t* __restrict__ idx,
                   const float* __restrict__ gamma, int N, int k, int Co, long long M,
                   float* __restrict__ sel, uint8_t* __restrict__ arg, float* __restrict__ esum,
                   double* __restrict__ stats) {
  extern __shared__ double sh_stats[];  // [2*Co] when stats != nullptr
  constexpr int SPW = 32 / LPP;          // work items a warp handles side by side
  const int lane = threadIdx.x & 31;
  const int sl = lane % LPP;
  const int G = Co / (4 * LPP);          // channel groups per point
  const long long S = (long long)gridDim.x * (GATHER_THREADS / 32) * SPW;
  const long long s = ((long long)blockIdx.x * (GATHER_THREADS / 32) + threadIdx.x / 32) * SPW +
                      lane / LPP;
  const long long PS = S / G;            // point streams
  const int g = (int)(s % G);
  const long long ps = s / G;
  const int c = g * 4 * LPP + sl * 4;    // this lane's first channel
  const int Co2 = 2 * Co;

  float sgn[4];
  {
    const float4 gm = *reinterpret_cast<const float4*>(gamma + c);
    sgn[0] = gm.x < 0.f ? -1.f : 1.f;
    sgn[1] = gm.y < 0.f ? -1.f : 1.f;
    sgn[2] = gm.z < 0.f ? -1.f : 1.f;
    sgn[3] = gm.w < 0.f ? -1.f : 1.f;
  }
  double dsum[4] = {0, 0, 0, 0}, dsq[4] = {0, 0, 0, 0};

  if (ps < PS) {
    for (long long m = ps; m < M; m += PS) {
      const long long base = (m / N) * N;  // first point of this cloud
      const int32_t* irow = idx + m * k;
      const float4 v4 = *reinterpret_cast<const float4*>(Y + m * Co2 + Co + c);
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      float best[4] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
      float se[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
      int bj[4] = {0, 0, 0, 0};
#pragma unroll 4
      for (int j = 0; j < k; ++j) {
        const int nj = __ldg(irow + j);
        const float4 u4 = *reinterpret_cast<const float4*>(Y + (base + nj) * Co2 + c);
        const float u[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float e = u[q] + v[q];
          se[q] += e;
          sq[q] = fmaf(e, e, sq[q]);
          const float t = sgn[q] * u[q];
          if (t > best[q]) { best[q] = t; bj[q] = j; }
        }
      }
      float4 o;
      o.x = sgn[0] * best[0] + v[0];
      o.y = sgn[1] * best[1] + v[1];
      o.z = sgn[2] * best[2] + v[2];
      o.w = sgn[3] * best[3] + v[3];
      *reinterpret_cast<float4*>(sel + m * Co + c) = o;
      *reinterpret_cast<uchar4*>(arg + m * Co + c) =
          make_uchar4((unsigned char)bj[0], (unsigned char)bj[1], (unsigned char)bj[2],
                      (unsigned char)bj[3]);
      if (esum) *reinterpret_cast<float4*>(esum + m * Co + c) = make_float4(se[0], se[1], se[2], se[3]);
#pragma unroll
      for (int q = 0; q < 4; ++q) { dsum[q] += (double)se[q]; dsq[q] += (double)sq[q]; }
    }
  }

  if (stats) {  // uniform across the grid
    for (int i = threadIdx.x; i < Co2; i += GATHER_THREADS) sh_stats[i] = 0.0;
    __syncthreads();
    if (ps < PS) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        atomicAdd(&sh_stats[c + q], dsum[q]);
        atomicAdd(&sh_stats[Co + c + q], dsq[q]);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Co2; i += GATHER_THREADS) atomicAdd(&stats[i], sh_stats[i]);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&stats[Co2], (double)M * (double)k);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ stats,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   const float* __restrict__ running_mean,
                                   const float* __restrict__ running_var, int training, float eps,
                                   int Co, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ a, float* __restrict__ b) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= Co) return;
  double mu, var;
  if (training) {
    const double count = stats[2 * Co];
    mu = stats[o] / count;
    var = stats[Co + o] / count - mu * mu;
    if (var < 0.0) var = 0.0;
  } else {
    mu = running_mean[o];
    var = running_var[o];
  }
  const double r = 1.0 / sqrt(var + (double)eps);
  const double aa = (double)gamma[o] * r;
  mean[o] = (float)mu;
  invstd[o] = (float)r;
  a[o] = (float)aa;
  b[o] = (float)((double)beta[o] - aa * mu);
}

// single block: every thread reads the old counter before thread 0 bumps it
__global__ void bn_update_running_kernel(const double* __restrict__ stats, int Co, float momentum,
                                         float* running_mean, float* running_var,
                                         int64_t* num_batches_tracked) {
  double factor = momentum;
  if (momentum < 0.f) factor = num_batches_tracked ? 1.0 / (double)(*num_batches_tracked + 1) : 1.0;
  const double count = stats[2 * Co];
  for (int o = threadIdx.x; o < Co; o += blockDim.x) {
    const double mu = stats[o] / count;
    double var = stats[Co + o] / count - mu * mu;
    if (var < 0.0) var = 0.0;
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    if (running_mean) running_mean[o] = (float)((1.0 - factor) * running_mean[o] + factor * mu);
    if (running_var) running_var[o] = (float)((1.0 - factor) * running_var[o] + factor * unbiased);
  }
  __syncthreads();
  if (threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
}

// [M,Co] -> [B,Co,N] with the BatchNorm affine and LeakyReLU fused into the transpose; the same
// values optionally also go out point-major (out_pm[m*ld_pm + o]) for consumers that want rows
// of channels (the 512-channel concat in front of conv5, models/dgcnn.py:100)
__global__ void __launch_bounds__(256)
edge_apply_kernel(const float* __restrict__ sel, const float* __restrict__ a,
                  const float* __restrict__ b, float slope, int N, int Co,
                  float* __restrict__ out, float* __restrict__ out_pm, long long ld_pm,
                  unsigned* __restrict__ amax) {
  __shared__ float tile[32][33];
  unsigned mx = 0u;   // max |y| of this thread as a bit pattern (non-negative floats order like unsigned)
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n0 = blockIdx.x * 32, o0 = blockIdx.y * 32, bb = blockIdx.z;
  const int o = o0 + tx;
  const float ao = o < Co ? a[o] : 0.f, bo = o < Co ? b[o] : 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int p = ty + 8 * r, n = n0 + p;
    float y = 0.f;
    if (n < N && o < Co) {
      const size_t m = (size_t)bb * N + n;
      y = ecb200::leaky(fmaf(ao, sel[m * Co + o], bo), slope);
      if (out_pm) out_pm[m * ld_pm + o] = y;
      mx = max(mx, __float_as_uint(fabsf(y)));
    }
    tile[p][tx] = y;
  }
  if (amax) {
    // the next layer's operand scale (ecb200_split_f16): one monotone atomic per block into one of
    // ECB200_AMAX_SLOTS slots, skipped once the slot is already as large
    mx = ecb200::block_max_u32(mx);
    unsigned* slot = amax + ((blockIdx.x + 3 * blockIdx.y + 5 * blockIdx.z) & (ECB200_AMAX_SLOTS - 1));
    if (tx == 0 && ty == 0 && mx > *reinterpret_cast<volatile unsigned*>(slot)) atomicMax(slot, mx);
  }
  if (!out) return;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int oc = o0 + ty + 8 * r, n = n0 + tx;
    if (oc < Co && n < N) out[((size_t)bb * Co + oc) * N + n] = tile[tx][ty + 8 * r];
  }
}

// get_graph_feature(), models/dgcnn.py:15-44: one thread per edge (n, j), looping over
// channels, so the [.., N, k]-innermost layouts are written fully coalesced.
__global__ void graph_feature_kernel(const float* __restrict__ x, const int32_t* __restrict__ idx,
                                     int C, int N, int k, int mode, float* __restrict__ out) {
  const int bb = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // n*k + j
  const long long NK = (long long)N * k;
  if (e >= NK) return;
  const int n = (int)(e / k);
  const int nj = idx[(size_t)bb * NK + e];
  const float* xb = x + (size_t)bb * C * N;
  if (mode == ECB200_GF_KNN_ONLY) {
    float* o = out + ((size_t)bb * NK + e) * C;
    if ((C & 3) == 0) {   // a thread's C outputs are contiguous: four gathered channels per 16-byte store
#pragma unroll 4
      for (int c = 0; c < C; c += 4)
        *reinterpret_cast<float4*>(o + c) = make_float4(xb[(size_t)c * N + nj], xb[(size_t)(c + 1) * N + nj],
                                                         xb[(size_t)(c + 2) * N + nj], xb[(size_t)(c + 3) * N + nj]);
    } else {
      for (int c = 0; c < C; ++c) o[c] = xb[(size_t)c * N + nj];
    }
    return;
  }
  if (mode == ECB200_GF_DISP_ONLY) {
    float* o = out + (size_t)bb * C * NK + e;
#pragma unroll 8   // independent gathers in flight: the loop is latency-bound otherwise
    for (int c = 0; c < C; ++c) o[(size_t)c * NK] = xb[(size_t)c * N + nj] - xb[(size_t)c * N + n];
    return;
  }
  float* o = out + (size_t)bb * 2 * C * NK + e;
  const bool centered = mode == ECB200_GF_CONCAT_CENTERED;
  // (left to the compiler's own unrolling: forcing 8 halved this loop's throughput, measured)
  for (int c = 0; c < C; ++c) {
    const float xi = xb[(size_t)c * N + n];
    const float xj = xb[(size_t)c * N + nj];
    o[(size_t)c * NK] = centered ? xj - xi : xj;
    o[(size_t)(C + c) * NK] = xi;
  }
}

__global__ void graph_feature_bwd_kernel(const float* __restrict__ gout,
                                         const int32_t* __restrict__ idx, int C, int N, int k,
                                         int mode, float* __restrict__ dx) {
  const int bb = blockIdx.y;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long NK = (long long)N * k;
  if (e >= NK) return;
  const int n = (int)(e / k);
  const int nj = idx[(size_t)bb * NK + e];
  float* dxb = dx + (size_t)bb * C * N;
  if (mode == ECB200_GF_KNN_ONLY) {
    const float* g = gout + ((size_t)bb * NK + e) * C;
    for (int c = 0; c < C; ++c) atomicAdd(dxb + (size_t)c * N + nj, g[c]);
    return;
  }
  if (mode == ECB200_GF_DISP_ONLY) {
    const float* g = gout + (size_t)bb * C * NK + e;
    for (int c = 0; c < C; ++c) {
      const float gv = g[(size_t)c * NK];
      atomicAdd(dxb + (size_t)c * N + nj, gv);
      atomicAdd(dxb + (size_t)c * N + n, -gv);
    }
    return;
  }
  const float* g = gout + (size_t)bb * 2 * C * NK + e;
  const bool centered = mode == ECB200_GF_CONCAT_CENTERED;
  for (int c = 0; c < C; ++c) {
    const float g1 = g[(size_t)c * NK];
    const float g2 = g[(size_t)(C + c) * NK];
    atomicAdd(dxb + (size_t)c * N + nj, g1);
    atomicAdd(dxb + (size_t)c * N + n, centered ? g2 - g1 : g2);
  }
}

template <int LPP>
int launch_gather(const float* Y, const int32_t* idx, const float* gamma, int N, int k, int Co,
                  long long M, float* sel, uint8_t* arg, float* esum, double* stats,
                  cudaStream_t st) {
  const int G = Co / (4 * LPP);
  // enough CTAs for ~8 per SM, rounded to a multiple of the SM count and of G streams
  long long items = M * G;
  long long per_cta = (long long)(GATHER_THREADS / 32) * (32 / LPP);
  long long want = ecb200::ceil_div64(items, per_cta);
  long long cap = 8LL * ecb200::kNumSMs;
  long long ctas = want < cap ? want : cap;
  if (ctas * per_cta < G) ctas = ecb200::ceil_div64(G, per_cta);
  const size_t smem = stats ? sizeof(double) * 2 * Co : 0;
  edge_gather_kernel<LPP><<<(unsigned)ctas, GATHER_THREADS, smem, st>>>(Y, idx, gamma, N, k, Co, M, sel,
                                                                      arg, esum, stats);
  ECB_LAUNCH_CHECK("edge_gather_kernel");
  return ECB200_OK;
}

}  // namespace

extern "C" int ecb200_edge_gather(const float* Y, const int32_t* idx, const float* gamma, int B,
                                  int N, int k, int Co, float* sel, uint8_t* arg, float* esum,
                                  double* stats, void* stream) {
  ECB_REQUIRE(Y && idx && gamma && sel && arg, "ecb200_edge_gather: null pointer");
  ECB_REQUIRE(B >= 1 && N >= 1, "ecb200_edge_gather: bad shape B=%d N=%d", B, N);
  ECB_REQUIRE(k >= 1 && k <= 255, "ecb200_edge_gather: k=%d out of range (arg slots are uint8)", k);
  ECB_REQUIRE(Co >= 4 && Co % 4 == 0 && Co <= 2048,
              "ecb200_edge_gather: Co=%d must be a multiple of 4 in [4, 2048]", Co);
  const long long M = (long long)B * N;
  cudaStream_t st = (cudaStream_t)stream;
  if (Co % 128 == 0) return launch_gather<32>(Y, idx, gamma, N, k, Co, M, sel, arg, esum, stats, st);
  if (Co % 64 == 0) return launch_gather<16>(Y, idx, gamma, N, k, Co, M, sel, arg, esum, stats, st);
  if (Co % 32 == 0) return launch_gather<8>(Y, idx, gamma, N, k, Co, M, sel, arg, esum, stats, st);
  if (Co % 16 == 0) return launch_gather<4>(Y, idx, gamma, N, k, Co, M, sel, arg, esum, stats, st);
  if (Co % 8 == 0) return launch_gather<2>(Y, idx, gamma, N, k, Co, M, sel, arg, esum, stats, st);
  return launch_gather<1>(Y, idx, gamma, N, k, Co, M, sel, arg, esum, stats, st);
}

extern "C" int ecb200_bn_finalize(const double* stats, const float* gamma, const float* beta,
                                  const float* running_mean, const float* running_var, int training,
                                  float eps, int Co, float* mean, float* invstd, float* a, float* b,
                                  void* stream) {
  ECB_REQUIRE(gamma && beta && mean && invstd && a && b, "ecb200_bn_finalize: null pointer");
  ECB_REQUIRE(Co >= 1, "ecb200_bn_finalize: Co=%d", Co);
  if (training) {
    ECB_REQUIRE(stats != nullptr, "ecb200_bn_finalize: training mode needs the statistics buffer");
  } else {
    ECB_REQUIRE(running_mean && running_var, "ecb200_bn_finalize: eval mode needs running statistics");
  }
  bn_finalize_kernel<<<ecb200::ceil_div(Co, 128), 128, 0, (cudaStream_t)stream>>>(
      stats, gamma, beta, running_mean, running_var, training, eps, Co, mean, invstd, a, b);
  ECB_LAUNCH_CHECK("bn_finalize_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_bn_update_running(const double* stats, int Co, float momentum,
                                        float* running_mean, float* running_var,
                                        int64_t* num_batches_tracked, void* stream) {
  ECB_REQUIRE(stats != nullptr && Co >= 1, "ecb200_bn_update_running: bad arguments");
  bn_update_running_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(stats, Co, momentum, running_mean,
                                                               running_var, num_batches_tracked);
  ECB_LAUNCH_CHECK("bn_update_running_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_edge_apply_amax(const float* sel, const float* a, const float* b, float slope,
                                      int B, int N, int Co, float* out, float* out_pm, long long ld_pm,
                                      float* amax, void* stream);
extern "C" int ecb200_edge_apply(const float* sel, const float* a, const float* b, float slope,
                                 int B, int N, int Co, float* out, float* out_pm, long long ld_pm,
                                 void* stream) {
  return ecb200_edge_apply_amax(sel, a, b, slope, B, N, Co, out, out_pm, ld_pm, nullptr, stream);
}

extern "C" int ecb200_edge_apply_amax(const float* sel, const float* a, const float* b, float slope,
                                      int B, int N, int Co, float* out, float* out_pm, long long ld_pm,
                                      float* amax, void* stream) {
  ECB_REQUIRE(sel && a && b && (out || out_pm), "ecb200_edge_apply: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && Co >= 1, "ecb200_edge_apply: bad shape");
  ECB_REQUIRE(!out_pm || ld_pm >= Co, "ecb200_edge_apply: ld_pm=%lld smaller than Co=%d", ld_pm, Co);
  dim3 grid(ecb200::ceil_div(N, 32), ecb200::ceil_div(Co, 32), B);
  if (amax) ECB_CUDA(cudaMemsetAsync(amax, 0, ECB200_AMAX_SLOTS * sizeof(float), (cudaStream_t)stream));
  edge_apply_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(sel, a, b, slope, N, Co, out, out_pm,
                                                                    ld_pm, reinterpret_cast<unsigned*>(amax));
  ECB_LAUNCH_CHECK("edge_apply_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_graph_feature(const float* x, const int32_t* idx, int B, int C, int N, int k,
                                    int mode, float* out, void* stream) {
  ECB_REQUIRE(x && idx && out, "ecb200_graph_feature: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && N >= 1 && k >= 1, "ecb200_graph_feature: bad shape");
  ECB_REQUIRE(mode >= 0 && mode <= 3, "ecb200_graph_feature: unknown mode %d", mode);
  dim3 grid((unsigned)ecb200::ceil_div64((long long)N * k, 256), B);
  graph_feature_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, idx, C, N, k, mode, out);
  ECB_LAUNCH_CHECK("graph_feature_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_graph_feature_bwd(const float* gout, const int32_t* idx, int B, int C, int N,
                                        int k, int mode, float* dx, void* stream) {
  ECB_REQUIRE(gout && idx && dx, "ecb200_graph_feature_bwd: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && N >= 1 && k >= 1, "ecb200_graph_feature_bwd: bad shape");
  ECB_REQUIRE(mode >= 0 && mode <= 3, "ecb200_graph_feature_bwd: unknown mode %d", mode);
  ECB_CUDA(cudaMemsetAsync(dx, 0, sizeof(float) * (size_t)B * C * N, (cudaStream_t)stream));
  dim3 grid((unsigned)ecb200::ceil_div64((long long)N * k, 256), B);
  graph_feature_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, idx, C, N, k, mode, dx);
  ECB_LAUNCH_CHECK("graph_feature_bwd_kernel");
  return ECB200_OK;
}
