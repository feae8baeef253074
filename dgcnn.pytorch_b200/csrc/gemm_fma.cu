// Per-point GEMMs of the split edge MLP, FP32 FMA, shared-memory tiled.
//   W.[x_j ; x_i] = W1.x_j + W2.x_i = U[j] + V[i]       
// so the reference's k-fold 1x1 Conv2d over [B,2C,N,k] (models/dgcnn.py:54-73)
// becomes ONE dense GEMM per layer over the B*N points, and its backward two more:
//   point_gemm : Y[M,2Co]     = X^T . Wcat^T        (X = x[B,C,N] read in place)
//   gemm_dx    : dx[B,C,N]    = dY . Wcat            (written straight back channel-major)
//   gemm_dw    : dWcat[2Co,C] = dY^T . X^T           (split over M, atomically reduced)
// One 64x64x16 tile kernel serves all three through small accessor policies.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, LD = 68;  // LD keeps rows 16-byte aligned
constexpr int GT = 256;                             // threads; each owns a 4x4 micro-tile

// x[B,C,N] viewed as the [M,C] matrix X^T (row m = b*N + n)
struct PointsRead {
  const float* x; int C, N;
  __device__ float at(long long m, int c) const {
    long long b = m / N; int n = (int)(m - b * N);
    return x[((size_t)b * C + c) * N + n];
  }
};
struct RowMajorRead {  // plain [rows, ld]
  const float* p; int ld;
  __device__ float at(long long r, int c) const { return p[(size_t)r * ld + c]; }
};
struct ColMajorRead {  // element (r, c) stored at p[c*ld + r]
  const float* p; int ld;
  __device__ float at(long long r, int c) const { return p[(size_t)c * ld + r]; }
};
struct RowMajorWrite {
  float* p; int ld;
  __device__ void put(long long r, int c, float v) const { p[(size_t)r * ld + c] = v; }
};
struct PointsWrite {  // dx[B,C,N] from the [M,C] result
  float* x; int C, N;
  __device__ void put(long long m, int c, float v) const {
    long long b = m / N; int n = (int)(m - b * N);
    x[((size_t)b * C + c) * N + n] = v;
  }
};
struct RowMajorAtomic {
  float* p; int ld;
  __device__ void put(long long r, int c, float v) const { atomicAdd(p + (size_t)r * ld + c, v); }
};

// C[M,Nn] (+)= A[M,K] . B[K,Nn].  A_FAST_M / B_FAST_N say which index is contiguous in
// memory (so the tile loads coalesce); C_FAST_M says which index of C is contiguous (so
// a warp's stores coalesce).  gridDim.z splits K into slabs of `kslab`.
template <class AR, class BR, class CW, bool A_FAST_M, bool B_FAST_N, bool C_FAST_M>
__global__ void __launch_bounds__(GT)
gemm_tile_kernel(AR A, BR Bm, CW Cw, long long M, int Nn, long long K, long long kslab) {
  __shared__ __align__(16) float As[BK][LD];
  __shared__ __align__(16) float Bs[BK][LD];
  const int t = threadIdx.x;
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long kbeg = (long long)blockIdx.z * kslab;
  const long long kend = min(K, kbeg + kslab);
  const int tr = C_FAST_M ? (t % 16) : (t / 16);  // micro-tile row group
  const int tc = C_FAST_M ? (t / 16) : (t % 16);  // micro-tile column group
  float acc[4][4] = {};

  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (A_FAST_M) { mm = t % BM; kk = t / BM + 4 * i; } else { kk = t % BK; mm = t / BK + 16 * i; }
      long long m = m0 + mm, kq = k0 + kk;
      As[kk][mm] = (m < M && kq < kend) ? A.at(m, kq) : 0.f;
      int nn, kb;
      if (B_FAST_N) { nn = t % BN; kb = t / BN + 4 * i; } else { kb = t % BK; nn = t / BK + 16 * i; }
      long long kq2 = k0 + kb;
      Bs[kb][nn] = (n0 + nn < Nn && kq2 < kend) ? Bm.at(kq2, n0 + nn) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][tr * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tc * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    long long m = m0 + tr * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tc * 4 + j;
      if (n < Nn) Cw.put(m, n, acc[i][j]);
    }
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ W, int Co, int C, int sub,
                                   float* __restrict__ Wcat) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * Co * C) return;
  int row = e / C, c = e % C;
  if (row < Co) {
    Wcat[e] = W[(size_t)row * 2 * C + c];
  } else {
    int o = row - Co;
    float w2 = W[(size_t)o * 2 * C + C + c];
    Wcat[e] = sub ? w2 - W[(size_t)o * 2 * C + c] : w2;
  }
}

// pack_weight + the tf32 hi/lo operand halves of Wcat and of Wcat^T in one pass (the forward
// point GEMM multiplies by Wcat [2Co,C], the backward dx GEMM by Wcat^T [C,2Co])
__global__ void prepare_weights_kernel(const float* __restrict__ W, int Co, int C, int sub,
                                       float* __restrict__ Wcat, float* __restrict__ hi,
                                       float* __restrict__ lo, float* __restrict__ hiT,
                                       float* __restrict__ loT) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * Co * C) return;
  int row = e / C, c = e % C;
  float v;
  if (row < Co) {
    v = W[(size_t)row * 2 * C + c];
  } else {
    int o = row - Co;
    float w2 = W[(size_t)o * 2 * C + C + c];
    v = sub ? w2 - W[(size_t)o * 2 * C + c] : w2;
  }
  Wcat[e] = v;
  const float h = ecb200::tf32_rna(v), l = ecb200::tf32_rna(v - h);
  if (hi) { hi[e] = h; lo[e] = l; }
  if (hiT) { hiT[(size_t)c * 2 * Co + row] = h; loT[(size_t)c * 2 * Co + row] = l; }
}

__global__ void unpack_weight_grad_kernel(const float* __restrict__ dWcat, int Co, int C, int sub,
                                          float* __restrict__ dW) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 2 * Co * C) return;
  int o = e / (2 * C), cc = e % (2 * C);
  if (cc < C) {
    float v = dWcat[(size_t)o * C + cc];
    if (sub) v -= dWcat[(size_t)(Co + o) * C + cc];  // V = (W2 - W1) x  =>  dW1 = dU^T X - dV^T X
    dW[e] = v;
  } else {
    dW[e] = dWcat[(size_t)(Co + o) * C + (cc - C)];
  }
}

}  // namespace

extern "C" int ecb200_pack_weight(const float* W, int Co, int C, int subtract_center, float* Wcat,
                                  void* stream) {
  ECB_REQUIRE(W && Wcat, "ecb200_pack_weight: null pointer");
  ECB_REQUIRE(Co >= 1 && C >= 1, "ecb200_pack_weight: bad shape Co=%d C=%d", Co, C);
  int n = 2 * Co * C;
  pack_weight_kernel<<<ecb200::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(W, Co, C,
                                                                                subtract_center, Wcat);
  ECB_LAUNCH_CHECK("pack_weight_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_unpack_weight_grad(const float* dWcat, int Co, int C, int subtract_center,
                                         float* dW, void* stream) {
  ECB_REQUIRE(dWcat && dW, "ecb200_unpack_weight_grad: null pointer");
  ECB_REQUIRE(Co >= 1 && C >= 1, "ecb200_unpack_weight_grad: bad shape Co=%d C=%d", Co, C);
  int n = 2 * Co * C;
  unpack_weight_grad_kernel<<<ecb200::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      dWcat, Co, C, subtract_center, dW);
  ECB_LAUNCH_CHECK("unpack_weight_grad_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_prepare_weights(const float* W, int Co, int C, int subtract_center, float* Wcat,
                                      float* hi, float* lo, float* hiT, float* loT, void* stream) {
  ECB_REQUIRE(W && Wcat, "ecb200_prepare_weights: null pointer");
  ECB_REQUIRE((hi == nullptr) == (lo == nullptr) && (hiT == nullptr) == (loT == nullptr),
              "ecb200_prepare_weights: hi/lo (and hiT/loT) come in pairs");
  ECB_REQUIRE(Co >= 1 && C >= 1, "ecb200_prepare_weights: bad shape Co=%d C=%d", Co, C);
  const int n = 2 * Co * C;
  prepare_weights_kernel<<<ecb200::ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
      W, Co, C, subtract_center, Wcat, hi, lo, hiT, loT);
  ECB_LAUNCH_CHECK("prepare_weights_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_point_gemm(const float* x, const float* Wcat, int B, int C, int N, int Co2,
                                 float* Y, void* stream) {
  ECB_REQUIRE(x && Wcat && Y, "ecb200_point_gemm: null pointer");
  ECB_REQUIRE(B >= 1 && C >= 1 && N >= 1 && Co2 >= 2, "ecb200_point_gemm: bad shape");
  long long M = (long long)B * N;
  dim3 grid((unsigned)ecb200::ceil_div64(M, BM), ecb200::ceil_div(Co2, BN), 1);
  // A = X^T [M,C] (m contiguous), B(k=c, n=o) = Wcat[o*C + c] (k contiguous), C = Y row-major
  gemm_tile_kernel<PointsRead, ColMajorRead, RowMajorWrite, true, false, false>
      <<<grid, GT, 0, (cudaStream_t)stream>>>(PointsRead{x, C, N}, ColMajorRead{Wcat, C},
                                              RowMajorWrite{Y, Co2}, M, Co2, C, C);
  ECB_LAUNCH_CHECK("gemm_tile_kernel<point_gemm>");
  return ECB200_OK;
}

extern "C" int ecb200_gemm_dx(const float* dY, const float* Wcat, int B, int C, int N, int Co2,
                              float* dx, void* stream) {
  ECB_REQUIRE(dY && Wcat && dx, "ecb200_gemm_dx: null pointer");
  ECB_REQUIRE(B >= 1 && C >= 1 && N >= 1 && Co2 >= 2, "ecb200_gemm_dx: bad shape");
  long long M = (long long)B * N;
  dim3 grid((unsigned)ecb200::ceil_div64(M, BM), ecb200::ceil_div(C, BN), 1);
  // A = dY [M,2Co] row-major (k contiguous), B = Wcat [2Co,C] row-major (n contiguous)
  gemm_tile_kernel<RowMajorRead, RowMajorRead, PointsWrite, false, true, true>
      <<<grid, GT, 0, (cudaStream_t)stream>>>(RowMajorRead{dY, Co2}, RowMajorRead{Wcat, C},
                                              PointsWrite{dx, C, N}, M, C, Co2, Co2);
  ECB_LAUNCH_CHECK("gemm_tile_kernel<gemm_dx>");
  return ECB200_OK;
}

// dWcat for a handful of input channels (the xyz layer, C = 3).  A block of 1024 threads =
// (1024 / 2Co) point lanes x 2Co output rows streams a slab of points: dY rows are read
// coalesced, the C coordinates of a point are a broadcast load.  The lanes are summed in shared
// memory and the block issues ONE atomic per (o, c): atomics onto the same few cache lines
// serialise in the L2, so there are only ~40 blocks.
template <int CC>
__global__ void __launch_bounds__(1024)
gemm_dw_smallc_kernel(const float* __restrict__ dY, const float* __restrict__ x, int C, int N, int Co2,
                      int slab, float* __restrict__ dWcat) {
  extern __shared__ float red[];  // [lanes][Co2][CC]
  const int lanes = blockDim.x / Co2;
  const int o = threadIdx.x % Co2, lane = threadIdx.x / Co2;
  const int bb = blockIdx.y;                    // one cloud per block row: no index division
  const int n0 = blockIdx.x * slab, n1 = min(N, n0 + slab);
  const float* dyb = dY + (size_t)bb * N * Co2 + o;
  const float* xb = x + (size_t)bb * C * N;
  float acc[CC];
#pragma unroll
  for (int c = 0; c < CC; ++c) acc[c] = 0.f;
  // 8 independent (dY, x) load groups in flight per thread: the loop is latency-bound
  for (int nb = n0 + lane; nb < n1; nb += 8 * lanes) {
    float dy[8], xv[8][CC];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int n = nb + u * lanes;
      const bool ok = n < n1;
      const int nn = ok ? n : n0;
      dy[u] = ok ? dyb[(size_t)nn * Co2] : 0.f;
#pragma unroll
      for (int c = 0; c < CC; ++c) xv[u][c] = c < C ? __ldg(xb + (size_t)c * N + nn) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int c = 0; c < CC; ++c) acc[c] = fmaf(dy[u], xv[u][c], acc[c]);
  }
#pragma unroll
  for (int c = 0; c < CC; ++c) red[((size_t)lane * Co2 + o) * CC + c] = acc[c];
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < CC; ++c) {
      float t = 0.f;
      for (int l = 0; l < lanes; ++l) t += red[((size_t)l * Co2 + o) * CC + c];
      if (c < C) atomicAdd(dWcat + (size_t)o * C + c, t);
    }
  }
}

extern "C" int ecb200_gemm_dw(const float* dY, const float* x, int B, int C, int N, int Co2,
                              float* dWcat, void* stream) {
  ECB_REQUIRE(dY && x && dWcat, "ecb200_gemm_dw: null pointer");
  ECB_REQUIRE(B >= 1 && C >= 1 && N >= 1 && Co2 >= 2, "ecb200_gemm_dw: bad shape");
  long long M = (long long)B * N;
  ECB_CUDA(cudaMemsetAsync(dWcat, 0, sizeof(float) * (size_t)Co2 * C, (cudaStream_t)stream));
  if (C <= 4 && Co2 <= 1024 && B <= 65535) {
    const int lanes = 1024 / Co2;
    const int threads = lanes * Co2;
    // about one block per SM: every block ends with one atomic per (o, c) on the same few lines
    int per_cloud = ecb200::ceil_div(ecb200::kNumSMs, B);
    int slab = ecb200::ceil_div(N, per_cloud);
    if (slab < 8 * lanes) slab = 8 * lanes;
    per_cloud = ecb200::ceil_div(N, slab);
    const size_t smem = sizeof(float) * (size_t)lanes * Co2 * 4;
    gemm_dw_smallc_kernel<4><<<dim3(per_cloud, B), threads, smem, (cudaStream_t)stream>>>(dY, x, C, N, Co2,
                                                                                         slab, dWcat);
    ECB_LAUNCH_CHECK("gemm_dw_smallc_kernel");
    return ECB200_OK;
  }
  // out[o,c] = sum_m dY[m,o] * X^T[m,c]:  A(o,m) = dY[m*Co2+o] (o contiguous),
  // B(m,c) = x[b,c,n] (m contiguous); the M-long reduction is cut into slabs.
  const int tiles = ecb200::ceil_div(Co2, BM) * ecb200::ceil_div(C, BN);
  long long slabs = (4LL * ecb200::kNumSMs + tiles - 1) / tiles;   // ~4 CTAs per SM in flight
  long long kslab = ecb200::ceil_div64(ecb200::ceil_div64(M, slabs), BK) * BK;
  if (kslab < 256) kslab = 256;
  slabs = ecb200::ceil_div64(M, kslab);
  dim3 grid(ecb200::ceil_div(Co2, BM), ecb200::ceil_div(C, BN), (unsigned)slabs);
  gemm_tile_kernel<ColMajorRead, PointsRead, RowMajorAtomic, true, false, false>
      <<<grid, GT, 0, (cudaStream_t)stream>>>(ColMajorRead{dY, Co2}, PointsRead{x, C, N},
                                              RowMajorAtomic{dWcat, C}, Co2, C, M, kslab);
  ECB_LAUNCH_CHECK("gemm_tile_kernel<gemm_dw>");
  return ECB200_OK;
}
