// kNN graph by FP32-FMA distance tiles feeding an on-chip top-k selector.
// Replaces knn() of /root/reference/models/dgcnn.py:6-12: the [B,N,N] matrix of
// -|xi|^2 + 2 xi.xj - |xj|^2 is produced tile by tile in registers and consumed
// immediately by the two-pass top-k selector of topk_select.cuh; only idx[B,N,k] reaches HBM.
//
// Ranking key.  For a fixed query i the reference's score differs from
//   s_ij = xi.xj - 0.5*|xj|^2
// only by the row constant -|xi|^2 and a factor 2, so s_ij ranks identically.
// The selector orders candidates by the 64-bit key (orderable(s) << 32 | ~j):
// larger score first, ties towards the smaller index j -- a total order, so the
// result does not depend on how candidates are split between threads.
#include <math_constants.h>

#include "common.cuh"
#include "topk_select.cuh"

namespace {

using namespace ecb200::topk;

constexpr int R = 64;         // query rows per CTA
constexpr int TJ = 64;        // candidates per tile
constexpr int NT = 128;       // threads per CTA: two per row, one per half of each tile
constexpr int HALF = TJ / 2;  // candidates a thread scores per tile (== NB register bins)
constexpr int KC_MAX = 128;   // channels staged in shared memory per chunk
constexpr int OFFER = 8;      // candidates offered between overflow checks
constexpr int CAP_MIN = 48;   // survivor slots per thread (expected use: ~12 at k=20, ~24 at k=40)
static_assert(HALF == NB, "one register bin per candidate slot of a tile");
using Surv = Survivors<NT>;

struct Tiles {
  float* Qs;  // [KC][R]   query rows, c-major
  float* Cs;  // [KC][TJ]  candidate tile, c-major
  float* hx;  // [TJ]      -0.5*|xj|^2, -inf past the end of the cloud
};

// scores of this thread's 32 candidates of tile j0:  acc[u] = xi.xj - 0.5*|xj|^2
__device__ __forceinline__ void score_tile(const float* __restrict__ xb, const float* __restrict__ xxb,
                                           int C, int N, int KC, int j0, int row0, int tid, int r,
                                           int h, const Tiles& t, float (&acc)[HALF]) {
  const int nchunks = (C + KC - 1) / KC;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int c0 = ch * KC;
    const int kc = min(KC, C - c0);
    __syncthreads();  // previous tile / chunk fully consumed
    for (int e = tid; e < kc * TJ; e += NT) {
      int c = e / TJ, jj = e % TJ;
      t.Cs[e] = (j0 + jj < N) ? xb[(size_t)(c0 + c) * N + j0 + jj] : 0.f;
    }
    if (nchunks > 1 || j0 == 0) {
      for (int e = tid; e < kc * R; e += NT) {
        int c = e / R, rr = e % R;
        t.Qs[e] = (row0 + rr < N) ? xb[(size_t)(c0 + c) * N + row0 + rr] : 0.f;
      }
    }
    if (ch == 0 && tid < TJ) t.hx[tid] = (j0 + tid < N) ? -0.5f * xxb[j0 + tid] : -CUDART_INF_F;
    __syncthreads();
    if (ch == 0) {
#pragma unroll
      for (int u = 0; u < HALF; ++u) acc[u] = t.hx[h * HALF + u];
    }
#pragma unroll 2
    for (int c = 0; c < kc; ++c) {
      const float q = t.Qs[c * R + r];
      const float4* cp = reinterpret_cast<const float4*>(t.Cs + c * TJ + h * HALF);
#pragma unroll
      for (int v = 0; v < HALF / 4; ++v) {
        float4 w = cp[v];  // same address in every lane: broadcast
        acc[4 * v + 0] = fmaf(q, w.x, acc[4 * v + 0]);
        acc[4 * v + 1] = fmaf(q, w.y, acc[4 * v + 1]);
        acc[4 * v + 2] = fmaf(q, w.z, acc[4 * v + 2]);
        acc[4 * v + 3] = fmaf(q, w.w, acc[4 * v + 3]);
      }
    }
  }
}

__global__ void __launch_bounds__(NT)
knn_fma_kernel(const float* __restrict__ x, const float* __restrict__ xx, int C, int N, int k,
               int KC, int cap, int sorted, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* sbin = reinterpret_cast<float*>(smem_raw + Surv::smem_bytes(cap));  // [NB][NT] sorted bin maxima
  int* scnt = reinterpret_cast<int*>(sbin + NB * NT);                      // [NT] survivor counts
  Tiles t;
  t.Qs = reinterpret_cast<float*>(scnt + NT);
  t.Cs = t.Qs + KC * R;
  t.hx = t.Cs + KC * TJ;

  const int b = blockIdx.y;
  const int row0 = blockIdx.x * R;
  const int tid = threadIdx.x;
  const int r = tid % R;
  const int h = tid / R;  // warp-uniform: which half of each candidate tile
  const float* xb = x + (size_t)b * C * N;
  const float* xxb = xx + (size_t)b * N;
  float acc[HALF];

  // ---- pass A: bin maxima -> per-row threshold
  float bin[NB];
#pragma unroll
  for (int u = 0; u < NB; ++u) bin[u] = -CUDART_INF_F;
  for (int j0 = 0; j0 < N; j0 += TJ) {
    score_tile(xb, xxb, C, N, KC, j0, row0, tid, r, h, t, acc);
#pragma unroll
    for (int u = 0; u < NB; ++u) bin[u] = fmaxf(bin[u], acc[u]);
  }
  sort32_desc(bin);
#pragma unroll
  for (int u = 0; u < NB; ++u) sbin[u * NT + tid] = bin[u];
  __syncthreads();
  // both threads of a row derive the same tau from the row's 64 bins
  const float tau = fmaxf(kth_of_sorted_columns<2>(sbin, r, R, NT, k), -3.0e38f);

  // ---- pass B: keep the candidates that reach tau
  Surv sv;
  sv.init(smem_raw, tid, cap, tau);
  for (int j0 = 0; j0 < N; j0 += TJ) {
    score_tile(xb, xxb, C, N, KC, j0, row0, tid, r, h, t, acc);
    const int jbase = j0 + h * HALF;
#pragma unroll
    for (int g = 0; g < HALF / OFFER; ++g) {
      sv.guard(k, OFFER);
#pragma unroll
      for (int u = 0; u < OFFER; ++u) {
        const int j = jbase + g * OFFER + u;
        if (j < N) sv.offer(acc[g * OFFER + u], j);
      }
    }
  }
  scnt[tid] = sv.cnt;
  __syncthreads();

  // ---- exact top-k: both threads of a row rank their own survivors against the union
  if (row0 + r < N) {
    const int cnts[2] = {scnt[r], scnt[r + R]};
    int32_t* out = idx + ((size_t)b * N + row0 + r) * k;
    // fewer than k survivors only if the input held NaNs: keep every slot in range
    if (h == 0 && cnts[0] + cnts[1] < k)
      for (int p = cnts[0] + cnts[1]; p < k; ++p) out[p] = N - 1;
    rank_and_write<NT, 2>(reinterpret_cast<const uint64_t*>(smem_raw), r, R, cnts, h, k, N - 1, out);
  }
}

// ---- xyz specialisation (C <= 3): the whole cloud sits in shared memory as packed
// (x, y, z, -0.5*|p|^2); a row's coordinates live in registers; 4 threads share a row,
// each scanning every fourth group of 32 candidates.  ~5 instructions per pair and pass.
constexpr int XR = 32;        // rows per CTA
constexpr int XT = 4;         // threads per row
constexpr int XNT = XR * XT;  // 128 threads
constexpr int XCAP_MIN = 24;  // survivor slots per thread (expected use: ~6 at k=20, ~12 at k=40); >= NB/2 (bins alias them)
constexpr int XTILE = 4096;   // candidates staged at a time (64 KB)
using XSurv = Survivors<XNT>;

__device__ __forceinline__ uint64_t raw_to_key(uint64_t raw) {
  return make_key(__uint_as_float((uint32_t)(raw >> 32)), (int)(uint32_t)raw);
}
__device__ __forceinline__ uint64_t key_to_raw(uint64_t key) {
  return ((uint64_t)__float_as_uint(key_score(key)) << 32) | key_index(key);
}

template <int C>
__global__ void __launch_bounds__(XNT)
knn_xyz_kernel(const float* __restrict__ x, const float* __restrict__ xx, int N, int k, int cap,
               int sorted, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the sorted bin maxima [NB][XNT] are only needed between the passes: they share the memory of
  // the survivor lists (cap >= 16 slots of 8 bytes per thread >= NB floats per thread)
  float* sbin = reinterpret_cast<float*>(smem_raw);
  int* scnt = reinterpret_cast<int*>(smem_raw + XSurv::smem_bytes(cap));    // [XNT]
  float4* cand = reinterpret_cast<float4*>(scnt + XNT);                     // [<= XTILE]

  const int b = blockIdx.y;
  const int row0 = blockIdx.x * XR;
  const int tid = threadIdx.x;
  const int r = tid % XR;
  const int h = tid / XR;  // warp index: which group of 32 within each 128 candidates
  const float* xb = x + (size_t)b * C * N;
  const float* xxb = xx + (size_t)b * N;
  const int row = min(row0 + r, N - 1);
  const float q0 = xb[row];
  const float q1 = C > 1 ? xb[(size_t)N + row] : 0.f;
  const float q2 = C > 2 ? xb[2 * (size_t)N + row] : 0.f;

  auto stage = [&](int t0, int tn) {  // candidates [t0, t0+tn) -> smem, padded to 128 with -inf
    __syncthreads();
    const int padded = (tn + XNT - 1) / XNT * XNT;
    for (int e = tid; e < padded; e += XNT) {
      float4 c = make_float4(0.f, 0.f, 0.f, -CUDART_INF_F);
      if (e < tn) {
        const int j = t0 + e;
        c.x = xb[j];
        if (C > 1) c.y = xb[(size_t)N + j];
        if (C > 2) c.z = xb[2 * (size_t)N + j];
        c.w = -0.5f * xxb[j];
      }
      cand[e] = c;
    }
    __syncthreads();
    return padded;
  };
  auto score = [&](const float4& c) { return fmaf(q0, c.x, fmaf(q1, c.y, fmaf(q2, c.z, c.w))); };

  // ---- pass A: bin maxima -> threshold
  float bin[NB];
#pragma unroll
  for (int u = 0; u < NB; ++u) bin[u] = -CUDART_INF_F;
  for (int t0 = 0; t0 < N; t0 += XTILE) {
    const int padded = stage(t0, min(XTILE, N - t0));
    for (int j0 = h * NB; j0 < padded; j0 += XNT) {
#pragma unroll
      for (int u = 0; u < NB; ++u) bin[u] = fmaxf(bin[u], score(cand[j0 + u]));
    }
  }
  sort32_desc(bin);
#pragma unroll
  for (int u = 0; u < NB; ++u) sbin[u * XNT + tid] = bin[u];
  __syncthreads();
  // -3e38 floor: padding candidates score -inf and must never pass, even when the cloud has
  // fewer than k non-empty bins
  const float tau = fmaxf(kth_of_sorted_columns<XT>(sbin, r, XR, XNT, k), -3.0e38f);
  __syncthreads();  // every thread has read the bins: their memory now holds the survivor lists

  // ---- pass B: survivors, appended as raw (score bits, j) words -- the ordered key is only
  //      built for the few entries that survive, after the sweep
  XSurv sv;
  sv.init(smem_raw, tid, cap, tau);
  for (int t0 = 0; t0 < N; t0 += XTILE) {
    const int padded = (N <= XTILE) ? (N + XNT - 1) / XNT * XNT : stage(t0, min(XTILE, N - t0));
    for (int j0 = h * NB; j0 < padded; j0 += XNT) {
#pragma unroll
      for (int g = 0; g < NB / 8; ++g) {
        if (sv.cnt > sv.cap - 8) {  // rare: keep the thread's own best k (works on ordered keys)
          for (int e = 0; e < sv.cnt; ++e) sv.buf[e * XNT] = raw_to_key(sv.buf[e * XNT]);
          sv.shrink_to(k);
          for (int e = 0; e < sv.cnt; ++e) sv.buf[e * XNT] = key_to_raw(sv.buf[e * XNT]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float sc = score(cand[j0 + g * 8 + u]);
          if (sc >= sv.thr) {
            sv.buf[sv.cnt * XNT] = ((uint64_t)__float_as_uint(sc) << 32) | (uint32_t)(t0 + j0 + g * 8 + u);
            ++sv.cnt;
          }
        }
      }
    }
  }
  for (int e = 0; e < sv.cnt; ++e) sv.buf[e * XNT] = raw_to_key(sv.buf[e * XNT]);
  scnt[tid] = sv.cnt;
  __syncthreads();

  // ---- exact top-k: the four threads of a row rank their own survivors against the union
  if (row0 + r < N) {
    int cnts[XT];
    int m = 0;
#pragma unroll
    for (int t = 0; t < XT; ++t) { cnts[t] = scnt[r + t * XR]; m += cnts[t]; }
    int32_t* out = idx + ((size_t)b * N + row0 + r) * k;
    if (h == 0 && m < k)  // only with NaN input
      for (int p = m; p < k; ++p) out[p] = N - 1;
    rank_and_write<XNT, XT>(reinterpret_cast<const uint64_t*>(smem_raw), r, XR, cnts, h, k, N - 1, out);
  }
}

__global__ void sqnorms_kernel(const float* __restrict__ x, int C, int N, long long M,
                               float* __restrict__ xx) {
  long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  long long b = m / N;
  int n = (int)(m - b * N);
  const float* p = x + (size_t)b * C * N + n;
  float s = 0.f;
  for (int c = 0; c < C; ++c) {
    float v = p[(size_t)c * N];
    s = fmaf(v, v, s);
  }
  xx[m] = s;
}

}  // namespace

extern "C" int ecb200_sqnorms(const float* x, int B, int C, int N, float* xx, void* stream) {
  ECB_REQUIRE(x && xx, "ecb200_sqnorms: null pointer");
  ECB_REQUIRE(B >= 1 && C >= 1 && N >= 1, "ecb200_sqnorms: bad shape B=%d C=%d N=%d", B, C, N);
  long long M = (long long)B * N;
  sqnorms_kernel<<<(unsigned)ecb200::ceil_div64(M, 256), 256, 0, (cudaStream_t)stream>>>(x, C, N, M,
                                                                                        xx);
  ECB_LAUNCH_CHECK("sqnorms_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_knn(const float* x, const float* xx, int B, int C, int N, int k, int sorted,
                          int32_t* idx, void* stream) {
  ECB_REQUIRE(x && xx && idx, "ecb200_knn: null pointer");
  ECB_REQUIRE(B >= 1 && C >= 1 && N >= 1, "ecb200_knn: bad shape B=%d C=%d N=%d", B, C, N);
  ECB_REQUIRE(B <= 65535, "ecb200_knn: B=%d exceeds 65535 clouds per call", B);
  // same trigger as Tensor.topk in the reference (dgcnn.py:11): k must not exceed N
  ECB_REQUIRE(k >= 1 && k <= N, "ecb200_knn: k=%d out of range for N=%d (selected index k out of range)", k, N);
  ECB_REQUIRE(k <= ECB200_MAX_K, "ecb200_knn: k=%d exceeds ECB200_MAX_K=%d", k, ECB200_MAX_K);
  if (C <= 3) {
    const int staged = N < XTILE ? (N + XNT - 1) / XNT * XNT : XTILE;
    const int cap = XSurv::capacity(k, 8, XCAP_MIN);
    const size_t smem = XSurv::smem_bytes(cap) + XNT * sizeof(int) + (size_t)staged * sizeof(float4);
    const size_t smem_max = XSurv::smem_bytes(XSurv::capacity(ECB200_MAX_K, 8, XCAP_MIN)) +
                            XNT * sizeof(int) + (size_t)XTILE * sizeof(float4);
    static thread_local bool seen_xyz[ecb200::kMaxDevices] = {};
    if (ecb200::first_use_on_device(seen_xyz)) {
      ECB_CUDA(cudaFuncSetAttribute(knn_xyz_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      ECB_CUDA(cudaFuncSetAttribute(knn_xyz_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
      ECB_CUDA(cudaFuncSetAttribute(knn_xyz_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
    }
    dim3 grid(ecb200::ceil_div(N, XR), B);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 1) knn_xyz_kernel<1><<<grid, XNT, smem, st>>>(x, xx, N, k, cap, sorted, idx);
    else if (C == 2) knn_xyz_kernel<2><<<grid, XNT, smem, st>>>(x, xx, N, k, cap, sorted, idx);
    else knn_xyz_kernel<3><<<grid, XNT, smem, st>>>(x, xx, N, k, cap, sorted, idx);
    ECB_LAUNCH_CHECK("knn_xyz_kernel");
    return ECB200_OK;
  }
  const int KC = C < KC_MAX ? C : KC_MAX;
  const int cap = Surv::capacity(k, OFFER, CAP_MIN);
  const size_t misc = (size_t)NB * NT * sizeof(float) + NT * sizeof(int);
  const size_t smem = Surv::smem_bytes(cap) + misc + (size_t)KC * (R + TJ) * sizeof(float) +
                      TJ * sizeof(float);
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if (ecb200::first_use_on_device(seen)) {
    const size_t smem_max = Surv::smem_bytes(Surv::capacity(ECB200_MAX_K, OFFER, CAP_MIN)) + misc +
                            (size_t)KC_MAX * (R + TJ) * sizeof(float) + TJ * sizeof(float);
    ECB_CUDA(cudaFuncSetAttribute(knn_fma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_max));
  }
  dim3 grid(ecb200::ceil_div(N, R), B);
  knn_fma_kernel<<<grid, NT, smem, (cudaStream_t)stream>>>(x, xx, C, N, k, KC, cap, sorted, idx);
  ECB_LAUNCH_CHECK("knn_fma_kernel");
  return ECB200_OK;
}
