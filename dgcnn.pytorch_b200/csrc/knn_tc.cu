// Tensor-core kNN for the feature-space layers (C = 32..128 channels).
//
// The reference's knn() (models/dgcnn.py:6-12) is a dense contraction x^T x followed by a
// row-wise top-k.  Here the contraction runs on the 5th-generation tensor cores:
//   * operands are point-major [M, C] fp32 split into hi = tf32(x) and lo = tf32(x - hi)
//     (ecb200_split_tf32); a tile of 128 points x 32 channels is one TMA box landing in
//     shared memory in the canonical K-major SWIZZLE_128B layout;
//   * D[128 queries x 128 candidates] += Ahi.Bhi^T + Ahi.Blo^T + Alo.Bhi^T  (3xTF32 error
//     compensation, kind::tf32, FP32 accumulators in TMEM) -- one elected thread issues;
//   * four epilogue warps read the accumulators with tcgen05.ld (lane = query row), add the
//     -0.5*|x_j|^2 column term and feed the two-pass selector of topk_select.cuh, so the
//     distance matrix never leaves the SM.  Accumulators are double-buffered in TMEM so the
//     next tile's MMAs overlap the selection of the current one.
// The query tile (A) stays resident in shared memory; candidate tiles (B) stream through a
// TMA / mbarrier ring.  Pass A and pass B of the selector are two sweeps of the same MMAs.
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "topk_select.cuh"

namespace {

using namespace ecb200::tc;
using namespace ecb200::topk;

constexpr int BM = 128;                  // query rows per CTA (= TMEM lanes)
constexpr int BN = 128;                  // candidates per MMA tile (= TMEM columns per stage)
constexpr int KB = 32;                   // channels per K-block: 32 fp32 = one 128-byte swizzle row
constexpr int TILE_BYTES = BM * KB * 4;  // 16 KB: one K-block of one operand half
constexpr int MAX_KB = 4;                // C <= 128 keeps the query tile resident
constexpr int NUM_EPI = 128;             // 4 epilogue warps
constexpr int NT = 64 + NUM_EPI;         // + producer warp + MMA warp
constexpr uint32_t TMEM_COLS = 2 * BN;   // two accumulator stages
constexpr int UMMA_K = 8;                // tf32: 32 bytes of K per instruction

struct SharedTail {  // lives after the operand tiles
  float hx[2][BN];
  uint64_t a_full, b_full[4], b_empty[4], t_full[2], t_empty[2];
  uint32_t tmem_slot;
};

__host__ __device__ constexpr int num_stages(int nkb) { return nkb <= 2 ? 4 : 2; }
__host__ __device__ constexpr size_t smem_bytes(int nkb) {
  return 1024 /* alignment slack */ + (size_t)(2 * nkb + 2 * num_stages(nkb)) * TILE_BYTES +
         sizeof(SharedTail);
}

template <int NBINS>
__device__ __forceinline__ void sort_bins_desc(float (&v)[NBINS]) {
#pragma unroll
  for (int size = 2; size <= NBINS; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < NBINS; ++i) {
        const int j = i ^ stride;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const float a = v[i], b = v[j];
          v[i] = desc ? fmaxf(a, b) : fminf(a, b);
          v[j] = desc ? fminf(a, b) : fmaxf(a, b);
        }
      }
    }
  }
}

// DEBUG = true: one sweep, raw scores written to dbg[B,N,N] (validation of the MMA plumbing)
template <int NBINS, bool DEBUG>
__global__ void __launch_bounds__(NT, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
              const float* __restrict__ xx, int N, int nkb, int k, uint64_t* __restrict__ surv_ws,
              int cap, int32_t* __restrict__ idx, float* __restrict__ dbg) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base =
      reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int S = num_stages(nkb);
  unsigned char* a_hi = base;                                  // [nkb][16 KB]
  unsigned char* a_lo = a_hi + (size_t)nkb * TILE_BYTES;       // [nkb][16 KB]
  unsigned char* b_st = a_lo + (size_t)nkb * TILE_BYTES;       // [S][hi 16 KB | lo 16 KB]
  SharedTail* T = reinterpret_cast<SharedTail*>(b_st + (size_t)S * 2 * TILE_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, rt = blockIdx.x;
  const int nct = (N + BN - 1) / BN;
  const int npass = DEBUG ? 1 : 2;
  const int cloud_row0 = b * N;  // first global row of this cloud in the [M, C] arrays

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_hi);
    prefetch_tensormap(&map_lo);
    mbar_init(&T->a_full, 1);
    for (int s = 0; s < S; ++s) { mbar_init(&T->b_full[s], 1); mbar_init(&T->b_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&T->t_full[s], 1); mbar_init(&T->t_empty[s], NUM_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&T->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T->tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      mbar_expect_tx(&T->a_full, (uint32_t)(2 * nkb * TILE_BYTES));
      for (int kb = 0; kb < nkb; ++kb) {
        tma_load_2d(a_hi + (size_t)kb * TILE_BYTES, &map_hi, &T->a_full, kb * KB, cloud_row0 + rt * BM);
        tma_load_2d(a_lo + (size_t)kb * TILE_BYTES, &map_lo, &T->a_full, kb * KB, cloud_row0 + rt * BM);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int pass = 0; pass < npass; ++pass)
        for (int ct = 0; ct < nct; ++ct)
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&T->b_empty[stage], phase ^ 1);
            unsigned char* dst = b_st + (size_t)stage * 2 * TILE_BYTES;
            mbar_expect_tx(&T->b_full[stage], 2 * TILE_BYTES);
            tma_load_2d(dst, &map_hi, &T->b_full[stage], kb * KB, cloud_row0 + ct * BN);
            tma_load_2d(dst + TILE_BYTES, &map_lo, &T->b_full[stage], kb * KB, cloud_row0 + ct * BN);
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_tf32(BM, BN);
      mbar_wait(&T->a_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int tile = 0;
      for (int pass = 0; pass < npass; ++pass)
        for (int ct = 0; ct < nct; ++ct, ++tile) {
          const int as = tile & 1;
          mbar_wait(&T->t_empty[as], ((tile >> 1) & 1) ^ 1);  // epilogue drained this stage
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait(&T->b_full[stage], phase);
            tc_fence_after();
            const uint32_t ah = smem_u32(a_hi + (size_t)kb * TILE_BYTES);
            const uint32_t al = smem_u32(a_lo + (size_t)kb * TILE_BYTES);
            const uint32_t bh = smem_u32(b_st + (size_t)stage * 2 * TILE_BYTES);
            const uint32_t bl = bh + TILE_BYTES;
#pragma unroll
            for (int k8 = 0; k8 < KB / UMMA_K; ++k8) {
              const uint32_t ko = (uint32_t)(k8 * UMMA_K * 4);  // byte offset inside the swizzle row
              const uint64_t dah = make_sw128_kmajor_desc(ah + ko), dal = make_sw128_kmajor_desc(al + ko);
              const uint64_t dbh = make_sw128_kmajor_desc(bh + ko), dbl = make_sw128_kmajor_desc(bl + ko);
              mma_tf32(d_tmem, dah, dbh, idesc, (kb | k8) != 0);
              mma_tf32(d_tmem, dah, dbl, idesc, 1);
              mma_tf32(d_tmem, dal, dbh, idesc, 1);
            }
            mma_commit(&T->b_empty[stage]);  // frees the stage once these MMAs have read it
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          mma_commit(&T->t_full[as]);  // accumulator ready for the epilogue
        }
    }
  } else {
    // ===================== epilogue: selection (thread = query row) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 64;
    const int row = rt * BM + q * 32 + lane;
    const bool valid = row < N;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    float bin[NBINS];
#pragma unroll
    for (int u = 0; u < NBINS; ++u) bin[u] = -CUDART_INF_F;
    uint64_t* mybuf = surv_ws ? surv_ws + (size_t)(cloud_row0 + (valid ? row : 0)) * cap : nullptr;
    int cnt = 0;
    float thr = CUDART_INF_F;
    int tile = 0;
    for (int pass = 0; pass < npass; ++pass) {
      if (pass == 1) {
        sort_bins_desc<NBINS>(bin);
        float tau = -CUDART_INF_F;
#pragma unroll
        for (int u = 0; u < NBINS; ++u)
          if (u == k - 1) tau = bin[u];
        thr = fmaxf(tau, -3.0e38f);  // masked candidates score -inf and must never pass
      }
      for (int ct = 0; ct < nct; ++ct, ++tile) {
        const int as = tile & 1;
        {
          const int j = ct * BN + et;
          T->hx[as][et] = (j < N) ? -0.5f * xx[cloud_row0 + j] : -CUDART_INF_F;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI) : "memory");  // epilogue warps only
        mbar_wait(&T->t_full[as], (tile >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c4 = 0; c4 < BN / 32; ++c4) {
          float v[32];
          __syncwarp();  // tcgen05.ld is warp-collective (.sync.aligned)
          tmem_ld_32x32(lane_base + (uint32_t)(as * BN + c4 * 32), v);
          const float4* hx4 = reinterpret_cast<const float4*>(&T->hx[as][c4 * 32]);
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            const float4 h4 = hx4[g];
            v[4 * g + 0] += h4.x; v[4 * g + 1] += h4.y; v[4 * g + 2] += h4.z; v[4 * g + 3] += h4.w;
          }
          if (DEBUG) {
            if (valid) {
#pragma unroll
              for (int u = 0; u < 32; ++u) {
                const int j = ct * BN + c4 * 32 + u;
                if (j < N) dbg[((size_t)(cloud_row0 + row)) * N + j] = v[u];
              }
            }
          } else if (pass == 0) {
            constexpr int HALVES = NBINS / 32;
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              const int bi = (c4 % HALVES) * 32 + u;
              bin[bi] = fmaxf(bin[bi], v[u]);
            }
          } else if (valid) {
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (cnt > cap - 8) {  // slow path: keep the row's own best k, raise the bar
                while (cnt > k) {
                  int arg = 0;
                  uint64_t mn = mybuf[0];
                  for (int e = 1; e < cnt; ++e) {
                    const uint64_t w = mybuf[e];
                    if (w < mn) { mn = w; arg = e; }
                  }
                  --cnt;
                  mybuf[arg] = mybuf[cnt];
                }
                uint64_t mn = mybuf[0];
                for (int e = 1; e < cnt; ++e) mn = min(mn, mybuf[e]);
                thr = fmaxf(thr, nextafterf(key_score(mn), CUDART_INF_F));
              }
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float s = v[g * 8 + u];
                if (s >= thr) {
                  mybuf[cnt] = make_key(s, ct * BN + c4 * 32 + g * 8 + u);
                  ++cnt;
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&T->t_empty[as]);
      }
    }
    if (!DEBUG && valid) {
      // exact top-k of the survivors: rank = number of strictly better keys = output slot
      int32_t* out = idx + (size_t)(cloud_row0 + row) * k;
      for (int p = cnt; p < k; ++p) out[p] = N - 1;  // only with NaN input
      for (int e = 0; e < cnt; ++e) {
        const uint64_t key = mybuf[e];
        int rank = 0;
        for (int f = 0; f < cnt; ++f) rank += (mybuf[f] > key);
        if (rank < k) out[rank] = (int32_t)min(key_index(key), (uint32_t)(N - 1));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// x[B,C,N] -> point-major hi/lo [M,C] (tf32-rounded halves of the fp32 value) and xx[M]
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, int C, int N, float* __restrict__ hi,
                  float* __restrict__ lo, float* __restrict__ xx) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n0 = blockIdx.x * 32, bb = blockIdx.y;
  const float* xb = x + (size_t)bb * C * N;
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = c0 + ty + 8 * r, n = n0 + tx;
      tile[ty + 8 * r][tx] = (c < C && n < N) ? xb[(size_t)c * N + n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int n = n0 + ty + 8 * r, c = c0 + tx;
      if (n < N && c < C) {
        const float v = tile[tx][ty + 8 * r];
        const float h = to_tf32(v);
        const size_t o = ((size_t)bb * N + n) * C + c;
        hi[o] = h;
        lo[o] = to_tf32(v - h);
      }
    }
    __syncthreads();
  }
  if (ty == 0 && n0 + tx < N) {  // same summation order as ecb200_sqnorms
    float s = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = xb[(size_t)c * N + n0 + tx];
      s = fmaf(v, v, s);
    }
    xx[(size_t)bb * N + n0 + tx] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, C] fp32 row-major, box = 32 channels x 128 rows, 128-byte swizzle, zero fill past the end
int make_point_map(CUtensorMap* m, const float* p, long long rows, int C) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    ecb200::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECB200_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
  const cuuint32_t box[2] = {KB, BM};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecb200::set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ECB200_ERR_CUDA;
  }
  return ECB200_OK;
}

template <int NBINS, bool DEBUG>
int launch_tc(const float* hi, const float* lo, const float* xx, int B, int C, int N, int k,
              uint64_t* ws, int cap, int32_t* idx, float* dbg, cudaStream_t st) {
  const int nkb = C / KB;
  CUtensorMap mh, ml;
  int rc = make_point_map(&mh, hi, (long long)B * N, C);
  if (rc) return rc;
  rc = make_point_map(&ml, lo, (long long)B * N, C);
  if (rc) return rc;
  auto kern = knn_tc_kernel<NBINS, DEBUG>;
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if (ecb200::first_use_on_device(seen))
    ECB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  (int)smem_bytes(MAX_KB)));
  dim3 grid(ecb200::ceil_div(N, BM), B);
  kern<<<grid, NT, smem_bytes(nkb), st>>>(mh, ml, xx, N, nkb, k, ws, cap, idx, dbg);
  ECB_LAUNCH_CHECK("knn_tc_kernel");
  return ECB200_OK;
}

}  // namespace

extern "C" int ecb200_split_tf32(const float* x, int B, int C, int N, float* hi, float* lo, float* xx,
                                 void* stream) {
  ECB_REQUIRE(x && hi && lo && xx, "ecb200_split_tf32: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && N >= 1, "ecb200_split_tf32: bad shape");
  dim3 grid(ecb200::ceil_div(N, 32), B);
  split_tf32_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, C, N, hi, lo, xx);
  ECB_LAUNCH_CHECK("split_tf32_kernel");
  return ECB200_OK;
}

// survivor slots per row in the workspace (expected use: ~31 at k = 20 with 32 bins, ~63 at
// k = 40 with 64 bins; overflow falls back to an exact in-place shrink)
static int survivor_cap(int k) { return k <= 20 ? 96 : 160; }

extern "C" size_t ecb200_knn_tc_workspace_bytes(int B, int N, int k) {
  return (size_t)B * N * (size_t)survivor_cap(k) * sizeof(uint64_t);
}

extern "C" int ecb200_knn_tc(const float* hi, const float* lo, const float* xx, int B, int C, int N,
                             int k, int sorted, int32_t* idx, void* workspace, size_t workspace_bytes,
                             void* stream) {
  (void)sorted;  // the rank-based final stage always yields nearest-first order
  ECB_REQUIRE(hi && lo && xx && idx && workspace, "ecb200_knn_tc: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_knn_tc: bad shape B=%d N=%d", B, N);
  ECB_REQUIRE(C % KB == 0 && C >= KB && C <= KB * MAX_KB,
              "ecb200_knn_tc: C=%d must be a multiple of 32 in [32, 128]", C);
  ECB_REQUIRE(k >= 1 && k <= N, "ecb200_knn_tc: k=%d out of range for N=%d (selected index k out of range)", k, N);
  ECB_REQUIRE(k <= 40, "ecb200_knn_tc: k=%d exceeds 40 (use ecb200_knn)", k);
  ECB_REQUIRE(workspace_bytes >= ecb200_knn_tc_workspace_bytes(B, N, k),
              "ecb200_knn_tc: workspace too small (%zu < %zu bytes)", workspace_bytes,
              ecb200_knn_tc_workspace_bytes(B, N, k));
  const int cap = survivor_cap(k);
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 20)
    return launch_tc<32, false>(hi, lo, xx, B, C, N, k, (uint64_t*)workspace, cap, idx, nullptr, st);
  return launch_tc<64, false>(hi, lo, xx, B, C, N, k, (uint64_t*)workspace, cap, idx, nullptr, st);
}

extern "C" int ecb200_debug_tc_scores(const float* hi, const float* lo, const float* xx, int B, int C,
                                      int N, float* scores, void* stream) {
  ECB_REQUIRE(hi && lo && xx && scores, "ecb200_debug_tc_scores: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_debug_tc_scores: bad shape");
  ECB_REQUIRE(C % KB == 0 && C >= KB && C <= KB * MAX_KB,
              "ecb200_debug_tc_scores: C=%d must be a multiple of 32 in [32, 128]", C);
  return launch_tc<32, true>(hi, lo, xx, B, C, N, 1, nullptr, 0, nullptr, scores, (cudaStream_t)stream);
}
