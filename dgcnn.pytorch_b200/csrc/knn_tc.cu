// Tensor-core kNN (128 query rows per CTA): feature-space layers and, through the same pipeline, the xyz layer.
//
// The reference's knn() (models/dgcnn.py:6-12) is a dense contraction x^T x followed by a
// row-wise top-k.  Here the contraction runs on the 5th-generation tensor cores:
//   * operands are point-major [M, C] split into hi + lo halves with 11-bit significands: PACKED FP16
//     pairs of the tensor scaled by a power of two (ecb200_split_f16, kind::f16: twice the MMA rate, the
//     default for C = 64 / 128) or tf32 values (ecb200_split_tf32, kind::tf32: C = 32 / 96).  A tile of
//     128 points x 32 words is one TMA box landing in shared memory in the canonical K-major
//     SWIZZLE_128B layout; in words the two operand types look the same to the kernel;
//   * D[128 queries x 128 candidates] += Ahi.Bhi^T + Ahi.Blo^T + Alo.Bhi^T  (error-compensated product,
//     FP32 accumulators in TMEM) -- one elected thread issues; with packed FP16 the column term
//     -0.5*|x_j|^2 is three more K slots of the same contraction (FOLD), so the accumulator is the score;
//   * two epilogue warpgroups (one per TMEM accumulator stage) read the accumulators with
//     tcgen05.ld (lane = query row) and feed the two-pass selector of topk_select.cuh, so the distance
//     matrix never leaves the SM; MMAs of the next tile overlap the selection of the current one;
//   * xyz layer (C <= 4): the three product terms and the column term of a point sit in ONE 16-deep K
//     step (TERMS = 1), one MMA per tile and sweep.
// The query tile (A) is copied once into tensor memory (TS-mode MMAs); candidate tiles (B) stream through a
// TMA / mbarrier ring, for the packed-FP16 kernels through per-cloud 3-D maps whose out-of-range rows read
// as NaN (a candidate that does not exist can never be selected).  Pass A and pass B of the selector are
// two sweeps of the same tiles.  The dense-store variant (DEBUG) of this kernel is the per-point GEMM
// Y = X.Wcat^T of the EdgeConv layers.  knn_tc2.cu is the 256-row sibling.
#include <cuda.h>
#include <cuda_fp16.h>
#include <math_constants.h>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "topk_select.cuh"
#include "knn_tc_shared.cuh"

namespace {

using namespace ecb200::tc;
using namespace ecb200::topk;
using namespace ecb200::knntc;

constexpr int BM = 128;                  // query rows per CTA (= TMEM lanes)
constexpr int BN = 128;                  // candidates per MMA tile (= TMEM columns per stage)
constexpr int TILE_BYTES = BM * KB * 4;  // 16 KB: one K-block of one operand half
constexpr int STAGE_BYTES = 2 * TILE_BYTES;  // a ring stage: (hi | lo) of a K-block, or two hi K-blocks
constexpr int MAX_KB = 4;                // C <= 128 keeps the query tile resident
constexpr int NUM_EPI = 128;             // threads per epilogue warpgroup (one per TMEM lane)
constexpr int NT = 64 + 2 * NUM_EPI;     // producer warp + MMA warp + two epilogue warpgroups
constexpr int LS = 2 * NUM_EPI;          // stride (in entries) between a thread's consecutive survivor slots
constexpr uint32_t TMEM_COLS = 512;      // two accumulator stages (2 x 128 columns) + the query tile
constexpr uint32_t A_COL0 = 2 * BN;      // hi at columns [256, 256+C), lo at [256+C, 256+2C)
constexpr int UMMA_K = 8;                // tf32: 32 bytes of K per instruction
constexpr int KMAX = 40;                 // largest k this kernel takes
constexpr int GUARD = 16;                // candidate columns between two overflow checks of a survivor list
constexpr int STG_LD = 36;               // dense-store variant: floats per staged row (32 + 4: conflict-free float4 access)
constexpr int STG_SLOTS = 18;            // ... its staging area, in units of a survivor slot row (2 KB): 8 warps x 32 rows x 144 B

struct SharedTail {  // lives after the operand ring and the survivor lists
  float hx[2][2][BN];              // [group][ping-pong] -0.5*|x_j|^2 of the current column tile
  int cnt_x[2][NUM_EPI];           // survivor-count exchange between the two threads of a row
  int sum_x[2][NUM_EPI];           // sum of the score-only ranks of a thread's entries (tie detection)
  float xmax_w[2 * NUM_EPI / 32];  // per-warp max |x_j|^2 over the candidates it staged
  uint64_t a_full, b_full[8], b_empty[8], t_full[2], t_empty[2];
  uint32_t tmem_slot;
  float cmax_s;                    // FOLD: the cloud's max |s x_j|^2
};

// shared memory: [S ring stages of 32 KB][cap survivor slots x 256 threads x 8 bytes][tail]
__host__ __device__ constexpr size_t smem_bytes(int S, int cap) {
  return 1024 /* alignment slack */ + (size_t)S * STAGE_BYTES + (size_t)cap * LS * sizeof(uint64_t) +
         sizeof(SharedTail);
}

// Optional timeline (diagnostics): CTA (0,0) stamps clock64() into tl[role*256 + i]
#define ECB_STAMP(role, i)                                                                  \
  do {                                                                                      \
    if (tl && blockIdx.x == 0 && blockIdx.y == 0 && (i) < 256) tl[(role) * 256 + (i)] = clock64(); \
  } while (0)

// DEBUG = true: one sweep, raw scores written to dbg[B,N,N] (validation of the MMA plumbing; also
//               the dense-store epilogue of the per-point GEMM)
// CL:           CTAs per cluster.  The CL row tiles of a cluster belong to one cloud and stream the
//               same candidate tiles: each CTA fetches 1/CL of every tile and TMA multicasts it to
//               all of them, so the L2 -> SM traffic of the candidates drops by CL.
// S:            ring stages of 32 KB.
// F16:          the operands are packed FP16 pairs (two channels per 32-bit word; values scaled by a
//               power of two into FP16's range by ecb200_split_f16) and the MMAs are kind::f16: the
//               same 11-bit significands as tf32 at twice the rate.  In words nothing else changes:
//               `nkb` counts blocks of 32 WORDS (64 channels) then.
// TERMS:        3 = error-compensated product hi.hi + hi.lo + lo.hi (pass A ranks with hi.hi alone);
//               1 = the hi arrays already carry the whole product (xyz layer: the three terms of a
//               3-channel point sit side by side in ONE 16-deep K step), both passes issue the same
//               single MMA and no margin is needed.
// ksteps:       K steps per 32-word block that hold data (4; 1 for the packed xyz operands).
// FOLD:         (packed FP16 only) the column term -0.5*|x_j|^2 is part of the contraction: three more K
//               slots hold 2^15 on the query side and the three FP16 pieces of -|s x_j|^2 / 2^16 on the
//               candidate side (the operands are scaled into [2^11, 2^12) so that the term fits FP16's
//               range), so the epilogue neither adds nor stages anything per candidate.  With TERMS = 3
//               the slots are one extra K step whose candidate rows come from `map_bn`; with TERMS = 1
//               they sit in free slots of the one K step.  cmax_g[b] = max_j |s x_j|^2 of cloud b.
template <int NBINS, bool DEBUG, int CL, int S, bool F16 = false, int TERMS = 3, bool FOLD = false>
__global__ void __launch_bounds__(NT, 1)
knn_tc_kernel(const float* __restrict__ a_hi_g, const float* __restrict__ a_lo_g,
              const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
              const __grid_constant__ CUtensorMap map_bn, const float* __restrict__ xx,
              const float* __restrict__ cmax_g, int Na, int N, int nkb, int ksteps, int k, int cap,
              int32_t* __restrict__ idx, float* __restrict__ dbg, long long* tl) {
  // A operand: rows [b*Na + rt*128, +128) of a_hi_g / a_lo_g [*, C] (the queries; for the GEMM use
  // the points), copied ONCE into tensor memory (lane = row, column = channel): the MMAs then
  // read only the B operand from shared memory, which halves their shared-memory traffic.
  // B operand: rows [b*N + ct*128, +128) of map_bhi/blo (the candidates; for the GEMM the rows of
  // Wcat), streamed through a TMA / mbarrier ring.  For kNN A and B are the same arrays, Na == N.
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment by pointer arithmetic on the __shared__ array (an integer round-trip
  // would turn every later access into a generic-address load/store)
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int C = nkb * KB;
  unsigned char* b_st = base;                                  // [S][32 KB]
  uint64_t* surv = reinterpret_cast<uint64_t*>(b_st + (size_t)S * STAGE_BYTES);   // [cap][256]
  SharedTail* T = reinterpret_cast<SharedTail*>(reinterpret_cast<unsigned char*>(surv) +
                                                (size_t)(DEBUG ? STG_SLOTS : cap) * LS * sizeof(uint64_t));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, rt = blockIdx.x;
  const int nct = (N + BN - 1) / BN;
  const int cloud_row0 = b * N;   // first global B row of this cloud
  const int a_row0 = b * Na;      // first global A row of this cloud
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
  // pass A multiplies the hi halves only: a stage then carries the hi halves of TWO K-blocks
  const int kpa = (nkb & 1) ? 1 : 2;

  // epilogue threads fetch their query row (first 64 channels) before anything else: the loads
  // are in flight while the barriers are initialised and tensor memory is allocated
  float4 pre[16], pre2[16];   // words 0..63 and 64..127 of the row (the second half only when C > 64)
  if (warp >= 2) {
    const int g = (warp - 2) >> 2, q = warp & 3;
    const int row = rt * BM + q * 32 + lane;
    const float* src = (g == 0 ? a_hi_g : a_lo_g) + (size_t)(a_row0 + row) * C;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      pre[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < Na && 4 * e < C) pre[e] = __ldg(reinterpret_cast<const float4*>(src) + e);
    }
    if (C > 64) {
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        pre2[e] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < Na && 64 + 4 * e < C) pre2[e] = __ldg(reinterpret_cast<const float4*>(src) + 16 + e);
      }
    }
  }

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_bhi);
    prefetch_tensormap(&map_blo);
    mbar_init(&T->a_full, 2 * NUM_EPI);
    for (int s = 0; s < S; ++s) { mbar_init(&T->b_full[s], 1); mbar_init(&T->b_empty[s], CL); }
    for (int s = 0; s < 2; ++s) { mbar_init(&T->t_full[s], 1); mbar_init(&T->t_empty[s], NUM_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&T->tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // the peers' barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = T->tmem_slot;
  if (threadIdx.x == 0) ECB_STAMP(5, 0);
  if (tl && threadIdx.x == 0) {  // diagnostics: wall-clock span and SM of every CTA
    unsigned long long t; unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    long long* e = tl + 6 * 256 + 3 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
    e[0] = (long long)t; e[2] = sm;
  }

  if (warp == 0) {
    // ===================== TMA producer (whole warp in the loop, one elected lane issues) =====
    {
      int stage = 0, n = 0;
      uint32_t phase = 0;
      constexpr int SLICE_ROWS = BM / CL;                 // rows of every tile this CTA fetches
      const uint32_t slice_off = crank * (uint32_t)(SLICE_ROWS * KB * 4);
      auto load = [&](unsigned char* dst, const CUtensorMap* m, uint64_t* bar, int c0, int r0) {
        if (FOLD && !DEBUG) tma_load_3d(dst, m, bar, c0, r0 - cloud_row0, b);   // per-cloud map: rows past N are NaN
        else if (CL > 1)    tma_load_2d_mc(dst + slice_off, m, bar, c0, r0 + (int)crank * SLICE_ROWS, CMASK);
        else                tma_load_2d(dst, m, bar, c0, r0);
      };
      if (!DEBUG && FOLD && TERMS == 3) {
        // first sweep: nkb hi K-blocks + the norm block of every tile, two blocks per stage
        for (int ct = 0; ct < nct; ++ct)
          for (int u0 = 0; u0 <= nkb; u0 += 2, ++n) {
            const int nun = min(2, nkb + 1 - u0);
            mbar_wait(&T->b_empty[stage], phase ^ 1);
            if (elect_one_sync()) {
              ECB_STAMP(0, n);
              unsigned char* dst = b_st + (size_t)stage * STAGE_BYTES;
              mbar_expect_tx(&T->b_full[stage], nun * TILE_BYTES);
              for (int j = 0; j < nun; ++j) {
                if (u0 + j < nkb) load(dst + j * TILE_BYTES, &map_bhi, &T->b_full[stage], (u0 + j) * KB, cloud_row0 + ct * BN);
                else              load(dst + j * TILE_BYTES, &map_bn, &T->b_full[stage], 0, cloud_row0 + ct * BN);
              }
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
      } else if (!DEBUG) {
        for (int sweep = 0; sweep < (TERMS == 1 ? 2 : 1); ++sweep)
        for (int ct = 0; ct < nct; ++ct)
          for (int kb = 0; kb < nkb; kb += kpa, ++n) {
            mbar_wait(&T->b_empty[stage], phase ^ 1);
            if (elect_one_sync()) {
              ECB_STAMP(0, n);
              unsigned char* dst = b_st + (size_t)stage * STAGE_BYTES;
              mbar_expect_tx(&T->b_full[stage], kpa * TILE_BYTES);
              for (int j = 0; j < kpa; ++j)
                load(dst + j * TILE_BYTES, &map_bhi, &T->b_full[stage], (kb + j) * KB, cloud_row0 + ct * BN);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
      }
      if (TERMS == 3)
      for (int ct = 0; ct < nct; ++ct)
        for (int kb = 0; kb < nkb + ((FOLD && !DEBUG) ? 1 : 0); ++kb, ++n) {
          mbar_wait(&T->b_empty[stage], phase ^ 1);
          if (elect_one_sync()) {
            ECB_STAMP(0, n);
            unsigned char* dst = b_st + (size_t)stage * STAGE_BYTES;
            if (kb < nkb) {
              mbar_expect_tx(&T->b_full[stage], 2 * TILE_BYTES);
              load(dst, &map_bhi, &T->b_full[stage], kb * KB, cloud_row0 + ct * BN);
              load(dst + TILE_BYTES, &map_blo, &T->b_full[stage], kb * KB, cloud_row0 + ct * BN);
            } else {   // the norm block of the tile, a stage of its own
              mbar_expect_tx(&T->b_full[stage], TILE_BYTES);
              load(dst, &map_bn, &T->b_full[stage], 0, cloud_row0 + ct * BN);
            }
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp in the loop, one elected lane issues) ========
    {
      const uint32_t idesc = F16 ? make_idesc_f16(BM, BN) : make_idesc_tf32(BM, BN);
      auto mma = [&](uint32_t d, uint32_t a, uint32_t bdesc, uint32_t acc) {
        if (F16) mma_f16_ts_lo(d, a, bdesc, idesc, acc);
        else     mma_tf32_ts_lo(d, a, bdesc, idesc, acc);
      };
      const uint32_t tmem_base = __shfl_sync(0xffffffffu, T->tmem_slot, 0);   // warp-uniform by construction
      mbar_wait(&T->a_full, 0);
      tc_fence_after();
      int stage = 0;
      uint32_t phase = 0;
      int tile = 0;
      const uint32_t ring_lo = sw128_kmajor_desc_lo(smem_u32(b_st));  // descriptor low word of stage 0, first half
      constexpr uint32_t STAGE_STEP = STAGE_BYTES >> 4, LO_STEP = TILE_BYTES >> 4, K8_STEP = (UMMA_K * 4) >> 4;
      const uint32_t a_col = tmem_base + A_COL0;
      auto release = [&](int st) {  // frees the stage (in every CTA of the cluster) once these MMAs have read it
        if (CL > 1) mma_commit_mc(&T->b_empty[st], CMASK);
        else        mma_commit(&T->b_empty[st]);
      };
      const uint32_t a_norm = a_col + (uint32_t)(2 * C);   // FOLD: the 2^15 constants behind the hi | lo halves
      if (!DEBUG && FOLD && TERMS == 3) {
        // first sweep with the norm block: units 0..nkb-1 = hi K-blocks, unit nkb = norm, two per stage
        for (int ct = 0; ct < nct; ++ct, ++tile) {
          const int as = tile & 1;
          mbar_wait(&T->t_empty[as], ((tile >> 1) & 1) ^ 1);
          tc_fence_after();
          ECB_STAMP(1, 2 * tile);
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int u0 = 0; u0 <= nkb; u0 += 2) {
            const int nun = min(2, nkb + 1 - u0);
            mbar_wait(&T->b_full[stage], phase);
            tc_fence_after();
            const uint32_t bh = ring_lo + (uint32_t)stage * STAGE_STEP;
            if (elect_one_sync()) {
              for (int j = 0; j < nun; ++j) {
                if (u0 + j < nkb) {
                  const uint32_t ah = a_col + (uint32_t)((u0 + j) * KB);
#pragma unroll
                  for (int k8 = 0; k8 < KB / UMMA_K; ++k8)
                    mma(d_tmem, ah + k8 * UMMA_K, bh + j * LO_STEP + k8 * K8_STEP, (u0 | j | k8) != 0);
                } else {
                  mma(d_tmem, a_norm, bh + j * LO_STEP, 1);
                }
              }
              release(stage);
              if (u0 + 2 > nkb) mma_commit(&T->t_full[as]);
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          ECB_STAMP(1, 2 * tile + 1);
        }
      } else if (!DEBUG) {
        // pass A: one product of the hi halves (ranked with an error margin), two K-blocks per stage;
        // with TERMS == 1 the second sweep is the same again
        for (int sweep = 0; sweep < (TERMS == 1 ? 2 : 1); ++sweep)
        for (int ct = 0; ct < nct; ++ct, ++tile) {
          const int as = tile & 1;  // stage g is consumed by epilogue warpgroup g
          mbar_wait(&T->t_empty[as], ((tile >> 1) & 1) ^ 1);  // that group drained this stage
          tc_fence_after();
          ECB_STAMP(1, 2 * tile);
          const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
          for (int kb = 0; kb < nkb; kb += kpa) {
            mbar_wait(&T->b_full[stage], phase);
            tc_fence_after();
            const uint32_t bh = ring_lo + (uint32_t)stage * STAGE_STEP;
            if (elect_one_sync()) {
              for (int j = 0; j < kpa; ++j) {
                const uint32_t ah = a_col + (uint32_t)((kb + j) * KB);
#pragma unroll
                for (int k8 = 0; k8 < KB / UMMA_K; ++k8)
                  if (k8 < ksteps)
                    mma(d_tmem, ah + k8 * UMMA_K, bh + j * LO_STEP + k8 * K8_STEP, (kb | j | k8) != 0);
              }
              release(stage);
              if (kb + kpa >= nkb) mma_commit(&T->t_full[as]);  // accumulator ready for the epilogue
            }
            __syncwarp();
            if (++stage == S) { stage = 0; phase ^= 1; }
          }
          ECB_STAMP(1, 2 * tile + 1);
        }
      }
      // pass B (and the only sweep of the dense-store variant): three terms, hi.hi + hi.lo + lo.hi
      if (TERMS == 3)
      for (int ct = 0; ct < nct; ++ct, ++tile) {
        const int as = tile & 1;
        mbar_wait(&T->t_empty[as], ((tile >> 1) & 1) ^ 1);
        tc_fence_after();
        ECB_STAMP(1, 2 * tile);
        const uint32_t d_tmem = tmem_base + (uint32_t)(as * BN);
        const int nst = nkb + ((FOLD && !DEBUG) ? 1 : 0);   // + the norm block's stage
        for (int kb = 0; kb < nst; ++kb) {
          mbar_wait(&T->b_full[stage], phase);
          tc_fence_after();
          const uint32_t ah = a_col + (uint32_t)(kb * KB);  // tensor-memory columns of A hi; lo at +C
          const uint32_t bh = ring_lo + (uint32_t)stage * STAGE_STEP;
          if (elect_one_sync()) {
            if (kb < nkb) {
#pragma unroll
              for (int k8 = 0; k8 < KB / UMMA_K; ++k8) {
                mma(d_tmem, ah + k8 * UMMA_K, bh + k8 * K8_STEP, (kb | k8) != 0);
                mma(d_tmem, ah + k8 * UMMA_K, bh + LO_STEP + k8 * K8_STEP, 1);
                mma(d_tmem, ah + (uint32_t)C + k8 * UMMA_K, bh + k8 * K8_STEP, 1);
              }
            } else {
              mma(d_tmem, a_norm, bh, 1);
            }
            release(stage);
            if (kb + 1 == nst) mma_commit(&T->t_full[as]);
          }
          __syncwarp();
          if (++stage == S) { stage = 0; phase ^= 1; }
        }
        ECB_STAMP(1, 2 * tile + 1);
      }
    }
  } else {
    // ===================== epilogue: selection (thread = query row) =====================
    // Two warpgroups: group g owns the tiles that land in accumulator stage g, so while one
    // group selects from a tile the other already works on the next one.  A row is
    // therefore served by two threads (one per group), each with its own bins / survivors.
    const int g = (warp - 2) >> 2;
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = (threadIdx.x - 64) & (NUM_EPI - 1);
    const int row = rt * BM + q * 32 + lane;
    const bool valid = row < Na;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN);
    {
      // query rows -> tensor memory: group 0 copies the hi halves, group 1 the lo halves
      const float* src = (g == 0 ? a_hi_g : a_lo_g) + (size_t)(a_row0 + row) * C;
      const uint32_t dst = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + (uint32_t)(g * C);
#pragma unroll
      for (int h = 0; h < 2; ++h) {       // the prefetched first 64 channels
        if (h * 32 < C) {
          uint32_t r[32];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float4 f = pre[8 * h + e];
            r[4 * e + 0] = __float_as_uint(f.x); r[4 * e + 1] = __float_as_uint(f.y);
            r[4 * e + 2] = __float_as_uint(f.z); r[4 * e + 3] = __float_as_uint(f.w);
          }
          __syncwarp();
          tmem_st_32x32(dst + (uint32_t)(h * 32), r);
        }
      }
      if (C > 64) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {       // the prefetched second 64 words
          if (64 + h * 32 < C) {
            uint32_t r[32];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 f = pre2[8 * h + e];
              r[4 * e + 0] = __float_as_uint(f.x); r[4 * e + 1] = __float_as_uint(f.y);
              r[4 * e + 2] = __float_as_uint(f.z); r[4 * e + 3] = __float_as_uint(f.w);
            }
            __syncwarp();
            tmem_st_32x32(dst + (uint32_t)(64 + h * 32), r);
          }
        }
      }
      if (FOLD && TERMS == 3 && !DEBUG && g == 0) {
        // the query side of the norm block: 2^15 (0x7800) in its first three K slots, the same for every row
        uint32_t r[8] = {0x78007800u, 0x00007800u, 0u, 0u, 0u, 0u, 0u, 0u};
        __syncwarp();
        tmem_st_32x8(tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + (uint32_t)(2 * C), r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&T->a_full);
    }
    if (FOLD && !DEBUG && warp == 2) {
      // the cloud's largest squared norm from the per-block maxima of the operand kernel (read by every
      // epilogue thread after the bar.sync of the threshold stage)
      const int nblk = (N + 31) / 32;
      float m = 0.f;
      for (int i = lane; i < nblk; i += 32) m = fmaxf(m, __ldg(cmax_g + (size_t)b * nblk + i));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) T->cmax_s = m;
    }
    float bin[NBINS];
#pragma unroll
    for (int u = 0; u < NBINS; ++u) bin[u] = -CUDART_INF_F;
    // this thread's survivor list: column `me` of surv[cap][256]; consecutive lanes touch
    // consecutive 8-byte words, so the (predicated) appends are conflict-free
    const int me = g * NUM_EPI + et;
    uint64_t* const sv = surv + me;
    const uint32_t sv_addr = smem_u32(sv);
    int cnt = 0;                  // entries in the list
    constexpr uint32_t SLOT = LS * sizeof(uint64_t);   // bytes between consecutive slots
    // |x_i|^2 for the pass-A margin: fetched now, needed after the first sweep
    const float xi = (!DEBUG && xx && valid) ? __ldg(xx + cloud_row0 + row) : 0.f;
    float thr = CUDART_INF_F;
    float xmax = 0.f;  // max |x_j|^2 over the candidates this thread staged (for the pass-A margin)
    int use = 0;  // how many times this group has consumed its accumulator stage
    // |x_j|^2 of column `et` of candidate tile t; -1 marks a column past the end of the cloud
    auto load_xx = [&](int t) -> float {
      const int j = t * BN + et;
      return (xx && j < N) ? __ldg(xx + cloud_row0 + j) : (xx ? -1.f : 0.f);
    };
    int xnext_t = g & 1;
    float xnext = (FOLD && !DEBUG) ? 0.f : load_xx(xnext_t);
    // the tile loop of one pass; `pass` is a compile-time constant in each instantiation, so the
    // two passes get separate code (and register allocations: the bins die after the first)
    auto run_tiles = [&](auto pass_tag) {
      constexpr int pass = decltype(pass_tag)::value;
      const int tile0 = DEBUG ? 0 : pass * nct;
      for (int ct = (tile0 + g) & 1; ct < nct; ct += 2, ++use) {
        float* hx = T->hx[g][use & 1];
        if (et == 0) ECB_STAMP(2 + g, 4 * use);
        if (!FOLD || DEBUG) {
          // |x_j|^2 of this tile's column was fetched one tile ahead (an L2 round trip off the
          // critical path); now fetch the next tile's (or the first tile's of the next pass)
          if (xnext_t != ct) xnext = load_xx(ct);  // only when a pass has no tile for this group
          const float xj = xnext;
          int nct_t = ct + 2;
          if (nct_t >= nct) nct_t = ((pass + 1) * nct + g) & 1;
          xnext = load_xx(nct_t);
          xnext_t = nct_t;
          xmax = fmaxf(xmax, xj);
          hx[et] = xx ? ((xj >= 0.f) ? -0.5f * xj : -CUDART_INF_F) : 0.f;
        }
        if (!FOLD || DEBUG) {
          if (g == 0) asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI) : "memory");
          else        asm volatile("bar.sync 2, %0;" ::"n"(NUM_EPI) : "memory");
        }
        if (et == 0) ECB_STAMP(2 + g, 4 * use + 1);
        mbar_wait(&T->t_full[g], use & 1);
        tc_fence_after();
        if (et == 0) ECB_STAMP(2 + g, 4 * use + 2);
        // one 32-column chunk of this row: add the column term, then store / bin / select
        auto process = [&](const uint32_t(&cur)[32], const int c4, const bool second_pass) {
          const float4* hx4 = reinterpret_cast<const float4*>(hx + c4 * 32);
          float v[32];
          if (FOLD && !DEBUG) {
            // the accumulator already is the score; candidates past the end of the cloud were loaded as NaN
            // (per-cloud tensor map with NaN fill) and score NaN: fmaxf drops them, >= rejects them
#pragma unroll
            for (int u = 0; u < 32; ++u) v[u] = __uint_as_float(cur[u]);
          } else {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float4 h4 = hx4[e];
              v[4 * e + 0] = __uint_as_float(cur[4 * e + 0]) + h4.x;
              v[4 * e + 1] = __uint_as_float(cur[4 * e + 1]) + h4.y;
              v[4 * e + 2] = __uint_as_float(cur[4 * e + 2]) + h4.z;
              v[4 * e + 3] = __uint_as_float(cur[4 * e + 3]) + h4.w;
            }
          }
          if (DEBUG) {  // dense store of the tile: raw scores, or the GEMM result Y = A.B^T
            if ((N & 3) == 0) {
              // A thread holds 32 consecutive columns of ITS row: stored directly, one instruction would
              // touch 32 rows with 16 bytes each.  The warp's 32 x 32 block goes through a private
              // staging tile instead and leaves as full 128-byte lines, four rows per instruction.
              float* stg = reinterpret_cast<float*>(surv) + (size_t)(warp - 2) * 32 * STG_LD;
              __syncwarp();
#pragma unroll
              for (int e = 0; e < 8; ++e)
                *reinterpret_cast<float4*>(stg + lane * STG_LD + 4 * e) =
                    make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
              __syncwarp();
              const int col = ct * BN + c4 * 32 + 4 * (lane & 7);
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = 4 * it + (lane >> 3);                  // row of the warp's block
                const int grow = rt * BM + q * 32 + r;
                const float4 w = *reinterpret_cast<const float4*>(stg + r * STG_LD + 4 * (lane & 7));
                if (grow < Na && col < N)
                  *reinterpret_cast<float4*>(dbg + ((size_t)(a_row0 + grow)) * N + col) = w;
              }
            } else if (valid) {
              float* orow = dbg + ((size_t)(a_row0 + row)) * N + ct * BN + c4 * 32;
              {
#pragma unroll
                for (int u = 0; u < 32; ++u)
                  if (ct * BN + c4 * 32 + u < N) orow[u] = v[u];
              }
            }
          } else if (!second_pass) {
#pragma unroll
            for (int u = 0; u < 32; ++u) {
              const int bi = (NBINS == 64 ? (c4 & 1) * 32 : 0) + u;
              bin[bi] = fmaxf(bin[bi], v[u]);
            }
          } else {
            // Candidates that reach the row's threshold are appended to the thread's own list in
            // shared memory: straight-line predicated code per candidate (compare, store, pointer
            // bump), no votes, no branches, no dependence between rows.  Rows past the end of the
            // cloud carry thr = +inf; masked columns score -inf.
            const int jb = ct * BN + c4 * 32;
#pragma unroll
            for (int h = 0; h < 32 / GUARD; ++h) {
              if (cnt > cap - GUARD) {  // rare (ties, clustered data): keep the thread's own best k
                thr = fmaxf(thr, shrink_survivors(sv, cnt, k, LS));
                cnt = k;
              }
#pragma unroll
              for (int u = h * GUARD; u < (h + 1) * GUARD; ++u) {
                if (v[u] >= thr) {   // entry = (score bits << 32) | j
                  // the slot address is formed in a scratch register per store: a running address
                  // register would be rewritten right behind every (predicated) store that reads
                  // it, and that write-after-read on a shared-memory store costs ~25 cycles
                  asm volatile(
                      "{\n\t"
                      ".reg .u32 t;\n\t"
                      "mad.lo.u32 t, %0, %1, %2;\n\t"
                      "st.shared.v2.b32 [t], {%3, %4};\n\t"
                      "}" ::"r"(cnt), "n"(SLOT), "r"(sv_addr), "r"(jb + u), "r"(__float_as_uint(v[u]))
                      : "memory");
                  ++cnt;
                }
              }
            }
          }
        };
        __syncwarp();  // tcgen05.ld / wait are warp-collective (.sync.aligned)
        if (DEBUG || pass == 0) {
          // 32 columns at a time, double-buffered: the tcgen05.ld of the next 32 columns is in
          // flight while the current ones are processed (the bins keep 32 registers busy)
          uint32_t ra[32], rb[32];
          tmem_ld_32x32_issue(lane_base, ra);
#pragma unroll
          for (int c4 = 0; c4 < BN / 32; ++c4) {
            uint32_t(&cur)[32] = (c4 & 1) ? rb : ra;
            uint32_t(&nxt)[32] = (c4 & 1) ? ra : rb;
            tmem_ld_wait(cur);
            if (c4 + 1 < BN / 32) {
              tmem_ld_32x32_issue(lane_base + (uint32_t)((c4 + 1) * 32), nxt);
            } else {
              tc_fence_before();
              mbar_arrive(&T->t_empty[g]);  // the whole stage has been read
            }
            process(cur, c4, false);
          }
        } else {
          // Second pass: the accumulator row moves into registers as early as the register
          // budget allows (96 columns, the last 32 as soon as the first chunk is done) and the
          // stage goes back to the MMA issuer, so the (three times longer) MMAs of this
          // group's next tile overlap most of the selection instead of waiting for it.
          uint32_t r0[32], r1[32], r2[32];
          tmem_ld_32x32_issue(lane_base, r0);
          tmem_ld_32x32_issue(lane_base + 32u, r1);
          tmem_ld_32x32_issue(lane_base + 64u, r2);
          tmem_ld_wait(r0);
          tmem_ld_wait(r1);
          tmem_ld_wait(r2);
          process(r0, 0, true);
          __syncwarp();
          tmem_ld_32x32_issue(lane_base + 96u, r0);
          tmem_ld_wait(r0);
          tc_fence_before();
          mbar_arrive(&T->t_empty[g]);
          process(r1, 1, true);
          process(r2, 2, true);
          process(r0, 3, true);
        }
        if (et == 0) ECB_STAMP(2 + g, 4 * use + 3);
      }
      if (et == 0) ECB_STAMP(4, 8 * g + pass);
    };
    run_tiles(std::integral_constant<int, 0>{});
    if (!DEBUG) {
      if (et == 0) ECB_STAMP(4, 8 * g + 5);
      sort_bins_desc<NBINS>(bin);
      if (et == 0) ECB_STAMP(4, 8 * g + 6);
      // The row's two threads pool their bins: the k-th largest of the union of two
      // descending lists is  max_i min(mine[i-1], theirs[k-i-1])  (i taken from mine).  The
      // partner's list travels through the (still empty) survivor area of shared memory.
      // exchange area [k][256] in the (still empty) survivor area, written REVERSED: slot i holds
      // the (k-i)-th largest bin, so the reader walks it with static offsets
      float* exch = reinterpret_cast<float*>(surv);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
      if (lane == 0) T->xmax_w[me >> 5] = xmax;  // read after the bar.sync below
      {
        float* wr = exch + (size_t)(k - 1) * LS + me;     // slot k-1-u  <-  bin[u]
#pragma unroll
        for (int u = 0; u < NBINS; ++u)
          if (u < k) wr[-u * LS] = bin[u];
        for (int i = 0; i < k - NBINS; ++i) exch[i * LS + me] = -CUDART_INF_F;   // k > 32: ranks that do not exist
      }
      asm volatile("bar.sync 3, %0;" ::"n"(2 * NUM_EPI) : "memory");
      const float* other = exch + ((g ^ 1) * NUM_EPI + et);
      float tau = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i <= KMAX; ++i) {
        // i entries from mine, k - i from theirs: min(mine[i-1], theirs[k-i-1])
        if (i <= k) {
          const float mine = i == 0 ? CUDART_INF_F : (i - 1 < NBINS ? bin[i - 1 < NBINS ? i - 1 : 0] : -CUDART_INF_F);
          const float theirs = i == k ? CUDART_INF_F : other[(i < KMAX ? i : 0) * LS];
          tau = fmaxf(tau, fminf(mine, theirs));
        }
      }
      // Pass A scored with the hi.hi product only: |score_A - score_B| <= 2^-11 (|hi_i||x_j| +
      // |x_i||hi_j|) <= 1.1 * 2^-10 |x_i| max_j|x_j|.  Lowering the bound by that margin keeps
      // it a lower bound of the row's k-th best 3xTF32 score (a few more survivors, no misses).
      float cmax = 0.f;
      if (FOLD) {
        cmax = T->cmax_s;
      } else {
#pragma unroll
        for (int w = 0; w < 2 * NUM_EPI / 32; ++w) cmax = fmaxf(cmax, T->xmax_w[w]);
      }
      // FP16 halves below 2^-14 are subnormal: absolute error 2^-25 per element on top of the
      // relative one, i.e. at most 2^-24 sqrt(channels) (|x_i| + |x_j|) on the score
      float margin = 1.1f * 0.0009765625f * sqrtf(xi * cmax) + 1e-30f;
      if (F16) margin += 5.96e-8f * sqrtf((float)(2 * C)) * (sqrtf(xi) + sqrtf(cmax));
      if (TERMS == 1) margin = 0.f;   // both sweeps compute the very same scores
      // masked candidates score -inf and must never pass; rows past the end keep nothing
      thr = valid ? fmaxf(tau - margin, -3.0e38f) : CUDART_INF_F;
      asm volatile("bar.sync 3, %0;" ::"n"(2 * NUM_EPI) : "memory");  // bins read: the area may take survivors
      if (et == 0) ECB_STAMP(4, 8 * g + 7);
      run_tiles(std::integral_constant<int, 1>{});
      if (!valid) cnt = 0;
      {
        // the union of a row's two lists is ranked in the (then dead) operand ring, one KB per slot:
        // a list longer than half of the ring's slots is cut to its own best k first (exact: at
        // most k entries of a list can make the row's top k).  Never taken with the 4-stage ring.
        constexpr int UNI_HALF = S * STAGE_BYTES / 1024 / 2;
        if (cnt > UNI_HALF) {
          shrink_survivors(sv, cnt, k, LS);
          cnt = k;
        }
      }
      // Top-k of the row's survivors (both groups), nearest first: rank = number of better
      // entries = output slot.  Fast path: every thread ranks its OWN entries against the union
      // by SCORE only (one compare on the ALU pipe + one add on the FMA pipe per pair).  Without
      // equal scores the ranks of a row are a permutation, i.e. they sum to total*(total-1)/2;
      // the two threads of a row compare that sum and, only if it falls short (equal scores:
      // duplicate points, lattices), redo the row exactly under the total order (score, then
      // smaller j).  Slots are staged in the (dead) operand ring and written out coalesced.
      T->cnt_x[g][et] = cnt;
      asm volatile("bar.sync 3, %0;" ::"n"(2 * NUM_EPI) : "memory");  // both lists complete; every MMA done
      if (et == 0) ECB_STAMP(4, 8 * g + 3);
      // the operand ring is dead now (every MMA has completed): it takes the row's UNION of both
      // lists, slot-major [slot][128 rows] (conflict-free, and a plain strided walk for the
      // ranking loop); at most 2k slots = 2k KB <= the ring's 96 KB
      uint64_t* uni = reinterpret_cast<uint64_t*>(b_st) + et;           // slot f of this row at uni[f * NUM_EPI]
      const int cnt0 = T->cnt_x[0][et], cnt1 = T->cnt_x[1][et];
      const int total = cnt0 + cnt1;
      // staged output [128 rows][k]: in the survivor area, which is dead once the lists have moved
      int32_t* out_s = reinterpret_cast<int32_t*>(surv);
      int32_t* orow = out_s + (q * 32 + lane) * k;
      {
        uint64_t* dstc = uni + (size_t)(g == 0 ? 0 : cnt0) * NUM_EPI;
        for (int e0 = 0; e0 < cnt; e0 += 8) {   // loads first, then stores: one shared-memory round trip per 8
          uint64_t w[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) w[u] = e0 + u < cnt ? sv[(e0 + u) * LS] : 0ull;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (e0 + u < cnt) dstc[(e0 + u) * NUM_EPI] = w[u];
        }
      }
      asm volatile("bar.sync 3, %0;" ::"n"(2 * NUM_EPI) : "memory");
      if (et == 0) ECB_STAMP(4, 8 * g + 2);
      if (tl && blockIdx.x == 0 && blockIdx.y == 0 && g == 0) tl[5 * 256 + 128 + et] = ((long long)cnt0 << 32) | (unsigned)cnt1;
      // the row's two threads split the union evenly (balanced work whatever the individual list
      // lengths): mine is [lo, hi)
      const int half = (total + 1) >> 1;
      const int lo = g * half, hi = min(total, lo + half);
      constexpr int CH = 9;   // own entries per sweep of the union: two sweeps cover total <= 36
      int ranksum = 0;
      for (int e0 = lo; e0 < hi; e0 += CH) {
        float so[CH], rf[CH];
        uint32_t jo[CH];
#pragma unroll
        for (int u = 0; u < CH; ++u) {
          const uint64_t w = e0 + u < hi ? uni[(e0 + u) * NUM_EPI] : 0ull;
          so[u] = e0 + u < hi ? __uint_as_float((uint32_t)(w >> 32)) : CUDART_INF_F;
          jo[u] = (uint32_t)w;
          rf[u] = 0.f;
        }
        const float* sc = reinterpret_cast<const float*>(uni) + 1;   // score word of slot f at sc[f * 2 * NUM_EPI]
#pragma unroll 4
        for (int f = 0; f < total; ++f) {
          const float sf = sc[f * 2 * NUM_EPI];
          // compare on the ALU pipe (FSETP), predicated add on the FMA pipe; written out because the
          // compiler's FSET.BF form measured ~4x slower here
#pragma unroll
          for (int u = 0; u < CH; ++u)
            asm("{\n\t"
                ".reg .pred p;\n\t"
                "setp.gt.f32 p, %1, %2;\n\t"
                "@p add.f32 %0, %0, 0f3F800000;\n\t"
                "}"
                : "+f"(rf[u])
                : "f"(sf), "f"(so[u]));
        }
#pragma unroll
        for (int u = 0; u < CH; ++u)
          if (e0 + u < hi) {
            const int r = (int)rf[u];
            ranksum += r;
            if (r < k) orow[r] = (int32_t)min(jo[u], (uint32_t)(N - 1));
          }
      }
      T->sum_x[g][et] = ranksum;
      if (et == 0) ECB_STAMP(5, 8 + 4 * g);
      asm volatile("bar.sync 3, %0;" ::"n"(2 * NUM_EPI) : "memory");
      if (et == 0) ECB_STAMP(5, 9 + 4 * g);
      if (valid && T->sum_x[0][et] + T->sum_x[1][et] != total * (total - 1) / 2) {
        // equal scores in this row: exact ranks under the total order
        for (int e0 = lo; e0 < hi; e0 += 4) {
          uint64_t own[4];
          int rank[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            own[u] = e0 + u < hi ? ordered_key(uni[(e0 + u) * NUM_EPI]) : ~0ull;
            rank[u] = 0;
          }
#pragma unroll 1
          for (int f = 0; f < total; ++f) {
            const uint64_t kf = ordered_key(uni[f * NUM_EPI]);
#pragma unroll
            for (int u = 0; u < 4; ++u) rank[u] += (kf > own[u]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
            if (e0 + u < hi && rank[u] < k) orow[rank[u]] = (int32_t)min(key_index(own[u]), (uint32_t)(N - 1));
        }
      }
      if (valid && g == 0)
        for (int p = total; p < k; ++p) orow[p] = N - 1;  // only with NaN input
      asm volatile("bar.sync 3, %0;" ::"n"(2 * NUM_EPI) : "memory");
      if (et == 0) ECB_STAMP(5, 10 + 4 * g);
      {
        // the CTA's rows are consecutive in idx: one contiguous, coalesced block of nrows*k words
        const int nrows = min(BM, Na - rt * BM);
        int32_t* dst = idx + (size_t)(cloud_row0 + rt * BM) * k;
        for (int w = me; w < nrows * k; w += 2 * NUM_EPI) dst[w] = out_s[w];
      }
      if (et == 0) ECB_STAMP(4, 8 * g + 4);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
  // no CTA of a cluster may exit while a peer can still signal its barriers
  if (CL > 1) cluster_sync_all();
  if (tl && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[6 * 256 + 3 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) + 1] = (long long)t;
  }
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

// ---- packed-FP16 operands -------------------------------------------------------------------------
// max |x| over a tensor, kept as the bit pattern of a non-negative float (ordered like an unsigned)
__global__ void __launch_bounds__(256)
absmax_kernel(const float* __restrict__ x, long long n, unsigned* __restrict__ amax) {
  unsigned m = 0u;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n4 = (reinterpret_cast<uintptr_t>(x) & 15u) == 0 ? n / 4 : 0;
  for (long long q = i; q < n4; q += stride) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + q);
    m = max(max(m, __float_as_uint(fabsf(v.x))), max(__float_as_uint(fabsf(v.y)),
            max(__float_as_uint(fabsf(v.z)), __float_as_uint(fabsf(v.w)))));
  }
  for (long long q = 4 * n4 + i; q < n; q += stride) m = max(m, __float_as_uint(fabsf(x[q])));
  m = ecb200::block_max_u32(m);
  // one of ECB200_AMAX_SLOTS slots per block (no single hot address); monotone, so it is skipped
  // once the slot is already as large (a stale read only costs a redundant atomic)
  unsigned* slot = amax + (blockIdx.x & (ECB200_AMAX_SLOTS - 1));
  if (threadIdx.x == 0 && m > *reinterpret_cast<volatile unsigned*>(slot)) atomicMax(slot, m);
}

// power of two that moves the tensor's largest magnitude into [2^11, 2^12): hi = fp16(s x) is then a
// normal number with an 11-bit significand for every element above 2^-25 of the maximum,
// lo = fp16(s x - hi) keeps the pair exact to max(2^-22 |s x|, 2^-25), and |s x_j|^2 / 2^16 <= 2^15
// for up to 128 channels, i.e. the column term -0.5 |x_j|^2 fits three FP16 pieces (norm_pieces)
__device__ __forceinline__ float f16_scale(unsigned amax_bits) {
  const float a = __uint_as_float(amax_bits);
  if (!(a > 0.f) || amax_bits >= 0x7f800000u) return 1.f;   // empty / all-zero / non-finite input
  int e;
  (void)frexpf(a, &e);                                      // a = m 2^e, m in [0.5, 1)
  return ldexpf(1.f, min(12 - e, 100));   // (a tensor of denormals keeps a finite scale)
}
// -0.5 q = 2^15 * (p0 + p1 + p2) with FP16 pieces p (q = |s x_j|^2 <= 2^31): exact to 2^-33 relative
__device__ __forceinline__ void norm_pieces(float q, __half (&p)[3]) {
  float w = -q * (1.f / 65536.f);
  p[0] = __float2half_rn(w);
  w -= __half2float(p[0]);
  p[1] = __float2half_rn(w);
  w -= __half2float(p[1]);
  p[2] = __float2half_rn(w);
}

// x[B,C,N] -> point-major packed halves hh = fp16(s x), hl = fp16(s x - hh) [M,C] and xxs[M] =
// |s x|^2 (= s^2 |x|^2 exactly: scaling by a power of two commutes with rounding); optionally the
// tf32 operand pair and |x|^2 of split_tf32_kernel in the same pass over x
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, int C, int N, const unsigned* __restrict__ amax,
                 __half* __restrict__ hh, __half* __restrict__ hl, __half* __restrict__ nb,
                 float* __restrict__ xxs, float* __restrict__ cmax,
                 float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ xx) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n0 = blockIdx.x * 32, bb = blockIdx.y;
  const float* xb = x + (size_t)bb * C * N;
  __shared__ float s_sh;
  if (ty == 0) {   // the tensor's max |x| = max over the slots
    unsigned m = __ldg(amax + tx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (tx == 0) s_sh = f16_scale(m);
  }
  __syncthreads();
  const float s = s_sh;
  for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int c = c0 + ty + 8 * r, n = n0 + tx;
      tile[ty + 8 * r][tx] = (c < C && n < N) ? xb[(size_t)c * N + n] : 0.f;
    }
    __syncthreads();
    if (hi) {
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
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {       // 32 points x 16 channel pairs
      const int p = ty * 32 + tx + 256 * r;
      const int pn = p >> 4, cp = p & 15;
      const int n = n0 + pn, c = c0 + 2 * cp;
      if (n < N && c < C) {             // C is even
        const float v0 = s * tile[2 * cp][pn], v1 = s * tile[2 * cp + 1][pn];
        const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
        const size_t o = ((size_t)bb * N + n) * C + c;
        *reinterpret_cast<__half2*>(hh + o) = __halves2half2(h0, h1);
        *reinterpret_cast<__half2*>(hl + o) = __halves2half2(__float2half_rn(v0 - __half2float(h0)),
                                                             __float2half_rn(v1 - __half2float(h1)));
      }
    }
    __syncthreads();
  }
  if (ty == 0 && n0 + tx < N) {  // same summation order as ecb200_sqnorms
    float a = 0.f, b = 0.f;
    for (int c = 0; c < C; ++c) {
      const float v = xb[(size_t)c * N + n0 + tx];
      a = fmaf(v, v, a);
      const float w = s * v;
      b = fmaf(w, w, b);
    }
    if (xx) xx[(size_t)bb * N + n0 + tx] = a;
    xxs[(size_t)bb * N + n0 + tx] = b;
    // the norm block's candidate row: [p0 p1 p2 0 ... 0] in the first 32 bytes of a 128-byte row
    __align__(16) __half r[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) r[u] = __float2half_rn(0.f);
    __half pc[3];
    norm_pieces(b, pc);
    r[0] = pc[0]; r[1] = pc[1]; r[2] = pc[2];
    uint4* nd = reinterpret_cast<uint4*>(nb + ((size_t)bb * N + n0 + tx) * 64);
    nd[0] = reinterpret_cast<const uint4*>(r)[0];
    nd[1] = reinterpret_cast<const uint4*>(r)[1];
  }
  if (ty == 0) {   // the cloud's max |s x_j|^2 (margin of the first sweep)
    float b = (n0 + tx < N) ? xxs[(size_t)bb * N + n0 + tx] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (tx == 0) cmax[(size_t)bb * gridDim.x + blockIdx.x] = b;   // per-block maxima: no atomics, no zero-fill
  }
}

// xyz layer (C <= 4 channels): all three terms of the compensated product AND the column term fit ONE
// 16-deep K step,
//   query row     A = [ h(0..C-1) | h(0..C-1) | l(0..C-1) | 2^15 2^15 2^15 | 0 ... ]
//   candidate row B = [ h(0..C-1) | l(0..C-1) | h(0..C-1) |  p0   p1   p2  | 0 ... ]
//   A.B = h.h + h.l + l.h - 0.5 |s x_j|^2
// Rows are 128 bytes (one swizzle row); only their first 32 bytes are ever read by the MMAs.
// Every CTA (256 points of one cloud) first finds the cloud's own largest magnitude -- a cloud is a few KB,
// re-read by each of its CTAs from L2 -- so the scale is per cloud and no separate reduction pass is needed.
template <int C>
__global__ void __launch_bounds__(256)
pack_xyz_f16_kernel(const float* __restrict__ x, int N, __half* __restrict__ arow,
                    __half* __restrict__ brow, float* __restrict__ xxs, float* __restrict__ cmax) {
  static_assert(3 * C + 3 <= 16, "three product terms and the norm term must fit one 16-deep K step");
  __shared__ unsigned m_s;
  const int bb = blockIdx.y;
  const float* xb = x + (size_t)bb * C * N;
  unsigned m = 0u;
#pragma unroll 8
  for (int i = threadIdx.x; i < C * N; i += blockDim.x) m = max(m, __float_as_uint(fabsf(__ldg(xb + i))));
  m = ecb200::block_max_u32(m);
  if (threadIdx.x == 0) m_s = m;
  __syncthreads();
  const float s = f16_scale(m_s);
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  float q = 0.f;
  if (n < N) {
    __align__(16) __half a[16], b[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) a[u] = b[u] = __float2half_rn(0.f);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = s * xb[(size_t)c * N + n];
      q = fmaf(v, v, q);
      const __half h = __float2half_rn(v);
      const __half l = __float2half_rn(v - __half2float(h));
      a[c] = h; a[C + c] = h; a[2 * C + c] = l;
      b[c] = h; b[C + c] = l; b[2 * C + c] = h;
    }
    // the column term -0.5 |s x_j|^2 in three more slots: 2^15 on the query side, its pieces on the other
    __half pc[3];
    norm_pieces(q, pc);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      a[3 * C + u] = __ushort_as_half((unsigned short)0x7800);
      b[3 * C + u] = pc[u];
    }
    const size_t o = ((size_t)bb * N + n) * 64;
    uint4* ad = reinterpret_cast<uint4*>(arow + o);
    uint4* bd = reinterpret_cast<uint4*>(brow + o);
    ad[0] = reinterpret_cast<const uint4*>(a)[0]; ad[1] = reinterpret_cast<const uint4*>(a)[1];
    bd[0] = reinterpret_cast<const uint4*>(b)[0]; bd[1] = reinterpret_cast<const uint4*>(b)[1];
    xxs[(size_t)bb * N + n] = q;
  }
  // per-32-point maxima of |s x_j|^2, the layout ecb200_split_f16 writes
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q = fmaxf(q, __shfl_xor_sync(0xffffffffu, q, o));
  if ((threadIdx.x & 31) == 0 && n < N) cmax[(size_t)bb * ((N + 31) / 32) + (n >> 5)] = q;
}

struct TcArgs {
  const float *a_hi, *a_lo, *b_hi, *b_lo, *xx;
  long long b_rows;   // rows of the B arrays
  int clouds, C, Na, Nb, k;   // C in 32-bit words per row (= channels for tf32, channels / 2 for packed fp16)
  int32_t* idx;
  float* dbg;
  long long* tl;
  int ksteps = KB / UMMA_K;   // K steps per 32-word block that hold data
  const float* bn = nullptr;  // FOLD: norm rows [b_rows, 32 words] (TERMS = 3)
  const float* cmax = nullptr;  // FOLD: per-cloud max |s x_j|^2
};

// clouds = grid.y; per cloud Na rows of A (queries / points) and Nb rows of B (candidates / Wcat rows)
template <bool DEBUG, int CL, int S, int CAP, bool F16 = false, int TERMS = 3, bool FOLD = false>
int launch_tc(const TcArgs& a, cudaStream_t st) {
  OperandMaps Bm;
  int rc;
  if (FOLD && !DEBUG) {   // per-cloud maps: candidate rows past the end of a cloud read as NaN
    rc = make_cloud_map(&Bm.hi, a.b_hi, a.clouds, a.Nb, a.C, BM);
    if (!rc) rc = make_cloud_map(&Bm.lo, a.b_lo, a.clouds, a.Nb, a.C, BM);
  } else {
    rc = make_operand(&Bm, a.b_hi, a.b_lo, a.b_rows, a.C, BM / CL);
  }
  if (rc) return rc;
  CUtensorMap Bn = Bm.hi;   // placeholder unless the norm block is a separate operand
  if (FOLD && TERMS == 3) {
    rc = make_cloud_map(&Bn, a.bn, a.clouds, a.Nb, KB, BM);
    if (rc) return rc;
  }
  const int nkb = a.C / KB;
  auto kern = knn_tc_kernel<32, DEBUG, CL, S, F16, TERMS, FOLD>;
  constexpr size_t smem = smem_bytes(S, DEBUG ? STG_SLOTS : CAP);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if (ecb200::first_use_on_device(seen))
    ECB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ecb200::ceil_div(a.Na, BM), a.clouds);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a.a_hi, a.a_lo, Bm.hi, Bm.lo, Bn, a.xx, a.cmax, a.Na, a.Nb, nkb,
                                     a.ksteps, a.k, CAP, a.idx, a.dbg, a.tl);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ecb200::set_error("launch of knn_tc_kernel failed: %s", cudaGetErrorString(e));
    return ECB200_ERR_CUDA;
  }
  return ECB200_OK;
}

// CTAs per cluster for a cloud of `tiles` row tiles: the largest of {4, 2, 1} that divides the
// tile count and does not exceed the request (ECB200_KNN_CLUSTER, a tuning knob read once).
// Default 1: on B200 the candidate tiles are served from L2 at ~4 KB/clk chip-wide, below its limit,
// and the lock-step of a cluster costs 3-7 % (measured at N = 1024: 58.2 / 60.5 / 62.3 us for 1 / 2 / 4);
// multicast pays only when many more row tiles share a cloud than the L2 can feed.
int cluster_size(int tiles) {
  static int want = -1;
  if (want < 0) {
    const char* e = getenv("ECB200_KNN_CLUSTER");
    want = e ? atoi(e) : 1;
    if (want < 1) want = 1;
  }
  if (want >= 4 && tiles % 4 == 0) return 4;
  if (want >= 2 && tiles % 2 == 0) return 2;
  return 1;
}

// survivor slots per thread (two threads per row): k + GUARD, the overflow check runs every GUARD
// candidate columns (expected use ~12 at k = 20, ~30 at k = 40); ring depth from what is left of the
// 227 KB: k <= 20 -> 4 stages + 36 slots (204 KB), k <= 40 -> 3 stages + 56 slots (212 KB)
int launch_knn(const TcArgs& a, cudaStream_t st) {
  const int cl = cluster_size(ecb200::ceil_div(a.Na, BM));
  if (a.k <= 20) {
    if (cl == 4) return launch_tc<false, 4, 4, 20 + GUARD>(a, st);
    if (cl == 2) return launch_tc<false, 2, 4, 20 + GUARD>(a, st);
    return launch_tc<false, 1, 4, 20 + GUARD>(a, st);
  }
  if (cl == 4) return launch_tc<false, 4, 3, KMAX + GUARD>(a, st);
  if (cl == 2) return launch_tc<false, 2, 3, KMAX + GUARD>(a, st);
  return launch_tc<false, 1, 3, KMAX + GUARD>(a, st);
}

// packed-FP16 operands (ecb200_split_f16 / ecb200_pack_xyz_f16): no cluster variants
template <int TERMS>
int launch_knn_f16(const TcArgs& a, cudaStream_t st) {
  if (a.k <= 20) return launch_tc<false, 1, 4, 20 + GUARD, true, TERMS, true>(a, st);
  return launch_tc<false, 1, 3, KMAX + GUARD, true, TERMS, true>(a, st);
}

__global__ void split_rows_tf32_kernel(const float* __restrict__ src, long long n, float* __restrict__ hi,
                                       float* __restrict__ lo) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = src[i];
  const float h = to_tf32(v);
  hi[i] = h;
  lo[i] = to_tf32(v - h);
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

// The survivor lists live in shared memory; no global workspace is needed any more.  The entry
// point keeps its (workspace, bytes) arguments for ABI stability: any non-null pointer will do.
extern "C" size_t ecb200_knn_tc_workspace_bytes(int B, int N, int k) {
  (void)B; (void)N; (void)k;
  return 256;
}

extern "C" int ecb200_knn_tc(const float* hi, const float* lo, const float* xx, int B, int C, int N,
                             int k, int sorted, int32_t* idx, void* workspace, size_t workspace_bytes,
                             void* stream) {
  (void)sorted;  // the rank-based final stage always yields nearest-first order
  (void)workspace_bytes;
  ECB_REQUIRE(hi && lo && xx && idx && workspace, "ecb200_knn_tc: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_knn_tc: bad shape B=%d N=%d", B, N);
  ECB_REQUIRE(C % KB == 0 && C >= KB && C <= KB * MAX_KB,
              "ecb200_knn_tc: C=%d must be a multiple of 32 in [32, 128]", C);
  ECB_REQUIRE(k >= 1 && k <= N, "ecb200_knn_tc: k=%d out of range for N=%d (selected index k out of range)", k, N);
  ECB_REQUIRE(k <= KMAX, "ecb200_knn_tc: k=%d exceeds %d (use ecb200_knn)", k, KMAX);
  TcArgs a = {hi, lo, hi, lo, xx, (long long)B * N, B, C, N, N, k, idx, nullptr, nullptr};
  return launch_knn(a, (cudaStream_t)stream);
}

extern "C" int ecb200_absmax(const float* x, long long n, float* amax, void* stream) {
  ECB_REQUIRE(x && amax && n >= 1, "ecb200_absmax: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  ECB_CUDA(cudaMemsetAsync(amax, 0, ECB200_AMAX_SLOTS * sizeof(float), st));
  const long long blocks = ecb200::ceil_div64(n, 256 * 16);
  absmax_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(x, n, reinterpret_cast<unsigned*>(amax));
  ECB_LAUNCH_CHECK("absmax_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_split_f16(const float* x, int B, int C, int N, const float* amax, void* hh, void* hl,
                                void* nb, float* xxs, float* cmax, float* hi, float* lo, float* xx,
                                void* stream) {
  ECB_REQUIRE(x && amax && hh && hl && nb && xxs && cmax, "ecb200_split_f16: null pointer");
  ECB_REQUIRE((hi == nullptr) == (lo == nullptr), "ecb200_split_f16: hi and lo come as a pair");
  ECB_REQUIRE(B >= 1 && B <= 65535 && C >= 2 && C % 2 == 0 && C <= 128 && N >= 1,
              "ecb200_split_f16: bad shape (C must be even and at most 128)");
  dim3 grid(ecb200::ceil_div(N, 32), B);
  split_f16_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
      x, C, N, reinterpret_cast<const unsigned*>(amax), static_cast<__half*>(hh), static_cast<__half*>(hl),
      static_cast<__half*>(nb), xxs, cmax, hi, lo, xx);
  ECB_LAUNCH_CHECK("split_f16_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_pack_xyz_f16(const float* x, int B, int C, int N, void* arow, void* brow, float* xxs,
                                   float* cmax, void* stream) {
  ECB_REQUIRE(x && arow && brow && xxs && cmax, "ecb200_pack_xyz_f16: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && C <= 4 && N >= 1, "ecb200_pack_xyz_f16: bad shape (C <= 4)");
  __half* ar = static_cast<__half*>(arow);
  __half* br = static_cast<__half*>(brow);
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 grid(ecb200::ceil_div(N, 256), B);
  switch (C) {
    case 1: pack_xyz_f16_kernel<1><<<grid, 256, 0, st>>>(x, N, ar, br, xxs, cmax); break;
    case 2: pack_xyz_f16_kernel<2><<<grid, 256, 0, st>>>(x, N, ar, br, xxs, cmax); break;
    case 3: pack_xyz_f16_kernel<3><<<grid, 256, 0, st>>>(x, N, ar, br, xxs, cmax); break;
    default: pack_xyz_f16_kernel<4><<<grid, 256, 0, st>>>(x, N, ar, br, xxs, cmax); break;
  }
  ECB_LAUNCH_CHECK("pack_xyz_f16_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_knn_tc_f16(const void* hh, const void* hl, const void* nb, const float* xxs,
                                 const float* cmax, int B, int C, int N, int k, int32_t* idx, long long* timeline,
                                 void* stream) {
  ECB_REQUIRE(hh && hl && nb && xxs && cmax && idx, "ecb200_knn_tc_f16: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_knn_tc_f16: bad shape B=%d N=%d", B, N);
  ECB_REQUIRE(C == 64 || C == 128, "ecb200_knn_tc_f16: C=%d must be 64 or 128", C);
  ECB_REQUIRE(k >= 1 && k <= N, "ecb200_knn_tc_f16: k=%d out of range for N=%d (selected index k out of range)", k, N);
  ECB_REQUIRE(k <= KMAX, "ecb200_knn_tc_f16: k=%d exceeds %d (use ecb200_knn)", k, KMAX);
  const float* h = static_cast<const float*>(hh);
  const float* l = static_cast<const float*>(hl);
  if (tc2_takes(C / 2, N, k, 3)) {   // 256 query rows per CTA, four epilogue warpgroups (knn_tc2.cu)
    Tc2Args a2 = {h, l, h, l, static_cast<const float*>(nb), xxs, cmax, (long long)B * N, B, C / 2, N, k,
                  KB / UMMA_K, idx, timeline};
    return launch_knn_tc2(a2, 3, (cudaStream_t)stream);
  }
  TcArgs a = {h, l, h, l, xxs, (long long)B * N, B, C / 2, N, N, k, idx, nullptr, timeline};
  a.bn = static_cast<const float*>(nb);
  a.cmax = cmax;
  return launch_knn_f16<3>(a, (cudaStream_t)stream);
}

extern "C" int ecb200_knn_tc_xyz(const void* arow, const void* brow, const float* xxs, const float* cmax, int B,
                                 int N, int k, int32_t* idx, long long* timeline, void* stream) {
  ECB_REQUIRE(arow && brow && xxs && cmax && idx, "ecb200_knn_tc_xyz: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_knn_tc_xyz: bad shape B=%d N=%d", B, N);
  ECB_REQUIRE(k >= 1 && k <= N, "ecb200_knn_tc_xyz: k=%d out of range for N=%d (selected index k out of range)", k, N);
  ECB_REQUIRE(k <= KMAX, "ecb200_knn_tc_xyz: k=%d exceeds %d (use ecb200_knn)", k, KMAX);
  const float* a_ = static_cast<const float*>(arow);
  const float* b_ = static_cast<const float*>(brow);
  if (tc2_takes(KB, N, k, 1)) {
    Tc2Args a2 = {a_, a_, b_, b_, nullptr, xxs, cmax, (long long)B * N, B, KB, N, k, 1, idx, timeline};
    return launch_knn_tc2(a2, 1, (cudaStream_t)stream);
  }
  TcArgs a = {a_, a_, b_, b_, xxs, (long long)B * N, B, KB, N, N, k, idx, nullptr, timeline};
  a.ksteps = 1;
  a.cmax = cmax;
  return launch_knn_f16<1>(a, (cudaStream_t)stream);
}

// diagnostic: the folded contraction itself, s^2 (x_i.x_j - 0.5 |x_j|^2), through the selection kernel's MMA
// sequence is not observable; this entry multiplies hi/lo only (no norm block) and returns s^2 x_i.x_j
extern "C" int ecb200_debug_tc_scores_f16(const void* hh, const void* hl, int B, int C, int N, float* scores,
                                          void* stream) {
  ECB_REQUIRE(hh && hl && scores, "ecb200_debug_tc_scores_f16: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_debug_tc_scores_f16: bad shape");
  ECB_REQUIRE(C == 64 || C == 128, "ecb200_debug_tc_scores_f16: C=%d must be 64 or 128", C);
  const float* h = static_cast<const float*>(hh);
  const float* l = static_cast<const float*>(hl);
  TcArgs a = {h, l, h, l, nullptr, (long long)B * N, B, C / 2, N, N, 1, nullptr, scores, nullptr};
  return launch_tc<true, 1, 5, 0, true, 3>(a, (cudaStream_t)stream);
}

extern "C" int ecb200_debug_tc_timeline(const float* hi, const float* lo, const float* xx, int B, int C,
                                        int N, int k, int32_t* idx, void* workspace,
                                        long long* timeline, void* stream) {
  ECB_REQUIRE(hi && lo && xx && idx && workspace && timeline, "ecb200_debug_tc_timeline: null pointer");
  ECB_REQUIRE(C % KB == 0 && C >= KB && C <= KB * MAX_KB && k <= KMAX && k <= N, "bad shape");
  TcArgs a = {hi, lo, hi, lo, xx, (long long)B * N, B, C, N, N, k, idx, nullptr, timeline};
  return launch_knn(a, (cudaStream_t)stream);
}

extern "C" int ecb200_debug_tc_scores(const float* hi, const float* lo, const float* xx, int B, int C,
                                      int N, float* scores, void* stream) {
  ECB_REQUIRE(hi && lo && xx && scores, "ecb200_debug_tc_scores: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1, "ecb200_debug_tc_scores: bad shape");
  ECB_REQUIRE(C % KB == 0 && C >= KB && C <= KB * MAX_KB,
              "ecb200_debug_tc_scores: C=%d must be a multiple of 32 in [32, 128]", C);
  TcArgs a = {hi, lo, hi, lo, xx, (long long)B * N, B, C, N, N, 1, nullptr, scores, nullptr};
  return launch_tc<true, 1, 5, 0>(a, (cudaStream_t)stream);
}

extern "C" int ecb200_split_rows_tf32(const float* src, long long n, float* hi, float* lo, void* stream) {
  ECB_REQUIRE(src && hi && lo && n >= 1, "ecb200_split_rows_tf32: bad arguments");
  split_rows_tf32_kernel<<<(unsigned)ecb200::ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(src, n, hi, lo);
  ECB_LAUNCH_CHECK("split_rows_tf32_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_point_gemm_tc(const float* xhi, const float* xlo, const float* whi, const float* wlo,
                                    long long M, int C, int Co2, float* Y, void* stream) {
  ECB_REQUIRE(xhi && xlo && whi && wlo && Y, "ecb200_point_gemm_tc: null pointer");
  ECB_REQUIRE(M >= 1 && M < (1LL << 31) && Co2 >= 1, "ecb200_point_gemm_tc: bad shape");
  ECB_REQUIRE(C % KB == 0 && C >= KB && C <= KB * MAX_KB,
              "ecb200_point_gemm_tc: C=%d must be a multiple of 32 in [32, 128]", C);
  // one "cloud": A = all M points, B = the 2Co rows of Wcat, dense store of the tiles into Y[M, 2Co]
  TcArgs a = {xhi, xlo, whi, wlo, nullptr, (long long)Co2, 1, C, (int)M, Co2, 1, nullptr, Y, nullptr};
  return launch_tc<true, 1, 5, 0>(a, (cudaStream_t)stream);
}
