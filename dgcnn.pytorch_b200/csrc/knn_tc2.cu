// Tensor-core kNN, 256 query rows per CTA (packed-FP16 operands, k <= 20).
//
// With FP16 halves the MMAs of knn_tc.cu take a third of a CTA's time: the kernel is bound by its
// selection epilogue, which two warpgroups (2 warps per scheduler) cannot keep issuing, and by a
// grid of B*N/128 CTAs that covers 148 SMs 1.73 times.  This variant gives every CTA TWO row tiles
// of 128 queries and four epilogue warpgroups (16 warps):
//   * a candidate tile (64 points) lands in shared memory once and feeds the MMAs of both row
//     tiles: half the L2 -> SM operand traffic per query;
//   * tensor memory: four accumulator stages of 64 columns (row tile x tile parity) + the two query
//     tiles (hi | lo halves, packed FP16) = 512 columns;
//   * warpgroup (R, p) owns row tile R and the candidate tiles of parity p, exactly like the two
//     groups of knn_tc.cu: a row is served by the two threads (R, 0) and (R, 1), each with its own
//     bins and survivor list;
//   * config 1 (B = 32, N = 1024) is ONE wave of 128 CTAs instead of two waves of 256.
// The selector (two sweeps: bins -> threshold -> survivors -> rank) is the one of knn_tc.cu; the
// ranking reads the row's two lists in place instead of copying them into a union first.
// Reference: knn(), models/dgcnn.py:6-12.
#include <cuda.h>
#include <math_constants.h>

#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "knn_tc_shared.cuh"
#include "tc_ptx.cuh"
#include "topk_select.cuh"

namespace {

using namespace ecb200::tc;
using namespace ecb200::topk;
using namespace ecb200::knntc;

constexpr int RT = 128;                       // rows per MMA (= TMEM lanes)
constexpr int NRT = 2;                        // row tiles per CTA
constexpr int BM = RT * NRT;                  // query rows per CTA
constexpr int BN = 64;                        // candidates per MMA tile (= TMEM columns per stage)
constexpr int TILE_BYTES = BN * KB * 4;       // 8 KB: one K-block of one operand half of a candidate tile
constexpr int STAGE_BYTES = 2 * TILE_BYTES;   // ring stage: (hi | lo) of a K-block, or two hi K-blocks
constexpr int NUM_EPI = 128;                  // threads per epilogue warpgroup
constexpr int NG = 2 * NRT;                   // warpgroups: (row tile, tile parity)
constexpr int NT = 64 + NG * NUM_EPI;         // producer warp + MMA warp + four epilogue warpgroups
constexpr int LS = NG * NUM_EPI;              // stride (entries) between a thread's consecutive survivor slots
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t A_COL0 = NG * BN;          // query tiles start behind the four accumulator stages
constexpr int UMMA_K = 8;                     // 32 bytes of K per instruction, in words
constexpr int KMAX = 20;                      // largest k of this variant
constexpr int GUARD = 8;                      // candidate columns between two overflow checks
constexpr int CAP = KMAX + GUARD;             // survivor slots per thread
constexpr int S = 5;                          // ring stages
constexpr int CH = 9;                         // own entries per sweep of the ranking

struct SharedTail {
  int cnt_x[NG][NUM_EPI];         // survivor counts, exchanged between the two threads of a row
  int sum_x[NG][NUM_EPI];         // sums of score-only ranks (tie detection)
  uint64_t a_full, b_full[S], b_empty[S], t_full[NG], t_empty[NG];
  uint32_t tmem_slot;
  float cmax_s[NRT];              // the cloud's max |s x_j|^2, one copy per row-tile pair
};

constexpr size_t SURV_BYTES = (size_t)CAP * LS * sizeof(uint64_t);
constexpr size_t OUT_BYTES = (size_t)BM * KMAX * sizeof(int32_t);
constexpr size_t SMEM_BYTES = 1024 + (size_t)S * STAGE_BYTES + SURV_BYTES + OUT_BYTES + sizeof(SharedTail);
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

// TERMS: 3 = hi.hi + hi.lo + lo.hi (the first sweep ranks with hi.hi and a margin); 1 = the hi arrays
// carry the whole product in one K step (xyz layer), both sweeps issue the same MMA, no margin.
// The column term -0.5*|x_j|^2 is always part of the contraction (FOLD of knn_tc.cu): with TERMS = 3 one
// more K step per tile whose candidate rows come from `map_bn`, with TERMS = 1 three slots of the one step.
template <int TERMS>
__global__ void __launch_bounds__(NT, 1)
knn_tc2_kernel(const float* __restrict__ a_hi_g, const float* __restrict__ a_lo_g,
               const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
               const __grid_constant__ CUtensorMap map_bn, const float* __restrict__ xx,
               const float* __restrict__ cmax_g, int N, int nkb, int ksteps, int k,
               int32_t* __restrict__ idx, long long* tl) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int C = nkb * KB;                     // words per operand row (<= 64)
  unsigned char* b_st = base;                                                        // [S][16 KB]
  uint64_t* surv = reinterpret_cast<uint64_t*>(b_st + (size_t)S * STAGE_BYTES);      // [CAP][512]
  int32_t* out_s = reinterpret_cast<int32_t*>(reinterpret_cast<unsigned char*>(surv) + SURV_BYTES);  // [256][k]
  SharedTail* T = reinterpret_cast<SharedTail*>(reinterpret_cast<unsigned char*>(out_s) + OUT_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y, rt = blockIdx.x;
  const int nct = (N + BN - 1) / BN;
  const int cloud_row0 = b * N;
  const int kpa = (nkb & 1) ? 1 : 2;          // hi K-blocks per stage in a single-term sweep
  const uint32_t A_PER_TILE = (uint32_t)(2 * C + (TERMS == 3 ? UMMA_K : 0));   // TMEM columns of a query tile: hi | lo | 2^15 constants

  // epilogue threads fetch their query row before anything else
  float4 pre[16];
  if (warp >= 2) {
    const int g = (warp - 2) >> 2, q = warp & 3;
    const int row = rt * BM + (g >> 1) * RT + q * 32 + lane;
    const float* src = ((g & 1) == 0 ? a_hi_g : a_lo_g) + (size_t)(cloud_row0 + row) * C;
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      pre[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row < N && 4 * e < C) pre[e] = __ldg(reinterpret_cast<const float4*>(src) + e);
    }
  }

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_bhi);
    prefetch_tensormap(&map_blo);
    mbar_init(&T->a_full, NG * NUM_EPI);
    for (int s = 0; s < S; ++s) { mbar_init(&T->b_full[s], 1); mbar_init(&T->b_empty[s], 1); }
    for (int s = 0; s < NG; ++s) { mbar_init(&T->t_full[s], 1); mbar_init(&T->t_empty[s], NUM_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&T->tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = T->tmem_slot;
  if (tl && threadIdx.x == 0) {  // diagnostics: wall-clock span and SM of every CTA
    unsigned long long t; unsigned sm;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    long long* e = tl + 3 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x);
    e[0] = (long long)t; e[2] = sm;
  }

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    auto advance = [&]() { if (++stage == S) { stage = 0; phase ^= 1; } };
    auto hi_sweep = [&]() {
      for (int ct = 0; ct < nct; ++ct)
        for (int kb = 0; kb < nkb; kb += kpa) {
          mbar_wait(&T->b_empty[stage], phase ^ 1);
          if (elect_one_sync()) {
            unsigned char* dst = b_st + (size_t)stage * STAGE_BYTES;
            mbar_expect_tx(&T->b_full[stage], kpa * TILE_BYTES);
            for (int j = 0; j < kpa; ++j)
              tma_load_3d(dst + j * TILE_BYTES, &map_bhi, &T->b_full[stage], (kb + j) * KB, ct * BN, b);
          }
          __syncwarp();
          advance();
        }
    };
    if (TERMS == 1) {
      hi_sweep();
      hi_sweep();
    } else {
      // first sweep: nkb hi K-blocks + the norm block of every tile, two blocks per stage
      for (int ct = 0; ct < nct; ++ct)
        for (int u0 = 0; u0 <= nkb; u0 += 2) {
          const int nun = min(2, nkb + 1 - u0);
          mbar_wait(&T->b_empty[stage], phase ^ 1);
          if (elect_one_sync()) {
            unsigned char* dst = b_st + (size_t)stage * STAGE_BYTES;
            mbar_expect_tx(&T->b_full[stage], nun * TILE_BYTES);
            for (int j = 0; j < nun; ++j) {
              if (u0 + j < nkb) tma_load_3d(dst + j * TILE_BYTES, &map_bhi, &T->b_full[stage], (u0 + j) * KB, ct * BN, b);
              else              tma_load_3d(dst + j * TILE_BYTES, &map_bn, &T->b_full[stage], 0, ct * BN, b);
            }
          }
          __syncwarp();
          advance();
        }
      // second sweep: (hi | lo) per K-block, then the norm block in a stage of its own
      for (int ct = 0; ct < nct; ++ct)
        for (int kb = 0; kb <= nkb; ++kb) {
          mbar_wait(&T->b_empty[stage], phase ^ 1);
          if (elect_one_sync()) {
            unsigned char* dst = b_st + (size_t)stage * STAGE_BYTES;
            if (kb < nkb) {
              mbar_expect_tx(&T->b_full[stage], 2 * TILE_BYTES);
              tma_load_3d(dst, &map_bhi, &T->b_full[stage], kb * KB, ct * BN, b);
              tma_load_3d(dst + TILE_BYTES, &map_blo, &T->b_full[stage], kb * KB, ct * BN, b);
            } else {
              mbar_expect_tx(&T->b_full[stage], TILE_BYTES);
              tma_load_3d(dst, &map_bn, &T->b_full[stage], 0, ct * BN, b);
            }
          }
          __syncwarp();
          advance();
        }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_f16(RT, BN);
    const uint32_t tbase = __shfl_sync(0xffffffffu, T->tmem_slot, 0);
    mbar_wait(&T->a_full, 0);
    tc_fence_after();
    int stage = 0;
    uint32_t phase = 0;
    int tile = 0;
    const uint32_t ring_lo = sw128_kmajor_desc_lo(smem_u32(b_st));
    constexpr uint32_t STAGE_STEP = STAGE_BYTES >> 4, LO_STEP = TILE_BYTES >> 4, K8_STEP = (UMMA_K * 4) >> 4;
    const uint32_t a_col = tbase + A_COL0;
    // one stage of a tile: `fn(R, d, aR)` issues that stage's MMAs for row tile R; `last` = the tile's
    // accumulators are complete after it
    auto run_stage = [&](int as, bool last, auto fn) {
      mbar_wait(&T->b_full[stage], phase);
      tc_fence_after();
      const uint32_t bh = ring_lo + (uint32_t)stage * STAGE_STEP;
      if (elect_one_sync()) {
#pragma unroll
        for (int R = 0; R < NRT; ++R) {
          fn(tbase + (uint32_t)((R * 2 + as) * BN), a_col + (uint32_t)(R * A_PER_TILE), bh);
          if (last) mma_commit(&T->t_full[R * 2 + as]);   // this row tile's accumulator is complete
        }
        mma_commit(&T->b_empty[stage]);      // the stage is free once these MMAs have read it
      }
      __syncwarp();
      if (++stage == S) { stage = 0; phase ^= 1; }
    };
    auto begin_tile = [&]() -> int {
      const int as = tile & 1;
      const uint32_t par = ((tile >> 1) & 1) ^ 1;
      mbar_wait(&T->t_empty[as], par);        // both row tiles' groups of this parity drained the stage
      mbar_wait(&T->t_empty[2 + as], par);
      tc_fence_after();
      return as;
    };
    if (TERMS == 1) {
      for (int sw = 0; sw < 2; ++sw)
        for (int ct = 0; ct < nct; ++ct, ++tile) {
          const int as = begin_tile();
          for (int kb = 0; kb < nkb; kb += kpa)
            run_stage(as, kb + kpa >= nkb, [&](uint32_t d, uint32_t aR, uint32_t bh) {
              for (int j = 0; j < kpa; ++j)
#pragma unroll
                for (int k8 = 0; k8 < KB / UMMA_K; ++k8)
                  if (k8 < ksteps)
                    mma_f16_ts_lo(d, aR + (uint32_t)((kb + j) * KB) + k8 * UMMA_K, bh + j * LO_STEP + k8 * K8_STEP, idesc,
                                  (kb | j | k8) != 0);
            });
        }
    } else {
      for (int ct = 0; ct < nct; ++ct, ++tile) {       // first sweep: hi.hi + norm
        const int as = begin_tile();
        for (int u0 = 0; u0 <= nkb; u0 += 2)
          run_stage(as, u0 + 2 > nkb, [&](uint32_t d, uint32_t aR, uint32_t bh) {
            const int nun = min(2, nkb + 1 - u0);
            for (int j = 0; j < nun; ++j) {
              if (u0 + j < nkb) {
#pragma unroll
                for (int k8 = 0; k8 < KB / UMMA_K; ++k8)
                  mma_f16_ts_lo(d, aR + (uint32_t)((u0 + j) * KB) + k8 * UMMA_K, bh + j * LO_STEP + k8 * K8_STEP, idesc,
                                (u0 | j | k8) != 0);
              } else {
                mma_f16_ts_lo(d, aR + (uint32_t)(2 * C), bh + j * LO_STEP, idesc, 1);
              }
            }
          });
      }
      for (int ct = 0; ct < nct; ++ct, ++tile) {       // second sweep: three terms + norm
        const int as = begin_tile();
        for (int kb = 0; kb <= nkb; ++kb)
          run_stage(as, kb == nkb, [&](uint32_t d, uint32_t aR, uint32_t bh) {
            if (kb < nkb) {
              const uint32_t ah = aR + (uint32_t)(kb * KB);
#pragma unroll
              for (int k8 = 0; k8 < KB / UMMA_K; ++k8) {
                mma_f16_ts_lo(d, ah + k8 * UMMA_K, bh + k8 * K8_STEP, idesc, (kb | k8) != 0);
                mma_f16_ts_lo(d, ah + k8 * UMMA_K, bh + LO_STEP + k8 * K8_STEP, idesc, 1);
                mma_f16_ts_lo(d, ah + (uint32_t)C + k8 * UMMA_K, bh + k8 * K8_STEP, idesc, 1);
              }
            } else {
              mma_f16_ts_lo(d, aR + (uint32_t)(2 * C), bh, idesc, 1);
            }
          });
      }
    }
  } else {
    // ===================== epilogue: selection (thread = query row x tile parity) =====================
    const int g = (warp - 2) >> 2;      // warpgroup
    const int R = g >> 1, p = g & 1;    // its row tile and tile parity
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int et = (threadIdx.x - 64) & (NUM_EPI - 1);
    const int row = rt * BM + R * RT + q * 32 + lane;
    const bool valid = row < N;
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * BN);
    auto pair_bar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(5 + R), "n"(2 * NUM_EPI) : "memory"); };
    {
      // query rows -> tensor memory: parity 0 copies the hi halves, parity 1 the lo halves of its row tile
      const uint32_t dst = tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + (uint32_t)R * A_PER_TILE + (uint32_t)(p * C);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
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
      if (TERMS == 3 && p == 0) {
        // the query side of the norm block: 2^15 (0x7800) in its first three K slots, the same for every row
        uint32_t r[8] = {0x78007800u, 0x00007800u, 0u, 0u, 0u, 0u, 0u, 0u};
        __syncwarp();
        tmem_st_32x8(tmem_base + ((uint32_t)(q * 32) << 16) + A_COL0 + (uint32_t)R * A_PER_TILE + (uint32_t)(2 * C), r);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&T->a_full);
    }
    if (p == 0 && et < 32) {
      // the cloud's largest squared norm from the per-block maxima of the operand kernel (read by the
      // pair's threads after the pair barrier of the threshold stage)
      const int nblk = (N + 31) / 32;
      float m = 0.f;
      for (int i = lane; i < nblk; i += 32) m = fmaxf(m, __ldg(cmax_g + (size_t)b * nblk + i));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0) T->cmax_s[R] = m;
    }
    float bin[NB];
#pragma unroll
    for (int u = 0; u < NB; ++u) bin[u] = -CUDART_INF_F;
    const int me = g * NUM_EPI + et;
    uint64_t* const sv = surv + me;
    const uint32_t sv_addr = smem_u32(sv);
    int cnt = 0;
    constexpr uint32_t SLOT = LS * sizeof(uint64_t);
    const float xi = valid ? __ldg(xx + cloud_row0 + row) : 0.f;
    float thr = CUDART_INF_F;
    int use = 0;  // how many times this group has consumed its accumulator stage
    auto run_tiles = [&](auto pass_tag) {
      constexpr int pass = decltype(pass_tag)::value;
      const int tile0 = pass * nct;
      for (int ct = (tile0 + p) & 1; ct < nct; ct += 2, ++use) {
        mbar_wait(&T->t_full[g], use & 1);
        tc_fence_after();
        auto process = [&](uint32_t(&cur)[32], const int c2, const bool second_pass) {
          // the accumulator already is the score; candidates past the end of the cloud were loaded as NaN
          // (per-cloud tensor map with NaN fill) and score NaN: fmaxf drops them, >= rejects them
          float v[32];
#pragma unroll
          for (int u = 0; u < 32; ++u) v[u] = __uint_as_float(cur[u]);
          if (!second_pass) {
#pragma unroll
            for (int u = 0; u < 32; ++u) bin[u] = fmaxf(bin[u], v[u]);
          } else {
            const int jb = ct * BN + c2 * 32;
#pragma unroll
            for (int h = 0; h < 32 / GUARD; ++h) {
              if (cnt > CAP - GUARD) {  // rare (ties, clustered data): keep the thread's own best k
                thr = fmaxf(thr, shrink_survivors(sv, cnt, k, LS));
                cnt = k;
              }
#pragma unroll
              for (int u = h * GUARD; u < (h + 1) * GUARD; ++u) {
                if (v[u] >= thr) {   // entry = (score bits << 32) | j
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
        __syncwarp();
        // both 32-column halves of the stage move to registers, then the stage goes back to the MMA issuer
        if (pass == 0) {
          // first sweep (the 32 bins are live): one 32-column half at a time
          uint32_t ra[32];
          tmem_ld_32x32_issue(lane_base, ra);
          tmem_ld_wait(ra);
          process(ra, 0, false);
          __syncwarp();
          tmem_ld_32x32_issue(lane_base + 32u, ra);
          tmem_ld_wait(ra);
          tc_fence_before();
          mbar_arrive(&T->t_empty[g]);
          process(ra, 1, false);
        } else {
          // second sweep: both halves move to registers, then the stage goes back to the MMA issuer
          uint32_t ra[32], rb[32];
          tmem_ld_32x32_issue(lane_base, ra);
          tmem_ld_32x32_issue(lane_base + 32u, rb);
          tmem_ld_wait(ra);
          tmem_ld_wait(rb);
          tc_fence_before();
          mbar_arrive(&T->t_empty[g]);
          process(ra, 0, true);
          process(rb, 1, true);
        }
      }
    };
    run_tiles(std::integral_constant<int, 0>{});
    sort_bins_desc<NB>(bin);
    // The row's two threads pool their bins: the k-th largest of the union of two descending lists is
    // max_i min(mine[i-1], theirs[k-i-1]).  The partner's list travels through the (still empty)
    // survivor area, written reversed so that the reader walks it with static offsets.
    float* exch = reinterpret_cast<float*>(surv);
    {
      float* wr = exch + (size_t)(k - 1) * LS + me;     // slot k-1-u  <-  bin[u]
#pragma unroll
      for (int u = 0; u < KMAX; ++u)
        if (u < k) wr[-u * LS] = bin[u];
    }
    pair_bar();
    const float* other = exch + ((g ^ 1) * NUM_EPI + et);
    float tau = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i <= KMAX; ++i) {
      if (i <= k) {
        const float mine = i == 0 ? CUDART_INF_F : bin[i > 0 ? i - 1 : 0];
        const float theirs = i == k ? CUDART_INF_F : other[(i < KMAX ? i : 0) * LS];
        tau = fmaxf(tau, fminf(mine, theirs));
      }
    }
    const float cmax = T->cmax_s[R];
    // first sweep scored with hi.hi only: lower the bound by the rigorous margin of knn_tc.cu
    float margin = 1.1f * 0.0009765625f * sqrtf(xi * cmax) + 1e-30f;
    margin += 5.96e-8f * sqrtf((float)(2 * C)) * (sqrtf(xi) + sqrtf(cmax));
    if (TERMS == 1) margin = 0.f;
    thr = valid ? fmaxf(tau - margin, -3.0e38f) : CUDART_INF_F;
    pair_bar();   // bins read: the area may take survivors
    run_tiles(std::integral_constant<int, 1>{});
    if (!valid) cnt = 0;
    if (cnt > KMAX + 4) {        // keeps the ranking sweeps short; exact (at most k of a list can make the top k)
      shrink_survivors(sv, cnt, k, LS);
      cnt = k;
    }
    // Top-k of the row's survivors (both lists), nearest first: rank = number of better entries =
    // output slot.  Fast path by score only; rows with equal scores are redone under the total order.
    T->cnt_x[g][et] = cnt;
    pair_bar();
    const int cnt0 = T->cnt_x[R * 2][et], cnt1 = T->cnt_x[R * 2 + 1][et];
    const int total = cnt0 + cnt1;
    const uint64_t* l0 = surv + (R * 2) * NUM_EPI + et;     // the row's list of parity 0; parity 1 behind it
    const uint64_t* l1 = l0 + NUM_EPI;
    auto entry = [&](int e) -> uint64_t { return e < cnt0 ? l0[e * LS] : l1[(e - cnt0) * LS]; };
    int32_t* orow = out_s + (R * RT + q * 32 + lane) * k;
    const int half = (total + 1) >> 1;
    const int lo = p * half, hi = min(total, lo + half);
    int ranksum = 0;
    for (int e0 = lo; e0 < hi; e0 += CH) {
      float so[CH], rf[CH];
      uint32_t jo[CH];
#pragma unroll
      for (int u = 0; u < CH; ++u) {
        const uint64_t w = e0 + u < hi ? entry(e0 + u) : 0ull;
        so[u] = e0 + u < hi ? __uint_as_float((uint32_t)(w >> 32)) : CUDART_INF_F;
        jo[u] = (uint32_t)w;
        rf[u] = 0.f;
      }
#pragma unroll
      for (int L = 0; L < 2; ++L) {
        const float* sc = reinterpret_cast<const float*>(L == 0 ? l0 : l1) + 1;   // score word of slot f
        const int n = L == 0 ? cnt0 : cnt1;
#pragma unroll 4
        for (int f = 0; f < n; ++f) {
          const float sf = sc[f * 2 * LS];
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
    pair_bar();
    if (valid && T->sum_x[R * 2][et] + T->sum_x[R * 2 + 1][et] != total * (total - 1) / 2) {
      // equal scores in this row: exact ranks under the total order
      for (int e0 = lo; e0 < hi; e0 += 4) {
        uint64_t own[4];
        int rank[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          own[u] = e0 + u < hi ? ordered_key(entry(e0 + u)) : ~0ull;
          rank[u] = 0;
        }
#pragma unroll 1
        for (int f = 0; f < total; ++f) {
          const uint64_t kf = ordered_key(entry(f));
#pragma unroll
          for (int u = 0; u < 4; ++u) rank[u] += (kf > own[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (e0 + u < hi && rank[u] < k) orow[rank[u]] = (int32_t)min(key_index(own[u]), (uint32_t)(N - 1));
      }
    }
    if (valid && p == 0)
      for (int s = total; s < k; ++s) orow[s] = N - 1;  // only with NaN input
    pair_bar();
    {
      // the row tile's rows are consecutive in idx: one contiguous, coalesced block of nrows*k words
      const int nrows = max(0, min(RT, N - rt * BM - R * RT));
      int32_t* dst = idx + (size_t)(cloud_row0 + rt * BM + R * RT) * k;
      const int32_t* src = out_s + R * RT * k;
      for (int w = p * NUM_EPI + et; w < nrows * k; w += 2 * NUM_EPI) dst[w] = src[w];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
  if (tl && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    tl[3 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) + 1] = (long long)t;
  }
}

template <int TERMS>
int launch(const ecb200::knntc::Tc2Args& a, cudaStream_t st) {
  OperandMaps Bm;   // per-cloud maps: candidate rows past the end of a cloud read as NaN
  int rc = make_cloud_map(&Bm.hi, a.b_hi, a.clouds, a.N, a.Cw, BN);
  if (!rc) rc = make_cloud_map(&Bm.lo, a.b_lo, a.clouds, a.N, a.Cw, BN);
  if (rc) return rc;
  CUtensorMap Bn = Bm.hi;   // placeholder unless the norm block is a separate operand
  if (TERMS == 3) {
    rc = make_cloud_map(&Bn, a.bn, a.clouds, a.N, KB, BN);
    if (rc) return rc;
  }
  auto kern = knn_tc2_kernel<TERMS>;
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if (ecb200::first_use_on_device(seen))
    ECB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  dim3 grid(ecb200::ceil_div(a.N, BM), a.clouds);
  kern<<<grid, NT, SMEM_BYTES, st>>>(a.a_hi, a.a_lo, Bm.hi, Bm.lo, Bn, a.xx, a.cmax, a.N, a.Cw / KB, a.ksteps, a.k,
                                     a.idx, a.tl);
  ECB_LAUNCH_CHECK("knn_tc2_kernel");
  return ECB200_OK;
}

}  // namespace

namespace ecb200 {
namespace knntc {

// ECB200_KNN_ROWS = 128 | 256 forces a variant (read at every call: a host-side getenv); otherwise by
// measurement on B200 (tools/check_f16.py, B = 32, N = 1024, k = 20): the xyz layer, whose MMAs are
// negligible, gains 7 % from the single wave (46.3 vs 50.0 us); the feature layers are bound by the
// issue rate of the selection epilogue either way (50.9 vs 51.3 us) and stay on the 128-row kernel.
bool tc2_takes(int Cw, int N, int k, int terms) {
  const char* e = getenv("ECB200_KNN_ROWS");
  const int rows = e ? atoi(e) : 0;
  // tensor memory: 4 x 64 accumulator columns + two query tiles of 2 Cw (+ 8 constants): three-term
  // products fit for Cw = 32 (64 channels) only
  if (!(k <= KMAX && N > RT && (Cw == KB || (Cw == 2 * KB && terms == 1)))) return false;
  if (rows == 128) return false;
  if (rows == 256) return true;
  return terms == 1;
}

int launch_knn_tc2(const Tc2Args& a, int terms, cudaStream_t st) {
  return terms == 1 ? launch<1>(a, st) : launch<3>(a, st);
}

}  // namespace knntc
}  // namespace ecb200
