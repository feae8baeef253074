// Backward GEMMs of the split edge MLP on the tensor cores (tcgen05 kind::tf32, 3xTF32).
//
//   gemm_dx : dx[B,C,N]    = dY[M,2Co] . Wcat[2Co,C]      K = 2Co (up to 512): both operands are
//             streamed K-block by K-block through a TMA / mbarrier ring (K-major tiles),
//             the result is written straight back in the reference's channel-major layout.
//   gemm_dw : dWcat[2Co,C] = dY^T . X                      K = M = B*N points: the reduction runs
//             over the ROWS of the point-major arrays, so both operands are consumed as
//             MN-major tiles (no transposed copies); the M-long reduction is cut into slabs
//             over CTAs and the partial 128 x C tiles are reduced with fp32 atomics.
//
// Operands are the tf32 hi/lo halves of the fp32 values (ecb200_split_rows_tf32 /
// ecb200_split_tf32); D += Ahi.Bhi + Ahi.Blo + Alo.Bhi accumulates in TMEM in FP32.
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2-5 = epilogue (tcgen05.ld, lane = output row).
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace ecb200::tc;

constexpr int BM = 128;                   // output rows per CTA (= TMEM lanes)
constexpr int KB = 32;                    // K elements per stage
constexpr int SUB = BM * KB * 4;          // 16 KB: one operand half for 128 rows x 32 k
constexpr int STAGES = 3;
constexpr int NT = 64 + 128;
constexpr uint32_t TMEM_COLS = 128;
constexpr int UMMA_K = 8;

struct Tail {
  uint64_t full[STAGES], empty[STAGES], done;
  uint32_t tmem_slot;
};
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * 4 * SUB + sizeof(Tail);

// MN-major tf32 operand.  For 4-byte types the tensor core only takes MN-major tiles in the
// "128-byte swizzle with 32-byte atomicity" layout (descriptor layout type 1; TMA
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 4 K-rows x 128 bytes (32 MN elements), the
// 32-byte chunks of a row XOR-ed with (row & 3).  Consecutive MN atoms are `lbo` bytes apart,
// consecutive K atoms 512 bytes; one kind::tf32 instruction (K = 8) spans two K atoms.
__device__ __forceinline__ uint64_t make_sw128_mnmajor_desc(uint32_t smem_addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo >> 4) << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;   // SWIZZLE_128B_BASE32B
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc_tf32_major(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct EpiDx {  // D[m, c] -> dx[b, c, n]   (m = b*N + n)
  float* dx; int C, N; long long M;
};
struct EpiDw {  // D[o, c] += into dWcat[o, c]
  float* dW; int C, rows;
};

// MN_MAJOR = false: A[M rows, K] and B[BN rows, K], K contiguous (TMA box 32 k x 128 rows).
// MN_MAJOR = true : A[K rows, M'] and B[K rows, BN], rows = reduction index (TMA box 32 cols x
//                   32 rows per 32-column atom).
// `three` = 1: D += Ahi.Bhi + Ahi.Blo + Alo.Bhi (3xTF32, fp32-equivalent); 0: D += A.B in plain TF32 from
// the `hi` maps only (which may then be the raw fp32 arrays: the tensor core reads their top 19 bits).
// blockIdx.z = tile of BN output columns (rows [z*BN, +BN) of B).
// C22: thread-block clusters of 2 x 2 output tiles (cluster dims (2,1,2), K-major operands, BN = 128): the
// two CTAs of a row pair share their B tile and the two of a column pair their A tile, so every CTA
// fetches HALF of each and TMA multicasts it to its partner -- the L2 -> SM operand traffic, which
// bounds a 128 x 128 fp32 tile at K = 512, is halved.
template <bool MN_MAJOR, class Epi, bool C22 = false>
__global__ void __launch_bounds__(NT, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap a_hi, const __grid_constant__ CUtensorMap a_lo,
               const __grid_constant__ CUtensorMap b_hi, const __grid_constant__ CUtensorMap b_lo,
               int BN, long long K, long long kslab, int three, Epi epi) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  Tail* T = reinterpret_cast<Tail*>(base + (size_t)STAGES * 4 * SUB);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x;                     // 128-row tile of the output
  const int nt = blockIdx.z;                     // BN-column tile of the output
  const long long k_beg = (long long)blockIdx.y * kslab;
  const long long k_end = k_beg + kslab < K ? k_beg + kslab : K;
  const int nkb = (int)((k_end - k_beg + KB - 1) / KB);

  uint32_t cx = 0, cz = 0;
  uint16_t amask = 0, bmask = 0, emask = 0;
  if (C22) {
    const uint32_t crank = cluster_ctarank();          // = x + 2 z inside the (2,1,2) cluster
    cx = crank & 1u; cz = crank >> 1;
    amask = (uint16_t)((1u << cx) | (1u << (cx + 2)));  // same row tile, both column tiles: share A
    bmask = (uint16_t)(3u << (2 * cz));                 // same column tile, both row tiles: share B
    emask = (uint16_t)(amask | bmask);                  // everyone whose multicasts land in my stages
  }
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&a_hi); prefetch_tensormap(&a_lo);
    prefetch_tensormap(&b_hi); prefetch_tensormap(&b_lo);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&T->full[s], 1); mbar_init(&T->empty[s], C22 ? 3 : 1); }
    mbar_init(&T->done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(&T->tmem_slot);
  tc_fence_before();
  __syncthreads();
  if (C22) cluster_sync_all();   // the partners' barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = T->tmem_slot;
  const uint32_t b_bytes = (uint32_t)(BN * KB * 4);  // one half of the B stage

  if (warp == 0) {
    // TMA producer: the whole warp stays in the loop, one elected lane issues (warp-uniform operands)
    int stage = 0;
    uint32_t phase = 0;
    const int brow = nt * BN;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&T->empty[stage], phase ^ 1);
      if (elect_one_sync()) {
        unsigned char* st = base + (size_t)stage * 4 * SUB;  // [A hi | A lo | B hi | B lo]
        mbar_expect_tx(&T->full[stage], three ? 2 * SUB + 2 * b_bytes : SUB + b_bytes);
        const int k0 = (int)(k_beg + (long long)kb * KB);
        if (C22) {
          // my half of the A tile (64 rows) to both CTAs of my row pair... of my column pair, and my
          // half of the B tile to both CTAs of my row pair; maps carry 64-row boxes
          constexpr int HALF = SUB / 2;
          tma_load_2d_mc(st + cz * HALF, &a_hi, &T->full[stage], k0, mt * BM + (int)cz * 64, amask);
          tma_load_2d_mc(st + 2 * SUB + cx * HALF, &b_hi, &T->full[stage], k0, brow + (int)cx * 64, bmask);
          if (three) {
            tma_load_2d_mc(st + SUB + cz * HALF, &a_lo, &T->full[stage], k0, mt * BM + (int)cz * 64, amask);
            tma_load_2d_mc(st + 3 * SUB + cx * HALF, &b_lo, &T->full[stage], k0, brow + (int)cx * 64, bmask);
          }
        } else if (!MN_MAJOR) {
          tma_load_2d(st, &a_hi, &T->full[stage], k0, mt * BM);
          tma_load_2d(st + 2 * SUB, &b_hi, &T->full[stage], k0, brow);
          if (three) {
            tma_load_2d(st + SUB, &a_lo, &T->full[stage], k0, mt * BM);
            tma_load_2d(st + 3 * SUB, &b_lo, &T->full[stage], k0, brow);
          }
        } else {
          // one box per 32-column atom: 32 columns x 32 reduction rows = 4 KB
          for (int a = 0; a < BM / 32; ++a) {
            tma_load_2d(st + a * 4096, &a_hi, &T->full[stage], mt * BM + a * 32, k0);
            if (three) tma_load_2d(st + SUB + a * 4096, &a_lo, &T->full[stage], mt * BM + a * 32, k0);
          }
          for (int a = 0; a < BN / 32; ++a) {
            tma_load_2d(st + 2 * SUB + a * 4096, &b_hi, &T->full[stage], brow + a * 32, k0);
            if (three) tma_load_2d(st + 3 * SUB + a * 4096, &b_lo, &T->full[stage], brow + a * 32, k0);
          }
        }
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // MMA issuer: whole warp in the loop, one elected lane issues
    const uint32_t idesc = make_idesc_tf32_major(BM, BN, MN_MAJOR, MN_MAJOR);
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, T->tmem_slot, 0);
    int stage = 0;
    uint32_t phase = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      mbar_wait(&T->full[stage], phase);
      tc_fence_after();
      const uint32_t st = smem_u32(base + (size_t)stage * 4 * SUB);
      if (elect_one_sync()) {
#pragma unroll
        for (int k8 = 0; k8 < KB / UMMA_K; ++k8) {
          uint64_t dah, dal, dbh, dbl;
          if (!MN_MAJOR) {
            const uint32_t ko = (uint32_t)(k8 * UMMA_K * 4);  // 32 bytes inside the swizzle row
            dah = make_sw128_kmajor_desc(st + ko);
            dal = make_sw128_kmajor_desc(st + SUB + ko);
            dbh = make_sw128_kmajor_desc(st + 2 * SUB + ko);
            dbl = make_sw128_kmajor_desc(st + 3 * SUB + ko);
          } else {
            const uint32_t ko = (uint32_t)(k8 * 1024);  // next 8-row K atom
            dah = make_sw128_mnmajor_desc(st + ko, 4096);
            dal = make_sw128_mnmajor_desc(st + SUB + ko, 4096);
            dbh = make_sw128_mnmajor_desc(st + 2 * SUB + ko, 4096);
            dbl = make_sw128_mnmajor_desc(st + 3 * SUB + ko, 4096);
          }
          mma_tf32(tmem_d, dah, dbh, idesc, (kb | k8) != 0);
          if (three) {
            mma_tf32(tmem_d, dah, dbl, idesc, 1);
            mma_tf32(tmem_d, dal, dbh, idesc, 1);
          }
        }
        if (C22) mma_commit_mc(&T->empty[stage], emask);
        else     mma_commit(&T->empty[stage]);
        if (kb + 1 == nkb) mma_commit(&T->done);
      }
      __syncwarp();
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue (thread = output row) =====================
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row inside the tile
    if (nkb > 0) {
      mbar_wait(&T->done, 0);
      tc_fence_after();
      if constexpr (Epi::STAGED) {
        // the operand ring is dead: stage the 128 x BN tile there ([128][BN+1] floats), then let the
        // epilogue reduce columns / write whole rows from shared memory
        float* zs = reinterpret_cast<float*>(base);
        const int ld = BN + 1;
        for (int c4 = 0; c4 < BN / 32; ++c4) {
          float v[32];
          __syncwarp();
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c4 * 32), v);
#pragma unroll
          for (int u = 0; u < 32; ++u) zs[r * ld + c4 * 32 + u] = v[u];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        epi.finish(zs, ld, mt * BM, nt * BN, BN, (int)threadIdx.x - 64);
      } else {
        for (int c4 = 0; c4 < BN / 32; ++c4) {
          float v[32];
          __syncwarp();
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c4 * 32), v);
          epi.store(mt * BM + r, c4 * 32, v);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<TMEM_COLS>(tmem_base);
  if (C22) cluster_sync_all();   // no CTA exits while a partner can still signal its barriers
}

struct EpiDxImpl : EpiDx {
  static constexpr bool STAGED = false;
  __device__ __forceinline__ void store(long long m, int c0, const float (&v)[32]) const {
    if (m >= M) return;
    const long long b = m / N;
    const int n = (int)(m - b * N);
    float* o = dx + ((size_t)b * C + c0) * N + n;  // lanes = consecutive points: coalesced per channel
#pragma unroll
    for (int u = 0; u < 32; ++u)
      if (c0 + u < C) o[(size_t)u * N] = v[u];
  }
};
struct EpiDwImpl : EpiDw {
  static constexpr bool STAGED = false;
  __device__ __forceinline__ void store(long long o, int c0, const float (&v)[32]) const {
    if (o >= rows) return;
    float* p = dW + (size_t)o * C + c0;
#pragma unroll
    for (int u = 0; u < 32; ++u)
      if (c0 + u < C) atomicAdd(p + u, v[u]);
  }
};

// Z[m, n0 + c] = D (row-major, whole 512-byte rows per warp instruction) and, per output column, the
// sum and sum of squares over the tile's rows (BatchNorm statistics of conv5, fp64 atomics): the
// statistics pass over the 128 MiB result and its re-read disappear.
struct EpiRowStats {
  static constexpr bool STAGED = true;
  float* Z; long long M; int ldz; double* stats; int E;
  __device__ __forceinline__ void finish(const float* zs, int ld, long long m0, int n0, int BN, int t) const {
    const int rows = (int)(M - m0 < BM ? M - m0 : BM);
    if (stats) {   // thread t = column t of the tile
      if (t < BN) {
        float s = 0.f, q = 0.f;
        for (int r = 0; r < rows; ++r) {
          const float z = zs[r * ld + t];
          s += z;
          q = fmaf(z, z, q);
        }
        atomicAdd(stats + n0 + t, (double)s);
        atomicAdd(stats + E + n0 + t, (double)q);
      }
      if (m0 == 0 && n0 == 0 && t == 0) atomicAdd(stats + 2 * E, (double)M);
    }
    const int w = t >> 5, l = t & 31;   // warp w writes rows w, w+4, ...: one coalesced row segment each
    for (int r = w; r < rows; r += 4) {
      float* o = Z + (size_t)(m0 + r) * ldz + n0;
      for (int c = l; c < BN; c += 32) o[c] = zs[r * ld + c];
    }
  }
};

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
// row-major [rows, cols] fp32, box = box_cols x box_rows, 128-byte swizzle, zero fill outside
int make_map(CUtensorMap* m, const float* p, long long rows, long long cols, int box_cols, int box_rows,
             CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    ecb200::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECB200_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecb200::set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ECB200_ERR_CUDA;
  }
  return ECB200_OK;
}

template <class K>
int opt_in_smem(K kern, bool (&seen)[ecb200::kMaxDevices]) {
  if (ecb200::first_use_on_device(seen))
    ECB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  return ECB200_OK;
}

__global__ void transpose_split_kernel(const float* __restrict__ W, int R, int C, float* __restrict__ hiT,
                                       float* __restrict__ loT) {
  // W [R, C] row-major -> hiT/loT [C, R] (tf32 halves): tiny (Wcat), one thread per element
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= R * C) return;
  const int c = e / R, r = e % R;
  const float v = W[(size_t)r * C + c];
  const float h = to_tf32(v);
  hiT[e] = h;
  loT[e] = to_tf32(v - h);
}

}  // namespace

extern "C" int ecb200_transpose_split_tf32(const float* W, int R, int C, float* hiT, float* loT,
                                           void* stream) {
  ECB_REQUIRE(W && hiT && loT && R >= 1 && C >= 1, "ecb200_transpose_split_tf32: bad arguments");
  transpose_split_kernel<<<ecb200::ceil_div(R * C, 256), 256, 0, (cudaStream_t)stream>>>(W, R, C, hiT, loT);
  ECB_LAUNCH_CHECK("transpose_split_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_gemm_dx_tc(const float* dYhi, const float* dYlo, const float* WTh, const float* WTl,
                                 int B, int C, int N, int Co2, float* dx, void* stream) {
  ECB_REQUIRE(dYhi && dYlo && WTh && WTl && dx, "ecb200_gemm_dx_tc: null pointer");
  ECB_REQUIRE(B >= 1 && N >= 1, "ecb200_gemm_dx_tc: bad shape");
  ECB_REQUIRE((C == 32 || C == 64 || C == 128) && Co2 % KB == 0,
              "ecb200_gemm_dx_tc: needs C in {32,64,128} and 2Co a multiple of 32 (C=%d 2Co=%d)", C, Co2);
  const long long M = (long long)B * N;
  CUtensorMap ah, al, bh, bl;
  int rc;
  if ((rc = make_map(&ah, dYhi, M, Co2, KB, BM))) return rc;
  if ((rc = make_map(&al, dYlo, M, Co2, KB, BM))) return rc;
  if ((rc = make_map(&bh, WTh, C, Co2, KB, C))) return rc;   // B = Wcat^T [C, 2Co], K contiguous
  if ((rc = make_map(&bl, WTl, C, Co2, KB, C))) return rc;
  auto kern = gemm_tc_kernel<false, EpiDxImpl>;
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if ((rc = opt_in_smem(kern, seen))) return rc;
  EpiDxImpl epi;
  epi.dx = dx; epi.C = C; epi.N = N; epi.M = M;
  dim3 grid((unsigned)ecb200::ceil_div64(M, BM), 1);
  kern<<<grid, NT, SMEM_BYTES, (cudaStream_t)stream>>>(ah, al, bh, bl, C, (long long)Co2, (long long)Co2, 1, epi);
  ECB_LAUNCH_CHECK("gemm_tc_kernel<dx>");
  return ECB200_OK;
}

extern "C" int ecb200_gemm_dw_tc(const float* dYhi, const float* dYlo, const float* xhi, const float* xlo,
                                 long long M, int C, int Co2, float* dWcat, void* stream) {
  ECB_REQUIRE(dYhi && dYlo && xhi && xlo && dWcat, "ecb200_gemm_dw_tc: null pointer");
  ECB_REQUIRE(M >= 1, "ecb200_gemm_dw_tc: bad shape");
  ECB_REQUIRE((C == 32 || C == 64 || C == 128) && Co2 % 32 == 0,
              "ecb200_gemm_dw_tc: needs C in {32,64,128} and 2Co a multiple of 32 (C=%d 2Co=%d)", C, Co2);
  cudaStream_t st = (cudaStream_t)stream;
  ECB_CUDA(cudaMemsetAsync(dWcat, 0, sizeof(float) * (size_t)Co2 * C, st));
  CUtensorMap ah, al, bh, bl;
  int rc;
  // MN-major operands: A = dY [K = M rows, 2Co cols], B = X [K = M rows, C cols]; box 32 cols x 32 rows
  if ((rc = make_map(&ah, dYhi, M, Co2, 32, KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&al, dYlo, M, Co2, 32, KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&bh, xhi, M, C, 32, KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  if ((rc = make_map(&bl, xlo, M, C, 32, KB, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B))) return rc;
  auto kern = gemm_tc_kernel<true, EpiDwImpl>;
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if ((rc = opt_in_smem(kern, seen))) return rc;
  const int mtiles = ecb200::ceil_div(Co2, BM);
  long long slabs = (ecb200::kNumSMs + mtiles - 1) / mtiles;      // about one CTA per SM
  long long kslab = ecb200::ceil_div64(ecb200::ceil_div64(M, slabs), KB) * KB;
  slabs = ecb200::ceil_div64(M, kslab);
  EpiDwImpl epi;
  epi.dW = dWcat; epi.C = C; epi.rows = Co2;
  dim3 grid(mtiles, (unsigned)slabs);
  kern<<<grid, NT, SMEM_BYTES, st>>>(ah, al, bh, bl, C, M, kslab, 1, epi);
  ECB_LAUNCH_CHECK("gemm_tc_kernel<dw>");
  return ECB200_OK;
}

// conv5 of the backbone (models/dgcnn.py:74-78,:102) as a per-point GEMM with its BatchNorm statistics
// in the epilogue: Z[M, E] = X[M, K] . W[E, K]^T, stats += [sum z | sum z^2 | M].  xlo / wlo NULL selects
// plain TF32 on the raw fp32 operands (what the library convolution does with cudnn.allow_tf32, the
// PyTorch default); with the tf32 halves of both operands the product is 3xTF32 (fp32-equivalent).
extern "C" int ecb200_embed_gemm(const float* xhi, const float* xlo, const float* whi, const float* wlo,
                                 long long M, int K, int E, float* Z, double* stats, void* stream) {
  ECB_REQUIRE(xhi && whi && Z, "ecb200_embed_gemm: null pointer");
  ECB_REQUIRE((xlo == nullptr) == (wlo == nullptr), "ecb200_embed_gemm: the lo halves come in pairs");
  ECB_REQUIRE(M >= 1 && K >= KB && K % KB == 0 && E >= 128 && E % 128 == 0,
              "ecb200_embed_gemm: needs K a multiple of 32 and E a multiple of 128 (K=%d E=%d)", K, E);
  const int three = xlo != nullptr;
  EpiRowStats epi;
  epi.Z = Z; epi.M = M; epi.ldz = E; epi.stats = stats; epi.E = E;
  dim3 grid((unsigned)ecb200::ceil_div64(M, BM), 1, (unsigned)(E / 128));
  const bool c22 = grid.x % 2 == 0 && grid.z % 2 == 0;   // 2 x 2 clusters need even tile counts
  const int box = c22 ? 64 : 128;
  CUtensorMap ah, al, bh, bl;
  int rc;
  if ((rc = make_map(&ah, xhi, M, K, KB, box))) return rc;
  if ((rc = make_map(&al, three ? xlo : xhi, M, K, KB, box))) return rc;
  if ((rc = make_map(&bh, whi, E, K, KB, box))) return rc;
  if ((rc = make_map(&bl, three ? wlo : whi, E, K, KB, box))) return rc;
  if (!c22) {
    auto kern = gemm_tc_kernel<false, EpiRowStats, false>;
    static thread_local bool seen[ecb200::kMaxDevices] = {};
    if ((rc = opt_in_smem(kern, seen))) return rc;
    kern<<<grid, NT, SMEM_BYTES, (cudaStream_t)stream>>>(ah, al, bh, bl, 128, (long long)K, (long long)K, three, epi);
    ECB_LAUNCH_CHECK("gemm_tc_kernel<embed>");
    return ECB200_OK;
  }
  auto kern = gemm_tc_kernel<false, EpiRowStats, true>;
  static thread_local bool seen2[ecb200::kMaxDevices] = {};
  if ((rc = opt_in_smem(kern, seen2))) return rc;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 2;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, ah, al, bh, bl, 128, (long long)K, (long long)K, three, epi);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    ecb200::set_error("launch of gemm_tc_kernel<embed, 2x2 cluster> failed: %s", cudaGetErrorString(e));
    return ECB200_ERR_CUDA;
  }
  return ECB200_OK;
}
