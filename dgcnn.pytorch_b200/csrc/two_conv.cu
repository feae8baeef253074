// Two-conv edge block, fused forward:
//     get_graph_feature(x, k) -> conv1 (1x1 Conv2d + BN2d + LeakyReLU) -> conv2 (same) -> max over k
// the shape of the reference's PositionEmbedding (models/layers.py:45-52) and of upstream's part-seg /
// sem-seg EdgeConv blocks.  The first conv splits like every EdgeConv layer,
//     e1_ij = W1.[x_j (- x_i) ; x_i] = U[idx_ij] + V[i]            (one per-point GEMM, Y = [U | V]),
// the second is a genuine per-edge GEMM z_ij = W2 . LeakyReLU(a1 e1_ij + b1) on a tensor that is never
// materialised.  One persistent kernel, per tile of P = 128/k points (P*k <= 128 edge rows):
//   1. gather: thread = edge row; U row + V row -> BN1 affine + LeakyReLU in registers -> tf32 hi / lo
//      halves written straight into shared memory in the K-major 128-byte-swizzle operand layout;
//   2. D[128 edges, C2] = H.W2^T on the tensor cores (tcgen05 kind::tf32, 3xTF32: hi.hi + hi.lo +
//      lo.hi, FP32 accumulators in tensor memory), W2's halves resident in shared memory (TMA, once);
//   3. epilogue: accumulators -> shared memory, then thread = output channel: max and min over the k
//      edges of each point (the one BN2 + LeakyReLU + max will select is known from sign(gamma2)),
//      and the channel's sum / sum of squares over all edges (BatchNorm2 statistics, fp64 at the end).
// Neither [B,2C,N,k] nor [B,C1,N,k] nor [B,C2,N,k] ever exists; the gathered rows come from L2.
#include <cuda.h>
#include <math_constants.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace {

using namespace ecb200::tc;

constexpr int ROWS = 128;       // edge rows per tile (= TMEM lanes)
constexpr int KB = 32;          // channels per K-block (one 128-byte swizzle row)
constexpr int NT = 128;         // one thread per edge row
constexpr int UMMA_K = 8;
constexpr int MAX_C1 = 64, MAX_C2 = 128;

struct Tail {
  uint64_t w_full, mma_done;
  uint32_t tmem_slot;
  float a1[MAX_C1], b1[MAX_C1];
};

// [A hi: C1/32 blocks of 16 KB][A lo][W2 hi: C1/32 blocks of C2*128 B][W2 lo][zs: 128 x (C2+1) floats][tail]
__host__ __device__ constexpr size_t smem_bytes(int C1, int C2) {
  return 1024 + (size_t)2 * ROWS * C1 * 4 + (size_t)2 * C2 * C1 * 4 + (size_t)ROWS * (C2 + 1) * 4 + sizeof(Tail);
}

template <int C2>
__global__ void __launch_bounds__(NT, 1)
two_conv_fwd_kernel(const float* __restrict__ Y, const int32_t* __restrict__ idx, const float* __restrict__ a1g,
                    const float* __restrict__ b1g, float slope1,
                    const __grid_constant__ CUtensorMap map_whi, const __grid_constant__ CUtensorMap map_wlo,
                    const float* __restrict__ gamma2, long long M, int N, int k, int C1,
                    float* __restrict__ sel2, double* __restrict__ stats2) {
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* base = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  const int nkb = C1 / KB;
  unsigned char* a_hi = base;
  unsigned char* a_lo = a_hi + (size_t)ROWS * C1 * 4;
  unsigned char* w_hi = a_lo + (size_t)ROWS * C1 * 4;
  unsigned char* w_lo = w_hi + (size_t)C2 * C1 * 4;
  float* zs = reinterpret_cast<float*>(w_lo + (size_t)C2 * C1 * 4);
  Tail* T = reinterpret_cast<Tail*>(zs + (size_t)ROWS * (C2 + 1));
  constexpr int ZLD = C2 + 1;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int P = ROWS / k;                       // points per tile
  const long long ntiles = (M + P - 1) / P;

  if (tid == 0) {
    prefetch_tensormap(&map_whi);
    prefetch_tensormap(&map_wlo);
    mbar_init(&T->w_full, 1);
    mbar_init(&T->mma_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<128>(&T->tmem_slot);
  for (int c = tid; c < C1; c += NT) { T->a1[c] = a1g[c]; T->b1[c] = b1g[c]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, T->tmem_slot, 0);
  if (warp == 0 && elect_one_sync()) {   // W2 halves: resident for the whole kernel
    const uint32_t wbytes = (uint32_t)(C2 * KB * 4);
    mbar_expect_tx(&T->w_full, 2 * nkb * wbytes);
    for (int kb = 0; kb < nkb; ++kb) {
      tma_load_2d(w_hi + (size_t)kb * wbytes, &map_whi, &T->w_full, kb * KB, 0);
      tma_load_2d(w_lo + (size_t)kb * wbytes, &map_wlo, &T->w_full, kb * KB, 0);
    }
  }
  mbar_wait(&T->w_full, 0);

  // epilogue role of this thread: output channel `col`, points pg, pg+PG, ...
  constexpr int PG = NT / C2;                   // threads per channel (1 for C2 = 128, 2 for 64, 4 for 32)
  const int col = tid % C2, pg = tid / C2;
  const bool take_max = gamma2[col] >= 0.f;
  double acc_s = 0.0, acc_q = 0.0;
  const uint32_t idesc = make_idesc_tf32(ROWS, C2);
  uint32_t phase = 0;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long m0 = tile * P;
    const int pv = (int)(M - m0 < P ? M - m0 : P);      // valid points of this tile
    // ---- 1. gather + BN1 + LeakyReLU + tf32 split -> operand tiles
    {
      const int p = tid / k, j = tid - p * k;
      const bool valid = p < pv;
      const long long m = m0 + p;
      const float *urow = nullptr, *vrow = nullptr;
      if (valid) {
        const long long cloud0 = (m / N) * N;
        const int nb = __ldg(idx + m * k + j);
        urow = Y + (cloud0 + nb) * (2LL * C1);
        vrow = Y + m * (2LL * C1) + C1;
      }
      const uint32_t sw = (uint32_t)(tid & 7);
      for (int kb = 0; kb < nkb; ++kb) {
        float4 u4[8], v4[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          u4[q] = valid ? __ldg(reinterpret_cast<const float4*>(urow + kb * KB) + q) : make_float4(0, 0, 0, 0);
          v4[q] = valid ? __ldg(reinterpret_cast<const float4*>(vrow + kb * KB) + q) : make_float4(0, 0, 0, 0);
        }
        unsigned char* hrow = a_hi + (size_t)kb * (ROWS * KB * 4) + (size_t)tid * 128;
        unsigned char* lrow = a_lo + (size_t)kb * (ROWS * KB * 4) + (size_t)tid * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float* a1 = T->a1 + kb * KB + 4 * q;
          const float* b1 = T->b1 + kb * KB + 4 * q;
          float h[4] = {u4[q].x + v4[q].x, u4[q].y + v4[q].y, u4[q].z + v4[q].z, u4[q].w + v4[q].w};
          float hi[4], lo[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float y = fmaf(a1[e], h[e], b1[e]);
            const float t = valid ? (y > 0.f ? y : y * slope1) : 0.f;
            hi[e] = to_tf32(t);
            lo[e] = to_tf32(t - hi[e]);
          }
          const uint32_t off = ((uint32_t)q ^ sw) * 16;       // 128-byte swizzle: chunk ^= row & 7
          *reinterpret_cast<float4*>(hrow + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(lrow + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    // generic-proxy writes -> visible to the tensor core (async proxy), then one thread issues
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
      if (elect_one_sync()) {
        for (int kb = 0; kb < nkb; ++kb) {
          const uint32_t ah = smem_u32(a_hi + (size_t)kb * (ROWS * KB * 4));
          const uint32_t al = smem_u32(a_lo + (size_t)kb * (ROWS * KB * 4));
          const uint32_t bh = smem_u32(w_hi + (size_t)kb * (C2 * KB * 4));
          const uint32_t bl = smem_u32(w_lo + (size_t)kb * (C2 * KB * 4));
#pragma unroll
          for (int k8 = 0; k8 < KB / UMMA_K; ++k8) {
            const uint32_t o = k8 * UMMA_K * 4;
            mma_tf32(tmem_base, make_sw128_kmajor_desc(ah + o), make_sw128_kmajor_desc(bh + o), idesc, (kb | k8) != 0);
            mma_tf32(tmem_base, make_sw128_kmajor_desc(ah + o), make_sw128_kmajor_desc(bl + o), idesc, 1);
            mma_tf32(tmem_base, make_sw128_kmajor_desc(al + o), make_sw128_kmajor_desc(bh + o), idesc, 1);
          }
        }
        mma_commit(&T->mma_done);
      }
      __syncwarp();
    }
    mbar_wait(&T->mma_done, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- 3a. accumulators -> shared memory (thread = edge row)
    {
      const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
      float* zrow = zs + (size_t)tid * ZLD;
#pragma unroll
      for (int c0 = 0; c0 < C2; c0 += 32) {
        float v[32];
        tmem_ld_32x32(lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int u = 0; u < 32; ++u) zrow[c0 + u] = v[u];
      }
    }
    tc_fence_before();
    __syncthreads();
    // ---- 3b. thread = output channel: max / min over the k edges of each point, statistics
    {
      float s = 0.f, q = 0.f;
      for (int p = pg; p < pv; p += PG) {
        const float* zp = zs + (size_t)(p * k) * ZLD + col;
        float mx = -CUDART_INF_F, mn = CUDART_INF_F;
        for (int j = 0; j < k; ++j) {
          const float z = zp[(size_t)j * ZLD];
          mx = fmaxf(mx, z);
          mn = fminf(mn, z);
          s += z;
          q = fmaf(z, z, q);
        }
        sel2[(m0 + p) * C2 + col] = take_max ? mx : mn;
      }
      acc_s += (double)s;
      acc_q += (double)q;
    }
    __syncthreads();   // zs and the operand tiles are free for the next tile
  }
  if (stats2) {
    atomicAdd(stats2 + col, acc_s);
    atomicAdd(stats2 + C2 + col, acc_q);
    if (blockIdx.x == 0 && tid == 0) atomicAdd(stats2 + 2 * C2, (double)M * (double)k);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem_base);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

int make_w_map(CUtensorMap* m, const float* p, int rows, int C) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<EncodeTiledFn>(fp);
  }
  if (!enc) {
    ecb200::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECB200_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
  const cuuint32_t box[2] = {KB, (cuuint32_t)rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ecb200::set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return ECB200_ERR_CUDA;
  }
  return ECB200_OK;
}

template <int C2>
int launch(const float* Y, const int32_t* idx, const float* a1, const float* b1, float slope1, const float* w2hi,
           const float* w2lo, const float* gamma2, long long M, int N, int k, int C1, float* sel2, double* stats2,
           cudaStream_t st) {
  CUtensorMap mh, ml;
  int rc = make_w_map(&mh, w2hi, C2, C1);
  if (rc) return rc;
  rc = make_w_map(&ml, w2lo, C2, C1);
  if (rc) return rc;
  auto kern = two_conv_fwd_kernel<C2>;
  const size_t smem = smem_bytes(C1, C2);
  static thread_local bool seen[ecb200::kMaxDevices] = {};
  if (ecb200::first_use_on_device(seen))
    ECB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(MAX_C1, C2)));
  const int P = ROWS / k;
  const long long tiles = (M + P - 1) / P;
  const int grid = (int)(tiles < ecb200::kNumSMs ? tiles : ecb200::kNumSMs);   // persistent: one CTA per SM
  kern<<<grid, NT, smem, st>>>(Y, idx, a1, b1, slope1, mh, ml, gamma2, M, N, k, C1, sel2, stats2);
  ECB_LAUNCH_CHECK("two_conv_fwd_kernel");
  return ECB200_OK;
}

}  // namespace

extern "C" int ecb200_two_conv_fwd(const float* Y, const int32_t* idx, const float* a1, const float* b1,
                                   float slope1, const float* w2hi, const float* w2lo, const float* gamma2,
                                   int B, int N, int k, int C1, int C2, float* sel2, double* stats2,
                                   void* stream) {
  ECB_REQUIRE(Y && idx && a1 && b1 && w2hi && w2lo && gamma2 && sel2, "ecb200_two_conv_fwd: null pointer");
  ECB_REQUIRE(B >= 1 && N >= 1 && k >= 1 && k <= ROWS, "ecb200_two_conv_fwd: bad shape B=%d N=%d k=%d", B, N, k);
  ECB_REQUIRE(C1 % KB == 0 && C1 >= KB && C1 <= MAX_C1, "ecb200_two_conv_fwd: C1=%d must be 32 or 64", C1);
  ECB_REQUIRE(C2 == 32 || C2 == 64 || C2 == 128, "ecb200_two_conv_fwd: C2=%d must be 32, 64 or 128", C2);
  const long long M = (long long)B * N;
  cudaStream_t st = (cudaStream_t)stream;
  if (C2 == 128) return launch<128>(Y, idx, a1, b1, slope1, w2hi, w2lo, gamma2, M, N, k, C1, sel2, stats2, st);
  if (C2 == 64) return launch<64>(Y, idx, a1, b1, slope1, w2hi, w2lo, gamma2, M, N, k, C1, sel2, stats2, st);
  return launch<32>(Y, idx, a1, b1, slope1, w2hi, w2lo, gamma2, M, N, k, C1, sel2, stats2, st);
}
