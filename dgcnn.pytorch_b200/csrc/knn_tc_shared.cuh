// Pieces shared by the tensor-core kNN kernels (knn_tc.cu: 128 query rows per CTA; knn_tc2.cu: 256).
#pragma once
#include <cuda.h>
#include <math_constants.h>
#include <stdint.h>

#include "common.cuh"
#include "tc_ptx.cuh"
#include "topk_select.cuh"

namespace ecb200 {
namespace knntc {

using namespace ecb200::tc;
using namespace ecb200::topk;

constexpr int KB = 32;   // 32-bit words of K per block: one 128-byte swizzle row

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

// raw survivor entry (score bits << 32 | j) -> totally ordered key (larger score, then smaller j)
__device__ __forceinline__ uint64_t ordered_key(uint64_t raw) {
  return make_key(__uint_as_float((uint32_t)(raw >> 32)), (int)(uint32_t)raw);
}
// Overflow of a thread's survivor list (only with massive ties or clustered data): keep its own
// best k entries in place and return the score a later candidate must reach to matter ("strictly
// better than the k-th kept": later candidates of equal score have a larger j, hence a smaller
// key).  Out of line and not unrolled: it must not bloat the hot loop's instruction footprint.
static __device__ __noinline__ float shrink_survivors(uint64_t* buf, int cnt, int k, int LS) {
#pragma unroll 1
  while (cnt > k) {
    int arg = 0;
    uint64_t mn = ordered_key(buf[0]);
#pragma unroll 1
    for (int e = 1; e < cnt; ++e) {
      const uint64_t w = ordered_key(buf[e * LS]);
      if (w < mn) { mn = w; arg = e; }
    }
    --cnt;
    buf[arg * LS] = buf[cnt * LS];
  }
  uint64_t mn = ordered_key(buf[0]);
#pragma unroll 1
  for (int e = 1; e < cnt; ++e) mn = min(mn, ordered_key(buf[e * LS]));
  return nextafterf(key_score(mn), CUDART_INF_F);
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
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

// [rows, C] fp32 row-major, box = 32 channels x box_rows rows, 128-byte swizzle, zero fill past the end
inline int make_point_map(CUtensorMap* m, const float* p, long long rows, int C, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    ecb200::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECB200_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)C * sizeof(float)};
  const cuuint32_t box[2] = {KB, (cuuint32_t)box_rows};
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

// [clouds, N, C] fp32 row-major as a 3-D map, box = 32 channels x box_rows rows of ONE cloud, 128-byte swizzle;
// rows past the end of a cloud read as NaN: a candidate that does not exist scores NaN, which no
// comparison of the selector accepts (fmaxf drops it, >= is false) -- no masking code in the epilogue
inline int make_cloud_map(CUtensorMap* m, const float* p, int clouds, int N, int C, int box_rows) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) {
    ecb200::set_error("cuTensorMapEncodeTiled is not available from this driver");
    return ECB200_ERR_CUDA;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)clouds};
  const cuuint64_t strides[2] = {(cuuint64_t)C * sizeof(float), (cuuint64_t)N * C * sizeof(float)};
  const cuuint32_t box[3] = {KB, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(p), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
  if (r != CUDA_SUCCESS) {
    ecb200::set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
    return ECB200_ERR_CUDA;
  }
  return ECB200_OK;
}

// [rows, C] fp32 row-major operand pair (hi, lo) as TMA maps
struct OperandMaps {
  CUtensorMap hi, lo;
};
inline int make_operand(OperandMaps* m, const float* hi, const float* lo, long long rows, int C, int box_rows) {
  int rc = make_point_map(&m->hi, hi, rows, C, box_rows);
  if (rc) return rc;
  return make_point_map(&m->lo, lo, rows, C, box_rows);
}


// 256-row variant (knn_tc2.cu): packed-FP16 operands, Cw = 32-bit words per operand row
struct Tc2Args {
  const float *a_hi, *a_lo, *b_hi, *b_lo, *bn, *xx, *cmax;   // bn: norm rows (TERMS = 3) or NULL
  long long b_rows;
  int clouds, Cw, N, k, ksteps;
  int32_t* idx;
  long long* tl;
};
bool tc2_takes(int Cw, int N, int k, int terms);
int launch_knn_tc2(const Tc2Args& a, int terms, cudaStream_t st);

}  // namespace knntc
}  // namespace ecb200
