// Two-pass per-row top-k selection shared by the FMA and tensor-core kNN kernels.
//
// Streaming a row's N scores through a k-entry structure costs ~k*ln(N/k) structure
// updates per row (about 100 at N=1024, k=20), each a divergent chain of dependent
// shared-memory accesses: measured on B200 that was 4 warp-instructions per (row,
// candidate) pair and >90 % of the kNN kernel (profiles/r1a_*).  Instead:
//
//   pass A  every thread folds the scores it sees into NB = 32 register bins,
//           bin[u] = max(bin[u], s)  -- one FMNMX per candidate, no branches.  The k-th
//           largest of a row's bin maxima is a lower bound tau on its k-th best score
//           (the k largest bin maxima are k distinct candidates), and a tight one: with
//           64 bins per row and k = 20 about 24 candidates reach it.
//   pass B  the scores are produced again and only candidates with s >= tau are kept:
//           appended to a small per-thread buffer in shared memory.
//   final   exact top-k of the survivors under the total order
//           key = (orderable(score) << 32) | ~j   (larger score first, then smaller j):
//           each survivor's rank among the row's survivors is its output slot.
//
// Scores are cheap to produce twice (3 FMAs per pair for xyz, tensor-core tiles for the
// feature layers); selection work drops to a few instructions per pair.
#pragma once
#include <math_constants.h>
#include <stdint.h>

namespace ecb200 {
namespace topk {

constexpr int NB = 32;  // register bins per thread

__device__ __forceinline__ uint32_t orderable(float s) {
  uint32_t u = __float_as_uint(s);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t make_key(float s, int j) {
  return ((uint64_t)orderable(s) << 32) | (uint32_t)(0xFFFFFFFFu - (uint32_t)j);
}
__device__ __forceinline__ float key_score(uint64_t key) {
  uint32_t u = (uint32_t)(key >> 32);
  u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t key_index(uint64_t key) {
  return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
}

// Bitonic network on 32 registers, descending.  All indices are compile-time constants
// after unrolling: 240 compare-exchanges, two FMNMX each.
__device__ __forceinline__ void sort32_desc(float (&v)[NB]) {
#pragma unroll
  for (int size = 2; size <= NB; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < NB; ++i) {
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

// k-th largest (1-based) of the union of T descending lists of NB floats; list t is the
// column `col0 + t*colstep` of a [NB][stride] shared-memory array.
template <int T>
__device__ __forceinline__ float kth_of_sorted_columns(const float* base, int col0, int colstep,
                                                       int stride, int k) {
  int pos[T];
  float head[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    pos[t] = 0;
    head[t] = base[col0 + t * colstep];
  }
  float val = -CUDART_INF_F;
  for (int it = 0; it < k; ++it) {
    int best = 0;
#pragma unroll
    for (int t = 1; t < T; ++t)
      if (head[t] > head[best]) best = t;
    val = head[best];
#pragma unroll
    for (int t = 0; t < T; ++t) {  // static indexing keeps pos/head in registers
      if (t == best) {
        ++pos[t];
        head[t] = pos[t] < NB ? base[pos[t] * stride + col0 + t * colstep] : -CUDART_INF_F;
      }
    }
  }
  return val;
}

// A thread's survivor list: a column of a [cap][NT] array of 64-bit keys.  cap >= k + the
// number of candidates offered between two guard() calls, so that a thread can always
// fall back to holding its own best k.
template <int NT>
struct Survivors {
  uint64_t* buf;
  int cnt, cap;
  float thr;  // candidates with score >= thr are kept

  static __host__ __device__ constexpr int capacity(int k, int headroom, int floor_) {
    return k + headroom > floor_ ? k + headroom : floor_;
  }
  static __host__ __device__ constexpr size_t smem_bytes(int cap) {
    return (size_t)NT * cap * sizeof(uint64_t);
  }

  __device__ __forceinline__ void init(unsigned char* smem, int tid, int cap_, float tau) {
    buf = reinterpret_cast<uint64_t*>(smem) + tid;
    cnt = 0;
    cap = cap_;
    thr = tau;
  }
  // candidates must arrive in ascending j within a thread
  __device__ __forceinline__ void offer(float s, int j) {
    if (s >= thr) {
      buf[cnt * NT] = make_key(s, j);
      ++cnt;
    }
  }
  // Slow path, only when ties or clustered data overfill the buffer: keep the thread's own
  // best k and raise its threshold to "strictly better than the k-th kept score" (later
  // candidates of equal score have a larger j, hence a smaller key, and can be dropped).
  __device__ void shrink_to(int k) {
    while (cnt > k) {
      int arg = 0;
      uint64_t mn = buf[0];
      for (int e = 1; e < cnt; ++e) {
        const uint64_t v = buf[e * NT];
        if (v < mn) { mn = v; arg = e; }
      }
      --cnt;
      buf[arg * NT] = buf[cnt * NT];
    }
    uint64_t mn = buf[0];
    for (int e = 1; e < cnt; ++e) mn = min(mn, buf[e * NT]);
    thr = fmaxf(thr, nextafterf(key_score(mn), CUDART_INF_F));
  }
  __device__ __forceinline__ void guard(int k, int headroom) {
    if (cnt > cap - headroom) shrink_to(k);
  }
};

// Exact, sorted top-k of a row whose survivors sit in T columns (column t = col0 +
// t*colstep of the [cap][NT] key array, holding cnts[t] keys).  Every one of the row's T
// threads ranks its OWN survivors against the whole union -- rank = number of strictly
// greater keys; keys are distinct, so ranks are a permutation -- and writes those with
// rank < k straight to out[rank].  All loads are independent (no serial selection chain),
// and the result comes out sorted best-first for free.
template <int NT, int T>
__device__ __forceinline__ void rank_and_write(const uint64_t* keys, int col0, int colstep,
                                               const int (&cnts)[T], int me, int k, int n_clamp,
                                               int32_t* __restrict__ out) {
  const uint64_t* mine = keys + col0 + me * colstep;
  for (int e = 0; e < cnts[me]; ++e) {
    const uint64_t key = mine[e * NT];
    int rank = 0;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const uint64_t* col = keys + col0 + t * colstep;
      for (int f = 0; f < cnts[t]; ++f) rank += (col[f * NT] > key);
    }
    if (rank < k) out[rank] = (int32_t)min(key_index(key), (uint32_t)n_clamp);
  }
}

}  // namespace topk
}  // namespace ecb200
