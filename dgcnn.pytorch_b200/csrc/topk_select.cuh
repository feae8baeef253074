// Per-row streaming top-k selector shared by the FMA and tensor-core kNN kernels.
//
// One thread owns one query row.  Candidates whose score passes the row's current
// threshold are appended to a small per-thread buffer in shared memory (cheap, almost
// divergence-free); when any lane of the warp is about to overflow, the whole warp
// flushes: every lane sifts its buffered candidates into its own k-entry min-heap (also
// in shared memory, column `tid` of a [k][NT] array, so accesses are conflict-free).
// Batching the heap updates keeps the lanes of a warp busy together instead of
// serialising one lane's insertion at a time.
//
// Order: 64-bit key = (orderable(score) << 32) | ~j  -- larger score first, ties towards
// the smaller candidate index j.  It is a total order, so the selected set does not
// depend on how the candidates were split between threads or on arrival order.
#pragma once
#include <math_constants.h>
#include <stdint.h>

namespace ecb200 {
namespace topk {

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
// empty slot = (-inf, j = INT_MAX): below every real candidate; key_score() of it is -inf
__device__ __forceinline__ uint64_t empty_key() { return make_key(-CUDART_INF_F, 0x7FFFFFFF); }

// Min-heap of n keys at h[0], h[NT], h[2*NT], ...: place `key` at the root and sift down.
template <int NT>
__device__ __forceinline__ void sift_from_root(uint64_t* h, int n, uint64_t key) {
  int p = 0;
  while (true) {
    int c = 2 * p + 1;
    if (c >= n) break;
    uint64_t kc = h[c * NT];
    if (c + 1 < n) {
      uint64_t k2 = h[(c + 1) * NT];
      if (k2 < kc) { kc = k2; c = c + 1; }
    }
    if (kc >= key) break;
    h[p * NT] = kc;
    p = c;
  }
  h[p * NT] = key;
}

template <int NT, int CAP>
struct RowSelector {
  uint64_t* heap;  // this thread's column of the [k][NT] heap array
  float* bs;       // this thread's column of the [CAP][NT] buffered scores
  int* bj;         //                                   ... and candidate indices
  int k, cnt;
  uint64_t thr_key;
  float thr_s;

  static __host__ __device__ constexpr size_t smem_bytes(int k) {
    return (size_t)NT * ((size_t)k * sizeof(uint64_t) + (size_t)CAP * (sizeof(float) + sizeof(int)));
  }

  __device__ __forceinline__ void init(unsigned char* smem, int k_, int tid) {
    uint64_t* hb = reinterpret_cast<uint64_t*>(smem);
    float* sb = reinterpret_cast<float*>(hb + (size_t)k_ * NT);
    int* jb = reinterpret_cast<int*>(sb + CAP * NT);
    heap = hb + tid;
    bs = sb + tid;
    bj = jb + tid;
    k = k_;
    cnt = 0;
    thr_key = empty_key();
    thr_s = -CUDART_INF_F;
    for (int p = 0; p < k; ++p) heap[p * NT] = thr_key;
  }

  // candidates must be offered in ascending j per thread: ">=" here plus the exact key
  // compare at flush time then realises the (score, smaller-j) order
  __device__ __forceinline__ void offer(float s, int j) {
    if (s >= thr_s) {
      bs[cnt * NT] = s;
      bj[cnt * NT] = j;
      ++cnt;
    }
  }

  __device__ __forceinline__ void insert_key(uint64_t key) {
    if (key > thr_key) {
      sift_from_root<NT>(heap, k, key);
      thr_key = heap[0];
    }
  }

  __device__ __forceinline__ void flush() {
    for (int r = 0; r < cnt; ++r) insert_key(make_key(bs[r * NT], bj[r * NT]));
    cnt = 0;
    thr_s = key_score(thr_key);
  }

  // warp-collective: call from converged code after offering at most CAP - limit candidates
  __device__ __forceinline__ void maybe_flush(int limit) {
    if (__any_sync(0xffffffffu, cnt > limit)) flush();
  }

  // heap sort in place: afterwards heap[0..k-1] is descending (best first)
  __device__ __forceinline__ void sort_descending() {
    for (int n = k - 1; n > 0; --n) {
      const uint64_t last = heap[n * NT];
      heap[n * NT] = heap[0];
      sift_from_root<NT>(heap, n, last);
    }
  }
};

}  // namespace topk
}  // namespace ecb200
