// SyncBatchNorm statistics exchange over NVLink peer memory, one kernel per exchange.
//
// The only collective on the EdgeConv path is the all-reduce(SUM) of a few hundred fp64 values
// per BatchNorm layer ([sum e | sum e^2 | count] forward, [sum g | sum g*xhat] backward;
// main_partseg_dist.py:189 -> torch SyncBatchNorm).  At 0.5-16 KB it is pure latency, so instead
// of a library all-reduce every rank PUSHES its vector into a slot of every peer's symmetric
// buffer with plain stores over NVLink and sums the `world` vectors that arrive in its own buffer
// in rank order (bit-identical on every rank).
//
// Low-latency protocol (as NCCL's LL): no separate flag, no fence.  Every fp64 value travels as
// two 8-byte words {32 data bits | 32-bit sequence number}; an aligned 8-byte store is delivered
// atomically, so a receiver that sees the expected sequence number in a word also sees its
// data.  One NVLink one-way latency per exchange instead of two round trips.
//
// Buffer layout (identical on every rank, zero-initialised once):
//   [2 parities][world][ECB200_PEER_MAX_VALUES][2] uint64
// Consecutive exchanges alternate parity.  Reuse of a parity half needs no extra barrier: a rank
// can only reach exchange s+2 after every peer has written s+1, which each peer does (stream
// order) only after it finished reading s; stale words carry an older sequence number.  `seq`
// is a device-resident counter, so the kernel replays correctly inside a CUDA graph.
#include "common.cuh"

namespace {

__device__ __forceinline__ void st_volatile_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512)
peer_allreduce_kernel(double* __restrict__ vals, int n, void* const* __restrict__ peer_bufs, int rank,
                      int world, unsigned long long* __restrict__ seq_counter) {
  constexpr size_t MAXV = ECB200_PEER_MAX_VALUES;
  const unsigned long long seq = *seq_counter + 1;
  const unsigned long long tag = (seq & 0xffffffffull) << 32;
  const int par = (int)(seq & 1ull);
  // 1. push my vector into slot [par][rank] of every rank's buffer (mine included)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long bits = (unsigned long long)__double_as_longlong(vals[i]);
    const unsigned long long w0 = (bits & 0xffffffffull) | tag, w1 = (bits >> 32) | tag;
    for (int p = 0; p < world; ++p) {
      unsigned long long* dst = reinterpret_cast<unsigned long long*>(peer_bufs[p]) +
                                (((size_t)par * world + rank) * MAXV + i) * 2;
      st_volatile_u64(dst, w0);
      st_volatile_u64(dst + 1, w1);
    }
  }
  // 2. collect: a word is valid once it carries this exchange's sequence number; sum in rank order
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(peer_bufs[rank]) +
                                  (size_t)par * world * MAXV * 2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < world; ++q) {
      const unsigned long long* w = src + ((size_t)q * MAXV + i) * 2;
      unsigned long long w0, w1;
      while (((w0 = ld_volatile_u64(w)) & 0xffffffff00000000ull) != tag) __nanosleep(20);
      while (((w1 = ld_volatile_u64(w + 1)) & 0xffffffff00000000ull) != tag) __nanosleep(20);
      acc += __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
    }
    vals[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *seq_counter = seq;
}

}  // namespace

extern "C" size_t ecb200_peer_buffer_bytes(int world) {
  return (size_t)2 * world * ECB200_PEER_MAX_VALUES * 2 * sizeof(unsigned long long);
}

extern "C" int ecb200_peer_allreduce(double* vals, int n, void* const* peer_bufs, int rank, int world,
                                     unsigned long long* seq_counter, void* stream) {
  ECB_REQUIRE(vals && peer_bufs && seq_counter, "ecb200_peer_allreduce: null pointer");
  ECB_REQUIRE(n >= 1 && n <= ECB200_PEER_MAX_VALUES, "ecb200_peer_allreduce: n=%d outside [1, %d]", n,
              ECB200_PEER_MAX_VALUES);
  ECB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "ecb200_peer_allreduce: bad rank/world");
  peer_allreduce_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(vals, n, peer_bufs, rank, world, seq_counter);
  ECB_LAUNCH_CHECK("peer_allreduce_kernel");
  return ECB200_OK;
}
