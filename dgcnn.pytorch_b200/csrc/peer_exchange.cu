// SyncBatchNorm statistics exchange over NVLink peer memory, one kernel per exchange.
//
// The only collective on the EdgeConv path is the all-reduce(SUM) of a few hundred fp64 values
// per BatchNorm layer ([sum e | sum e^2 | count] forward, [sum g | sum g*xhat] backward;
// main_partseg_dist.py:189 -> torch SyncBatchNorm).  At 0.5-16 KB it is pure latency, so instead
// of a library all-reduce every rank PUSHES its vector into a slot of every peer's symmetric
// buffer with plain stores over NVLink, raises a flag there, waits for the flags of all peers in
// its own buffer and sums the `world` vectors in rank order (bit-identical on every rank).
//
// Buffer layout (identical on every rank, zero-initialised once):
//   data  : [2 parities][world][ECB200_PEER_MAX_VALUES] double
//   flags : [2 parities][world] uint64      at byte offset 2*world*MAX*8
// Consecutive exchanges alternate parity.  Reuse of a parity half needs no extra barrier: a rank
// can only reach exchange s+2 after every peer has written s+1, which each peer does (stream
// order) only after it finished reading s.  `seq` is a device-resident counter, so the kernel
// replays correctly inside a CUDA graph.
#include "common.cuh"

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(512)
peer_allreduce_kernel(double* __restrict__ vals, int n, void* const* __restrict__ peer_bufs, int rank,
                      int world, unsigned long long* __restrict__ seq_counter) {
  constexpr int MAXV = ECB200_PEER_MAX_VALUES;
  const unsigned long long seq = *seq_counter + 1;
  const int par = (int)(seq & 1ull);
  const size_t flag_off = (size_t)2 * world * MAXV * sizeof(double);
  // 1. push my vector into slot [par][rank] of every rank's buffer (mine included)
  for (int p = 0; p < world; ++p) {
    double* dst = reinterpret_cast<double*>(peer_bufs[p]) + ((size_t)par * world + rank) * MAXV;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = vals[i];
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise my flag at every rank, 3. wait for every rank's flag here
  if (threadIdx.x < world) {
    unsigned long long* theirs = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<unsigned char*>(peer_bufs[threadIdx.x]) + flag_off) + (size_t)par * world + rank;
    st_release_sys(theirs, seq);
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(
        reinterpret_cast<const unsigned char*>(peer_bufs[rank]) + flag_off) + (size_t)par * world + threadIdx.x;
    while (ld_acquire_sys(mine) < seq) __nanosleep(32);
  }
  __syncthreads();
  // 4. sum in rank order; .cv loads: these addresses were read two exchanges ago, L1 may be stale
  const double* src = reinterpret_cast<const double*>(peer_bufs[rank]) + (size_t)par * world * MAXV;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int q = 0; q < world; ++q) acc += __ldcv(src + (size_t)q * MAXV + i);
    vals[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *seq_counter = seq;
}

}  // namespace

extern "C" size_t ecb200_peer_buffer_bytes(int world) {
  return (size_t)2 * world * ECB200_PEER_MAX_VALUES * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
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
