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

// What the exchange feeds.  Every exchange of the path is followed by a tiny per-channel "finalize"
// (BatchNorm affine in the forward, the c1/c2 coefficients in the backward); doing it in the same
// kernel saves one launch per exchange -- twelve per training step.
struct Finalize {
  int kind;            // 0 none, 1 forward (ecb200_bn_finalize, training), 2 backward (ecb200_bwd_finalize, training)
  int Co;
  float eps;
  const float *gamma, *beta;           // kind 1
  float *mean, *invstd, *a, *b;        // kind 1 outputs
  const double *local, *count;         // kind 2: this rank's [sum g | sum g*xhat]; global edge count
  const float *a_in, *invstd_in;       // kind 2
  float *dgamma, *dbeta, *c1, *c2;     // kind 2 outputs
};

__global__ void __launch_bounds__(1024)
peer_allreduce_kernel(double* __restrict__ vals, int n, void* const* __restrict__ peer_bufs, int rank,
                      int world, unsigned long long* __restrict__ seq_counter, Finalize fin) {
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
  // (the words of up to 8 peers are loaded before any of them is checked: one memory round trip per
  // group instead of one per word -- at 8 ranks the serial form cost 16 dependent loads per value)
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int q0 = 0; q0 < world; q0 += 8) {
      unsigned long long w0[8], w1[8];
      bool pending = true;
      unsigned spins = 0;
      while (pending) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (q0 + u < world) {
            const unsigned long long* w = src + ((size_t)(q0 + u) * MAXV + i) * 2;
            w0[u] = ld_volatile_u64(w);
            w1[u] = ld_volatile_u64(w + 1);
          }
        pending = false;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (q0 + u < world)
            pending |= ((w0[u] & 0xffffffff00000000ull) != tag) | ((w1[u] & 0xffffffff00000000ull) != tag);
        if (pending) {
          __nanosleep(20);
          // a peer that never arrives (crashed rank) must not hang the GPU: after ~2^22 polls (seconds)
          // give up and poison the value -- the step's loss turns NaN instead of the device wedging
          if (++spins > (1u << 22)) {
            acc = __longlong_as_double(0x7ff8000000000000LL);
            break;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u)   // rank order: bit-identical sums on every rank
        if (q0 + u < world) acc += __longlong_as_double((long long)((w0[u] & 0xffffffffull) | (w1[u] << 32)));
    }
    vals[i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) *seq_counter = seq;
  if (fin.kind == 1) {          // == bn_finalize_kernel (training) on the reduced statistics
    const double count = vals[2 * fin.Co];
    for (int o = threadIdx.x; o < fin.Co; o += blockDim.x) {
      const double mu = vals[o] / count;
      double var = vals[fin.Co + o] / count - mu * mu;
      if (var < 0.0) var = 0.0;
      const double r = 1.0 / sqrt(var + (double)fin.eps);
      const double aa = (double)fin.gamma[o] * r;
      fin.mean[o] = (float)mu;
      fin.invstd[o] = (float)r;
      fin.a[o] = (float)aa;
      fin.b[o] = (float)((double)fin.beta[o] - aa * mu);
    }
  } else if (fin.kind == 2) {   // == bwd_finalize_kernel (training): vals = the reduced [sum g | sum g*xhat]
    const double count = *fin.count;
    for (int o = threadIdx.x; o < fin.Co; o += blockDim.x) {
      fin.dbeta[o] = (float)fin.local[o];
      fin.dgamma[o] = (float)fin.local[fin.Co + o];
      fin.c1[o] = (float)((double)fin.a_in[o] * vals[o] / count);
      fin.c2[o] = (float)((double)fin.a_in[o] * vals[fin.Co + o] * (double)fin.invstd_in[o] / count);
    }
  }
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
  Finalize fin = {};
  peer_allreduce_kernel<<<1, n > 512 ? 1024 : 512, 0, (cudaStream_t)stream>>>(vals, n, peer_bufs, rank, world, seq_counter, fin);
  ECB_LAUNCH_CHECK("peer_allreduce_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_peer_allreduce_bn_finalize(double* stats, int Co, void* const* peer_bufs, int rank,
                                                 int world, unsigned long long* seq_counter,
                                                 const float* gamma, const float* beta, float eps, float* mean,
                                                 float* invstd, float* a, float* b, void* stream) {
  ECB_REQUIRE(stats && peer_bufs && seq_counter && gamma && beta && mean && invstd && a && b,
              "ecb200_peer_allreduce_bn_finalize: null pointer");
  ECB_REQUIRE(Co >= 1 && 2 * Co + 1 <= ECB200_PEER_MAX_VALUES, "ecb200_peer_allreduce_bn_finalize: Co=%d", Co);
  ECB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "ecb200_peer_allreduce_bn_finalize: bad rank/world");
  Finalize fin = {};
  fin.kind = 1; fin.Co = Co; fin.eps = eps; fin.gamma = gamma; fin.beta = beta;
  fin.mean = mean; fin.invstd = invstd; fin.a = a; fin.b = b;
  peer_allreduce_kernel<<<1, Co > 256 ? 1024 : 512, 0, (cudaStream_t)stream>>>(stats, 2 * Co + 1, peer_bufs, rank, world, seq_counter, fin);
  ECB_LAUNCH_CHECK("peer_allreduce_kernel<bn_finalize>");
  return ECB200_OK;
}

extern "C" int ecb200_peer_allreduce_bwd_finalize(const double* bstats_local, double* bstats_global, int Co,
                                                  void* const* peer_bufs, int rank, int world,
                                                  unsigned long long* seq_counter, const double* count_dev,
                                                  const float* a, const float* invstd, float* dgamma, float* dbeta,
                                                  float* c1, float* c2, void* stream) {
  ECB_REQUIRE(bstats_local && bstats_global && peer_bufs && seq_counter && count_dev && a && invstd && dgamma &&
                  dbeta && c1 && c2, "ecb200_peer_allreduce_bwd_finalize: null pointer");
  ECB_REQUIRE(Co >= 1 && 2 * Co <= ECB200_PEER_MAX_VALUES, "ecb200_peer_allreduce_bwd_finalize: Co=%d", Co);
  ECB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world, "ecb200_peer_allreduce_bwd_finalize: bad rank/world");
  Finalize fin = {};
  fin.kind = 2; fin.Co = Co; fin.local = bstats_local; fin.count = count_dev; fin.a_in = a; fin.invstd_in = invstd;
  fin.dgamma = dgamma; fin.dbeta = dbeta; fin.c1 = c1; fin.c2 = c2;
  peer_allreduce_kernel<<<1, Co > 256 ? 1024 : 512, 0, (cudaStream_t)stream>>>(bstats_global, 2 * Co, peer_bufs, rank, world, seq_counter, fin);
  ECB_LAUNCH_CHECK("peer_allreduce_kernel<bwd_finalize>");
  return ECB200_OK;
}
