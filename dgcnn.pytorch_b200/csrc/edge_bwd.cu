// EdgeConv backward without the edge tensor.
//
// The reference's autograd walks max -> LeakyReLU -> BatchNorm2d -> Conv2d -> cat/repeat
// -> index over [B,Co,N,k] / [B,2C,N,k] tensors.  With e_ij = U_j + V_i,
// g_io = dL/dout_io * LeakyReLU'(a*sel+b) placed at the arg slot j*(i,o), mu/r the batch
// mean / inverse std, a = gamma*r, Mtot = B*N*k:
//   dbeta = sum g           dgamma = sum g*(sel-mu)*r
//   c1 = a*dbeta/Mtot       c2 = a*dgamma*r/Mtot
//   dV_i = a*g_i - k*c1 - c2*(sum_j e_ij - k*mu)
//   dU_j = sum_{(i,o): idx[i,j*]=j} a*g_i  -  c1*deg_j - c2*(deg_j*(U_j-mu) + T_j)
// with deg_j the in-degree of j and T_j = sum of V_i over the in-edges of j, which needs
// the destination-major (reverse) graph built here.  In eval mode c1 = c2 = 0.
#include "common.cuh"

namespace {

constexpr int BT = 256;

__global__ void __launch_bounds__(256)
bwd_prep_kernel(const float* __restrict__ gout, const float* __restrict__ gout_pm, long long ld_pm,
                const float* __restrict__ sel,
                const float* __restrict__ a, const float* __restrict__ b,
                const float* __restrict__ mean, const float* __restrict__ invstd, float slope, int N,
                int Co, float* __restrict__ g, double* __restrict__ bstats) {
  __shared__ float tile[32][33];
  __shared__ double red[2][8][32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int n0 = blockIdx.x * 32, o0 = blockIdx.y * 32, bb = blockIdx.z;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int oc = o0 + ty + 8 * r, n = n0 + tx;
    tile[ty + 8 * r][tx] = (gout && oc < Co && n < N) ? gout[((size_t)bb * Co + oc) * N + n] : 0.f;
  }
  __syncthreads();
  const int o = o0 + tx;
  double db = 0.0, dg = 0.0;
  if (o < Co) {
    const float ao = a[o], bo = b[o], mu = mean[o], rs = invstd[o];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int p = ty + 8 * r, n = n0 + p;
      if (n < N) {
        const size_t m = (size_t)bb * N + n;
        const float s = sel[m * Co + o];
        const float y = fmaf(ao, s, bo);
        float gin = tile[tx][p];
        if (gout_pm) gin += gout_pm[m * ld_pm + o];  // gradient of the point-major copy of the output
        const float gg = gin * (y > 0.f ? 1.f : slope);
        g[m * Co + o] = gg;
        db += (double)gg;
        dg += (double)(gg * ((s - mu) * rs));
      }
    }
  }
  red[0][ty][tx] = db;
  red[1][ty][tx] = dg;
  __syncthreads();
  if (ty < 2 && o < Co) {
    double s = 0.0;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[ty][r][tx];
    atomicAdd(&bstats[ty * Co + o], s);
  }
}

__global__ void bwd_finalize_kernel(const double* __restrict__ local, const double* __restrict__ glob,
                                    const double* __restrict__ count_dev, const float* __restrict__ a,
                                    const float* __restrict__ invstd, int training, int Co,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta,
                                    float* __restrict__ c1, float* __restrict__ c2) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= Co) return;
  dbeta[o] = (float)local[o];
  dgamma[o] = (float)local[Co + o];
  if (training) {
    const double count = *count_dev;
    c1[o] = (float)((double)a[o] * glob[o] / count);
    c2[o] = (float)((double)a[o] * glob[Co + o] * (double)invstd[o] / count);
  } else {
    c1[o] = 0.f;
    c2[o] = 0.f;
  }
}

// ---- reverse graph -----------------------------------------------------------------
__global__ void rev_count_kernel(const int32_t* __restrict__ idx, int N, int k, long long E,
                                 int32_t* __restrict__ cursor) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long base = (e / ((long long)N * k)) * N;
  atomicAdd(&cursor[base + idx[e]], 1);
}

__global__ void __launch_bounds__(1024)
rev_scan_kernel(const int32_t* __restrict__ deg, int N, int k, long long M,
                int32_t* __restrict__ rowptr) {
  __shared__ int wsum[32];
  const int bb = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int32_t* d = deg + (size_t)bb * N;
  int32_t* rp = rowptr + (size_t)bb * N;
  const int per = (N + blockDim.x - 1) / blockDim.x;
  const int beg = min(N, tid * per), end = min(N, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += d[i];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    wsum[lane] = v;  // inclusive over warps
  }
  __syncthreads();
  int run = (int)((long long)bb * N * k) + (inc - s) + (w > 0 ? wsum[w - 1] : 0);
  for (int i = beg; i < end; ++i) {
    rp[i] = run;
    run += d[i];
  }
  if (bb == (int)gridDim.x - 1 && tid == 0) rowptr[M] = (int32_t)(M * k);
}

__global__ void rev_fill_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ rowptr,
                                int N, int k, long long E, int32_t* __restrict__ cursor,
                                int32_t* __restrict__ src) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const long long base = (e / ((long long)N * k)) * N;
  const long long dst = base + idx[e];
  const int pos = atomicSub(&cursor[dst], 1) - 1;
  src[rowptr[dst] + pos] = (int32_t)(e / k);
}

// Shared-memory reverse graph, small clouds (N*k < 40960): CTA (s, b) owns the destination points [lo, hi) of cloud b.  It
// reads the cloud's whole neighbour list twice (L2-resident): once to histogram the in-degrees of
// its destinations (shared-memory atomics) and to count the edges that go to smaller
// destinations (its global base offset), once to fill.  Replaces the count / scan / fill triple
// (global atomics, three launches and a memset) whenever the counters fit shared memory.
__global__ void __launch_bounds__(1024)
rev_cloud_rows_kernel(const int32_t* __restrict__ idx, int N, int k, long long M, int span,
                 int32_t* __restrict__ rowptr, int32_t* __restrict__ src) {
  extern __shared__ int sh[];          // [span] degree -> cursor, [span] exclusive offsets
  int* deg = sh;
  int* off = sh + span;
  __shared__ int wsum[32];
  __shared__ int below_s;
  const int bb = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int lo = blockIdx.x * span, hi = min(N, lo + span), cnt = hi - lo;
  const int E = N * k;
  const int32_t* ib = idx + (size_t)bb * E;
  for (int i = tid; i < cnt; i += blockDim.x) deg[i] = 0;
  if (tid == 0) below_s = 0;
  __syncthreads();
  int below = 0;
  for (int i = tid; i < N; i += blockDim.x) {   // thread = source point: no index division
    const int32_t* row = ib + (size_t)i * k;
#pragma unroll 4
    for (int j = 0; j < k; ++j) {
      const int d = row[j];
      if (d < lo) ++below;
      else if (d < hi) atomicAdd(&deg[d - lo], 1);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
  if (lane == 0 && below) atomicAdd(&below_s, below);
  __syncthreads();
  // exclusive scan of deg[0..cnt): each thread owns a contiguous chunk
  const int per = (cnt + blockDim.x - 1) / blockDim.x;
  const int beg = min(cnt, tid * per), end = min(cnt, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += deg[i];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    wsum[lane] = v;
  }
  __syncthreads();
  const int gbase = (int)((long long)bb * E) + below_s;   // global offset of destination `lo`
  int run = (inc - s) + (w > 0 ? wsum[w - 1] : 0);
  for (int i = beg; i < end; ++i) {
    off[i] = run;
    rowptr[(size_t)bb * N + lo + i] = gbase + run;
    run += deg[i];
  }
  if (bb == (int)gridDim.y - 1 && blockIdx.x == gridDim.x - 1 && tid == 0) rowptr[M] = (int32_t)(M * k);
  __syncthreads();
  // fill: deg[] now counts down as the cursor of each destination row
  for (int i = tid; i < N; i += blockDim.x) {
    const int32_t* row = ib + (size_t)i * k;
    const int32_t me = (int32_t)((long long)bb * N + i);
#pragma unroll 4
    for (int j = 0; j < k; ++j) {
      const int d = row[j];
      if (d >= lo && d < hi) {
        const int pos = atomicSub(&deg[d - lo], 1) - 1;
        src[gbase + off[d - lo] + pos] = me;
      }
    }
  }
}

// The same for larger clouds (measured on B200: 301 vs 426 us per step at N = 2048, k = 40, but 82 vs 61 us at
// N = 1024, k = 20, where the thread-per-source-point form above stays).  CTA (s, b) owns the destination
// points [lo, hi) of cloud b.  It
// reads the cloud's whole neighbour list twice (L2-resident, as one flat coalesced stream): once to
// histogram the in-degrees of its destinations (shared-memory atomics) and to count the edges that
// go to smaller destinations (its global base offset), once to fill.  The CTA's slice of `src` is
// contiguous (CSR rows lo..hi): it is assembled in shared memory and written out as one coalesced
// block when it fits (`seg_cap` entries), instead of one scattered 4-byte store per edge.  Replaces
// the count / scan / fill triple (global atomics, three launches and a memset) whenever the
// counters fit shared memory.
__global__ void __launch_bounds__(1024)
rev_cloud_kernel(const int32_t* __restrict__ idx, int N, int k, long long M, int span, int seg_cap,
                 int32_t* __restrict__ rowptr, int32_t* __restrict__ src) {
  extern __shared__ int sh[];          // [span] degree -> cursor, [span] exclusive offsets, [seg_cap] segment
  int* deg = sh;
  int* off = sh + span;
  int* seg = sh + 2 * span;
  __shared__ int wsum[32];
  __shared__ int below_s;
  const int bb = blockIdx.y, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int lo = blockIdx.x * span, hi = min(N, lo + span), cnt = hi - lo;
  const int E = N * k;
  const int32_t* ib = idx + (size_t)bb * E;
  const unsigned inv_k = 0xffffffffu / (unsigned)k + 1u;   // e / k = (e * inv_k) >> 32, exact for e < 2^20 * ... (E < 2^26, k <= 64)
  for (int i = tid; i < cnt; i += blockDim.x) deg[i] = 0;
  if (tid == 0) below_s = 0;
  __syncthreads();
  int below = 0;
#pragma unroll 4
  for (int e = tid; e < E; e += blockDim.x) {
    const int d = __ldg(ib + e);
    if (d < lo) ++below;
    else if (d < hi) atomicAdd(&deg[d - lo], 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
  if (lane == 0 && below) atomicAdd(&below_s, below);
  __syncthreads();
  // exclusive scan of deg[0..cnt): each thread owns a contiguous chunk
  const int per = (cnt + blockDim.x - 1) / blockDim.x;
  const int beg = min(cnt, tid * per), end = min(cnt, beg + per);
  int s = 0;
  for (int i = beg; i < end; ++i) s += deg[i];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? wsum[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += t;
    }
    wsum[lane] = v;
  }
  __syncthreads();
  const int gbase = (int)((long long)bb * E) + below_s;   // global offset of destination `lo`
  const int total = wsum[(blockDim.x >> 5) - 1];           // edges into [lo, hi)
  int run = (inc - s) + (w > 0 ? wsum[w - 1] : 0);
  for (int i = beg; i < end; ++i) {
    off[i] = run;
    rowptr[(size_t)bb * N + lo + i] = gbase + run;
    run += deg[i];
  }
  if (bb == (int)gridDim.y - 1 && blockIdx.x == gridDim.x - 1 && tid == 0) rowptr[M] = (int32_t)(M * k);
  __syncthreads();
  // fill: deg[] now counts down as the cursor of each destination row
  const bool staged = total <= seg_cap;
  const int me0 = (int)((long long)bb * N);
#pragma unroll 4
  for (int e = tid; e < E; e += blockDim.x) {
    const int d = __ldg(ib + e);
    if (d >= lo && d < hi) {
      const int pos = atomicSub(&deg[d - lo], 1) - 1;
      const int32_t me = me0 + (int)(((unsigned long long)(unsigned)e * inv_k) >> 32);
      if (staged) seg[off[d - lo] + pos] = me;
      else        src[gbase + off[d - lo] + pos] = me;
    }
  }
  if (staged) {
    __syncthreads();
    for (int i = tid; i < total; i += blockDim.x) src[gbase + i] = seg[i];
  }
}

// ---- dense BatchNorm terms of dU through the reverse graph ---------------------------
template <int LPP>
__global__ void __launch_bounds__(BT)
bwd_dense_kernel(const float* __restrict__ Y, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ src, const float* __restrict__ mean,
                 const float* __restrict__ c1, const float* __restrict__ c2, int training, int Co,
                 long long M, float* __restrict__ dY, const float* __restrict__ dU_in,
                 float* __restrict__ dYhi, float* __restrict__ dYlo) {
  constexpr int SPW = 32 / LPP;
  const int lane = threadIdx.x & 31, sl = lane % LPP;
  const int G = Co / (4 * LPP), Co2 = 2 * Co;
  const long long S = (long long)gridDim.x * (BT / 32) * SPW;
  const long long s = ((long long)blockIdx.x * (BT / 32) + threadIdx.x / 32) * SPW + lane / LPP;
  const long long PS = S / G;
  const int c = (int)(s % G) * 4 * LPP + sl * 4;
  const long long ps = s / G;
  if (ps >= PS) return;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c);
  const float4 k1 = *reinterpret_cast<const float4*>(c1 + c);
  const float4 k2 = *reinterpret_cast<const float4*>(c2 + c);
  for (long long m = ps; m < M; m += PS) {
    float4 out = make_float4(0.f, 0.f, 0.f, 0.f);
    if (training) {
      const int beg = rowptr[m], end = rowptr[m + 1];
      float4 T = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
      for (int e = beg; e < end; ++e) {
        const long long sp = __ldg(src + e);
        const float4 v = *reinterpret_cast<const float4*>(Y + sp * Co2 + Co + c);
        T.x += v.x; T.y += v.y; T.z += v.z; T.w += v.w;
      }
      const float deg = (float)(end - beg);
      const float4 u = *reinterpret_cast<const float4*>(Y + m * Co2 + c);
      out.x = -k1.x * deg - k2.x * (deg * (u.x - mu.x) + T.x);
      out.y = -k1.y * deg - k2.y * (deg * (u.y - mu.y) + T.y);
      out.z = -k1.z * deg - k2.z * (deg * (u.z - mu.z) + T.z);
      out.w = -k1.w * deg - k2.w * (deg * (u.w - mu.w) + T.w);
    }
    if (dU_in) {
      // fused mode: the sparse part was scattered into dU_in [M,Co] before; the sum goes out as the
      // tf32 hi/lo operand pair of the tensor-core GEMMs (no fp32 dY, no separate split pass)
      const float4 sp = *reinterpret_cast<const float4*>(dU_in + m * Co + c);
      out.x += sp.x; out.y += sp.y; out.z += sp.z; out.w += sp.w;
      float4 h, l;
      h.x = ecb200::tf32_rna(out.x); l.x = ecb200::tf32_rna(out.x - h.x);
      h.y = ecb200::tf32_rna(out.y); l.y = ecb200::tf32_rna(out.y - h.y);
      h.z = ecb200::tf32_rna(out.z); l.z = ecb200::tf32_rna(out.z - h.z);
      h.w = ecb200::tf32_rna(out.w); l.w = ecb200::tf32_rna(out.w - h.w);
      *reinterpret_cast<float4*>(dYhi + m * Co2 + c) = h;
      *reinterpret_cast<float4*>(dYlo + m * Co2 + c) = l;
    } else {
      *reinterpret_cast<float4*>(dY + m * Co2 + c) = out;
    }
  }
}

// ---- dV and the sparse scatter of a*g into dU ----------------------------------------
template <int LPP>
__global__ void __launch_bounds__(BT)
bwd_scatter_kernel(const float* __restrict__ g, const float* __restrict__ esum,
                   const uint8_t* __restrict__ arg, const int32_t* __restrict__ idx,
                   const float* __restrict__ a, const float* __restrict__ mean,
                   const float* __restrict__ c1, const float* __restrict__ c2, int N, int k, int Co,
                   long long M, float* __restrict__ dY, float* __restrict__ dU_acc,
                   float* __restrict__ dYhi, float* __restrict__ dYlo) {
  constexpr int SPW = 32 / LPP;
  const int lane = threadIdx.x & 31, sl = lane % LPP;
  const int G = Co / (4 * LPP), Co2 = 2 * Co;
  const long long S = (long long)gridDim.x * (BT / 32) * SPW;
  const long long s = ((long long)blockIdx.x * (BT / 32) + threadIdx.x / 32) * SPW + lane / LPP;
  const long long PS = S / G;
  const int c = (int)(s % G) * 4 * LPP + sl * 4;
  const long long ps = s / G;
  if (ps >= PS) return;
  const float4 av = *reinterpret_cast<const float4*>(a + c);
  const float4 mu = *reinterpret_cast<const float4*>(mean + c);
  const float4 k1 = *reinterpret_cast<const float4*>(c1 + c);
  const float4 k2 = *reinterpret_cast<const float4*>(c2 + c);
  const float kf = (float)k;
  for (long long m = ps; m < M; m += PS) {
    const long long base = (m / N) * N;
    const float4 gv = *reinterpret_cast<const float4*>(g + m * Co + c);
    const float4 es = *reinterpret_cast<const float4*>(esum + m * Co + c);
    const uchar4 aj = *reinterpret_cast<const uchar4*>(arg + m * Co + c);
    const float4 ag = make_float4(av.x * gv.x, av.y * gv.y, av.z * gv.z, av.w * gv.w);
    float4 dv;
    dv.x = ag.x - kf * k1.x - k2.x * (es.x - kf * mu.x);
    dv.y = ag.y - kf * k1.y - k2.y * (es.y - kf * mu.y);
    dv.z = ag.z - kf * k1.z - k2.z * (es.z - kf * mu.z);
    dv.w = ag.w - kf * k1.w - k2.w * (es.w - kf * mu.w);
    const int32_t* irow = idx + m * k;
    if (dU_acc) {
      // fused mode: dV is final here and goes out as tf32 hi/lo; the sparse part of dU accumulates
      // in dU_acc [M,Co] (zero-filled by the caller), finished by bwd_dense_kernel
      float4 h, l;
      h.x = ecb200::tf32_rna(dv.x); l.x = ecb200::tf32_rna(dv.x - h.x);
      h.y = ecb200::tf32_rna(dv.y); l.y = ecb200::tf32_rna(dv.y - h.y);
      h.z = ecb200::tf32_rna(dv.z); l.z = ecb200::tf32_rna(dv.z - h.z);
      h.w = ecb200::tf32_rna(dv.w); l.w = ecb200::tf32_rna(dv.w - h.w);
      *reinterpret_cast<float4*>(dYhi + m * Co2 + Co + c) = h;
      *reinterpret_cast<float4*>(dYlo + m * Co2 + Co + c) = l;
      atomicAdd(dU_acc + (base + irow[aj.x]) * Co + c + 0, ag.x);
      atomicAdd(dU_acc + (base + irow[aj.y]) * Co + c + 1, ag.y);
      atomicAdd(dU_acc + (base + irow[aj.z]) * Co + c + 2, ag.z);
      atomicAdd(dU_acc + (base + irow[aj.w]) * Co + c + 3, ag.w);
    } else {
      *reinterpret_cast<float4*>(dY + m * Co2 + Co + c) = dv;
      atomicAdd(dY + (base + irow[aj.x]) * Co2 + c + 0, ag.x);
      atomicAdd(dY + (base + irow[aj.y]) * Co2 + c + 1, ag.y);
      atomicAdd(dY + (base + irow[aj.z]) * Co2 + c + 2, ag.z);
      atomicAdd(dY + (base + irow[aj.w]) * Co2 + c + 3, ag.w);
    }
  }
}

template <int LPP>
unsigned group_grid(long long M, int Co) {
  const int G = Co / (4 * LPP);
  const long long per_cta = (long long)(BT / 32) * (32 / LPP);
  long long want = ecb200::ceil_div64(M * G, per_cta);
  const long long cap = 8LL * ecb200::kNumSMs;
  long long ctas = want < cap ? want : cap;
  if (ctas * per_cta < G) ctas = ecb200::ceil_div64(G, per_cta);
  return (unsigned)ctas;
}

#define ECB_DISPATCH_LPP(Co, CALL)            \
  do {                                        \
    if ((Co) % 128 == 0) { CALL(32); }        \
    else if ((Co) % 64 == 0) { CALL(16); }    \
    else if ((Co) % 32 == 0) { CALL(8); }     \
    else if ((Co) % 16 == 0) { CALL(4); }     \
    else if ((Co) % 8 == 0) { CALL(2); }      \
    else { CALL(1); }                         \
  } while (0)

}  // namespace

extern "C" int ecb200_bwd_prep(const float* gout, const float* gout_pm, long long ld_pm, const float* sel,
                               const float* a, const float* b, const float* mean, const float* invstd,
                               float slope, int B, int N, int Co, float* g, double* bstats, void* stream) {
  ECB_REQUIRE((gout || gout_pm) && sel && a && b && mean && invstd && g && bstats,
              "ecb200_bwd_prep: null pointer");
  ECB_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && Co >= 1, "ecb200_bwd_prep: bad shape");
  ECB_REQUIRE(!gout_pm || ld_pm >= Co, "ecb200_bwd_prep: ld_pm=%lld smaller than Co=%d", ld_pm, Co);
  dim3 grid(ecb200::ceil_div(N, 32), ecb200::ceil_div(Co, 32), B);
  bwd_prep_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(gout, gout_pm, ld_pm, sel, a, b, mean, invstd, slope,
                                                                 N, Co, g, bstats);
  ECB_LAUNCH_CHECK("bwd_prep_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_bwd_finalize(const double* bstats_local, const double* bstats_global,
                                   const double* count_dev, const float* a, const float* invstd, int training,
                                   int Co, float* dgamma, float* dbeta, float* c1, float* c2,
                                   void* stream) {
  ECB_REQUIRE(bstats_local && bstats_global && a && invstd && dgamma && dbeta && c1 && c2,
              "ecb200_bwd_finalize: null pointer");
  ECB_REQUIRE(Co >= 1 && (!training || count_dev), "ecb200_bwd_finalize: bad shape / missing count");
  bwd_finalize_kernel<<<ecb200::ceil_div(Co, 256), 256, 0, (cudaStream_t)stream>>>(
      bstats_local, bstats_global, count_dev, a, invstd, training, Co, dgamma, dbeta, c1, c2);
  ECB_LAUNCH_CHECK("bwd_finalize_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_reverse_graph(const int32_t* idx, int B, int N, int k, int32_t* rowptr,
                                    int32_t* src, int32_t* cursor, void* stream) {
  ECB_REQUIRE(idx && rowptr && src && cursor, "ecb200_reverse_graph: null pointer");
  ECB_REQUIRE(B >= 1 && N >= 1 && k >= 1, "ecb200_reverse_graph: bad shape");
  const long long M = (long long)B * N, E = M * k;
  ECB_REQUIRE(E < (1LL << 31), "ecb200_reverse_graph: B*N*k = %lld does not fit int32", E);
  cudaStream_t st = (cudaStream_t)stream;
  {
    // destination ranges per cloud: about one CTA per SM, each range's counters in shared memory
    int parts = ecb200::ceil_div(ecb200::kNumSMs, B);
    if (parts > 8) parts = 8;   // every CTA re-reads the cloud's whole neighbour list
    int span = ecb200::ceil_div(N, parts);
    if ((size_t)span * 2 * sizeof(int) > 160 * 1024) span = 160 * 1024 / (2 * sizeof(int));
    parts = ecb200::ceil_div(N, span);
    if (B <= 65535 && parts <= 64 && (long long)N * k < 40960) {
      const size_t smem = (size_t)span * 2 * sizeof(int);
      static thread_local bool seen_rows[ecb200::kMaxDevices] = {};
      if (ecb200::first_use_on_device(seen_rows))
        ECB_CUDA(cudaFuncSetAttribute(rev_cloud_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      rev_cloud_rows_kernel<<<dim3(parts, B), 1024, smem, st>>>(idx, N, k, M, span, rowptr, src);
      ECB_LAUNCH_CHECK("rev_cloud_rows_kernel");
      return ECB200_OK;
    }
    if (B <= 65535 && parts <= 64 && (long long)N * k < (1LL << 26) && k <= 64) {
      // the CTA's slice of src is staged in shared memory when it fits: 1.5x its expected size
      // (in-degrees are uneven), within what is left of 200 KB
      long long seg = (long long)span * k * 3 / 2;
      const long long room = (200 * 1024 - (long long)span * 2 * (long long)sizeof(int)) / (long long)sizeof(int);
      if (seg > room) seg = room > 0 ? room : 0;
      const size_t smem = ((size_t)span * 2 + (size_t)seg) * sizeof(int);
      static thread_local bool seen[ecb200::kMaxDevices] = {};
      if (ecb200::first_use_on_device(seen))
        ECB_CUDA(cudaFuncSetAttribute(rev_cloud_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      rev_cloud_kernel<<<dim3(parts, B), 1024, smem, st>>>(idx, N, k, M, span, (int)seg, rowptr, src);
      ECB_LAUNCH_CHECK("rev_cloud_kernel");
      return ECB200_OK;
    }
  }
  ECB_CUDA(cudaMemsetAsync(cursor, 0, sizeof(int32_t) * (size_t)M, st));
  const unsigned eb = (unsigned)ecb200::ceil_div64(E, 256);
  rev_count_kernel<<<eb, 256, 0, st>>>(idx, N, k, E, cursor);
  ECB_LAUNCH_CHECK("rev_count_kernel");
  rev_scan_kernel<<<B, 1024, 0, st>>>(cursor, N, k, M, rowptr);
  ECB_LAUNCH_CHECK("rev_scan_kernel");
  rev_fill_kernel<<<eb, 256, 0, st>>>(idx, rowptr, N, k, E, cursor, src);
  ECB_LAUNCH_CHECK("rev_fill_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_bwd_dense(const float* Y, const int32_t* rowptr, const int32_t* src,
                                const float* mean, const float* c1, const float* c2, int training,
                                int B, int N, int Co, float* dY, const float* dU_in, float* dYhi,
                                float* dYlo, void* stream) {
  ECB_REQUIRE(Y && mean && c1 && c2, "ecb200_bwd_dense: null pointer");
  ECB_REQUIRE(dU_in ? (dYhi && dYlo) : dY != nullptr,
              "ecb200_bwd_dense: needs dY (plain mode) or dU_in + dYhi + dYlo (fused mode)");
  ECB_REQUIRE(!training || (rowptr && src), "ecb200_bwd_dense: training needs the reverse graph");
  ECB_REQUIRE(B >= 1 && N >= 1 && Co >= 4 && Co % 4 == 0, "ecb200_bwd_dense: bad shape");
  const long long M = (long long)B * N;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(L)                                                                                  \
  bwd_dense_kernel<L><<<group_grid<L>(M, Co), BT, 0, st>>>(Y, rowptr, src, mean, c1, c2, training, \
                                                         Co, M, dY, dU_in, dYhi, dYlo)
  ECB_DISPATCH_LPP(Co, CALL);
#undef CALL
  ECB_LAUNCH_CHECK("bwd_dense_kernel");
  return ECB200_OK;
}

extern "C" int ecb200_bwd_scatter(const float* g, const float* esum, const uint8_t* arg,
                                  const int32_t* idx, const float* a, const float* mean,
                                  const float* c1, const float* c2, int B, int N, int k, int Co,
                                  float* dY, float* dU_acc, float* dYhi, float* dYlo, void* stream) {
  ECB_REQUIRE(g && esum && arg && idx && a && mean && c1 && c2, "ecb200_bwd_scatter: null pointer");
  ECB_REQUIRE(dU_acc ? (dYhi && dYlo) : dY != nullptr,
              "ecb200_bwd_scatter: needs dY (plain mode) or dU_acc + dYhi + dYlo (fused mode)");
  ECB_REQUIRE(B >= 1 && N >= 1 && k >= 1 && Co >= 4 && Co % 4 == 0, "ecb200_bwd_scatter: bad shape");
  const long long M = (long long)B * N;
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(L)                                                                                   \
  bwd_scatter_kernel<L><<<group_grid<L>(M, Co), BT, 0, st>>>(g, esum, arg, idx, a, mean, c1, c2, N, \
                                                           k, Co, M, dY, dU_acc, dYhi, dYlo)
  ECB_DISPATCH_LPP(Co, CALL);
#undef CALL
  ECB_LAUNCH_CHECK("bwd_scatter_kernel");
  return ECB200_OK;
}
