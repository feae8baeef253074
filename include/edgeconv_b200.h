/*
 * edgeconv_b200.h -- C ABI of libedgeconv_b200.so: the DGCNN EdgeConv hot path
 * (kNN graph + neighbour gather + edge MLP + BatchNorm + LeakyReLU + max over k)
 * as hand-written sm_100a CUDA kernels.
 *
 * This is the drop-in boundary for /root/reference/models/dgcnn.py.  The
 * reference is pure Python calling torch ops; a maintainer binds these entry
 * points with ctypes behind the same Python signatures:
 *
 *   knn(x, k)                         models/dgcnn.py:6-12   -> ecb200_sqnorms + ecb200_knn
 *   get_graph_feature(x, k, ...)      models/dgcnn.py:15-44  -> ecb200_graph_feature(_bwd)
 *   conv{n}(...) ; x.max(dim=-1)      models/dgcnn.py:54-73, :84-98
 *                                      -> ecb200_pack_weight, ecb200_point_gemm,
 *                                         ecb200_edge_gather, ecb200_bn_finalize,
 *                                         ecb200_edge_apply  (forward)
 *                                         ecb200_bwd_* + ecb200_gemm_dx/dw (backward)
 *
 * Conventions
 *   - every pointer is a DEVICE pointer into memory owned by the caller
 *     (PyTorch's allocator); the library never allocates, frees or synchronises,
 *     it only enqueues kernels on `stream` (a cudaStream_t passed as void*).
 *   - all floating-point tensors are contiguous fp32 unless stated; "stats"
 *     accumulators are fp64.  Neighbour indices are int32, local to their cloud
 *     (0..N-1); the Python `knn()` widens to int64 as the reference returns.
 *   - x is the reference's channel-major layout [B, C, N].  Point-major arrays
 *     are [M, *] with M = B*N and row m = b*N + n.
 *   - return value: 0 = success, otherwise an ECB200_ERR_* code; the message is
 *     available from ecb200_last_error() (thread-local).  No C++ exception
 *     crosses the boundary.  There is no CPU fallback.
 *   - re-entrant: no global mutable state; safe to call from one host thread
 *     per GPU (nn.DataParallel, main_cls.py:62) with the device already current.
 */
#ifndef EDGECONV_B200_H_
#define EDGECONV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define ECB200_VERSION 100            /* major*100 + minor */
#define ECB200_MAX_K 64               /* neighbours per point supported by the selector */
#define ECB200_AMAX_SLOTS 32          /* floats of an `amax` buffer (ecb200_absmax, ecb200_edge_apply_amax) */

enum {
  ECB200_OK = 0,
  ECB200_ERR_ARG = 1,                 /* bad shape / null pointer / unsupported size */
  ECB200_ERR_CUDA = 2                 /* a CUDA runtime call or launch failed */
};

/* graph-feature output layouts, get_graph_feature() models/dgcnn.py:37-42 */
enum {
  ECB200_GF_CONCAT = 0,               /* [B,2C,N,k]: (x_j , x_i)            dgcnn.py:42 */
  ECB200_GF_KNN_ONLY = 1,             /* [B,N,k,C] : x_j                    dgcnn.py:37-38 */
  ECB200_GF_DISP_ONLY = 2,            /* [B,C,N,k] : x_j - x_i              dgcnn.py:39-40 */
  ECB200_GF_CONCAT_CENTERED = 3       /* [B,2C,N,k]: (x_j - x_i , x_i)      test.ipynb cell 7 */
};

int ecb200_version(void);
const char* ecb200_last_error(void);

/* ---- kNN graph: replaces knn(), models/dgcnn.py:6-12 ------------------------------ */

/* xx[m] = sum_c x[b,c,n]^2  (dgcnn.py:8).  x [B,C,N] -> xx [B*N]. */
int ecb200_sqnorms(const float* x, int B, int C, int N, float* xx, void* stream);

/* idx[b,i,0..k-1] = the k points j of cloud b with the largest
 * -|x_i|^2 + 2 x_i.x_j - |x_j|^2 (dgcnn.py:9-11); ties broken towards the smaller j.
 * sorted != 0: nearest first, as Tensor.topk returns them; sorted == 0: the same set in
 * unspecified order (enough for EdgeConv, whose reductions over k are order-free).
 * FP32 FMA distance tiles feeding an on-chip top-k selector: the [B,N,N] matrix is never
 * written.  Requires 1 <= k <= min(N, 64). */
int ecb200_knn(const float* x, const float* xx, int B, int C, int N, int k, int sorted,
               int32_t* idx, void* stream);

/* Tensor-core variant for the feature-space layers (C a multiple of 32 in [32,128],
 * k <= 40): tcgen05 kind::tf32 tiles with 3xTF32 error compensation, FP32 accumulators in
 * TMEM, operand tiles moved by TMA, the same on-chip selector.
 *   ecb200_split_tf32: x[B,C,N] -> point-major hi = tf32(x), lo = tf32(x - hi) [B*N, C] and
 *                      xx[B*N] (same values as ecb200_sqnorms)
 *   ecb200_knn_tc:     idx as ecb200_knn; `workspace` (ecb200_knn_tc_workspace_bytes) holds the
 *                      per-row survivor lists of the selector's second pass.
 *   ecb200_debug_tc_scores: diagnostic -- the raw scores x_i.x_j - 0.5|x_j|^2 [B,N,N] produced
 *                      by the same MMA pipeline (only sensible for small N). */
int ecb200_split_tf32(const float* x, int B, int C, int N, float* hi, float* lo,
                      float* xx, void* stream);
size_t ecb200_knn_tc_workspace_bytes(int B, int N, int k);
int ecb200_knn_tc(const float* hi, const float* lo, const float* xx, int B, int C, int N,
                  int k, int sorted, int32_t* idx, void* workspace, size_t workspace_bytes,
                  void* stream);
int ecb200_debug_tc_scores(const float* hi, const float* lo, const float* xx, int B, int C, int N,
                           float* scores, void* stream);
/* The same kNN with packed FP16 operands (tcgen05 kind::f16): fp16 and tf32 carry the same 11-bit
 * significand, so the compensated product hi.hi + hi.lo + lo.hi is as accurate as 3xTF32 once the
 * tensor has been moved into fp16's range by a power of two -- at twice the tensor-pipe rate.  The
 * column term -0.5|x_j|^2 is part of the contraction (three more K slots: 2^15 on the query side, the
 * three fp16 pieces of -|s x_j|^2 / 2^16 on the candidate side), so the accumulator IS the score.
 *   ecb200_absmax:       max_s amax[s] = max |x| over n values; amax has ECB200_AMAX_SLOTS floats
 *                        (several slots, so that thousands of blocks do not hammer one address; the
 *                        callee zero-fills them first, consumers take the maximum over the slots)
 *   ecb200_split_f16:    s = 2^(12 - exponent(amax)); point-major halves hh = fp16(s x), hl = fp16(s x - hh)
 *                        [B*N, C] (C even, <= 128); nb [B*N, 64] halves = the norm block's candidate rows
 *                        (three pieces + zeros in the first 32 bytes of a 128-byte row); xxs[B*N] = s^2|x|^2;
 *                        cmax[B, ceil(N/32)] = max of xxs over each 32 points (the kNN kernels reduce
 *                        them to the cloud's maximum: no atomics, no zero-fill); optionally
 *                        (hi/lo/xx non-NULL) the outputs of ecb200_split_tf32 from the same pass over x
 *   ecb200_knn_tc_f16:   idx as ecb200_knn (nearest first); C = 64 or 128, k <= 40;
 *                        `timeline` NULL, or the diagnostic buffer of ecb200_debug_tc_timeline
 *   ecb200_pack_xyz_f16 / ecb200_knn_tc_xyz: the xyz layer (C <= 4) on the same pipeline -- the three
 *                        terms of a point's compensated product AND the norm term sit side by side in
 *                        ONE 16-deep K step (query rows [h|h|l|2^15 x3], candidate rows [h|l|h|pieces],
 *                        128-byte rows), one MMA per tile and sweep; the scale is per cloud, found by
 *                        the pack kernel, which also writes cmax[B, ceil(N/32)]
 *   ecb200_debug_tc_scores_f16: diagnostic -- s^2 x_i.x_j [B,N,N] from the hi/lo halves */
int ecb200_absmax(const float* x, long long n, float* amax, void* stream);
int ecb200_split_f16(const float* x, int B, int C, int N, const float* amax, void* hh, void* hl,
                     void* nb, float* xxs, float* cmax, float* hi, float* lo, float* xx, void* stream);
int ecb200_knn_tc_f16(const void* hh, const void* hl, const void* nb, const float* xxs,
                      const float* cmax, int B, int C, int N, int k, int32_t* idx, long long* timeline,
                      void* stream);
int ecb200_pack_xyz_f16(const float* x, int B, int C, int N, void* arow, void* brow, float* xxs,
                        float* cmax, void* stream);
int ecb200_knn_tc_xyz(const void* arow, const void* brow, const float* xxs, const float* cmax, int B,
                      int N, int k, int32_t* idx, long long* timeline, void* stream);
int ecb200_debug_tc_scores_f16(const void* hh, const void* hl, int B, int C, int N, float* scores,
                               void* stream);
/* hi = tf32(src), lo = tf32(src - hi), element-wise over n values (operand prep for the
 * tensor-core GEMM: Wcat is already K-major) */
int ecb200_split_rows_tf32(const float* src, long long n, float* hi, float* lo, void* stream);
/* diagnostic: ecb200_knn_tc with CTA (0,0) stamping clock64() into timeline[0 .. 6*256) and every
 * CTA its {globaltimer at start, at end, SM id} into timeline[6*256 + 3*cta ..] (int64;
 * 6*256 + 3*B*ceil(N/128) entries) */
int ecb200_debug_tc_timeline(const float* hi, const float* lo, const float* xx, int B, int C, int N,
                             int k, int32_t* idx, void* workspace, long long* timeline, void* stream);

/* ---- materialised graph feature: replaces get_graph_feature(), dgcnn.py:15-44 ------ */
int ecb200_graph_feature(const float* x, const int32_t* idx, int B, int C, int N, int k,
                         int mode, float* out, void* stream);
/* dx[B,C,N] (zero-filled by the callee) = gradient of the above w.r.t. x */
int ecb200_graph_feature_bwd(const float* gout, const int32_t* idx, int B, int C, int N, int k,
                             int mode, float* dx, void* stream);

/* ---- EdgeConv forward: replaces conv{n} + max over k, dgcnn.py:54-73, :84-98 ------- */

/* Wcat[2Co,C]: rows 0..Co-1 = W[:, :C], rows Co..2Co-1 = W[:, C:]  (minus W[:, :C] when
 * subtract_center, i.e. the canonical (x_j - x_i, x_i) edge feature).  W is the
 * Conv2d weight [Co, 2C] (dgcnn.py:55). */
int ecb200_pack_weight(const float* W, int Co, int C, int subtract_center, float* Wcat,
                       void* stream);

/* ecb200_pack_weight plus, in the same pass, the tf32 hi/lo halves of Wcat ([2Co,C]; operands of
 * ecb200_point_gemm_tc) and of Wcat^T ([C,2Co]; operands of ecb200_gemm_dx_tc).  Either pair may
 * be NULL. */
int ecb200_prepare_weights(const float* W, int Co, int C, int subtract_center, float* Wcat,
                           float* hi, float* lo, float* hiT, float* loT, void* stream);

/* Y[M,2Co] = [U | V],  U = x^T W1^T,  V = x^T W2'^T : the one dense per-point
 * GEMM that replaces the k-fold 1x1 convolution over [B,2C,N,k]. */
int ecb200_point_gemm(const float* x, const float* Wcat, int B, int C, int N, int Co2,
                      float* Y, void* stream);
/* The same GEMM on the tensor cores (tcgen05 kind::tf32, 3xTF32, TMA-fed, TMEM accumulators;
 * the pipeline of ecb200_knn_tc with a dense-store epilogue) from the point-major hi/lo
 * operands the kNN of the same layer already needs: xhi/xlo [M,C], whi/wlo [2Co,C].
 * C a multiple of 32 in [32,128]. */
int ecb200_point_gemm_tc(const float* xhi, const float* xlo, const float* whi, const float* wlo,
                         long long M, int C, int Co2, float* Y, void* stream);

/* For every point i and channel o, over its k neighbours j: e = U[idx[i,j],o] + V[i,o];
 *   sel[i,o]  = max_j e if gamma[o] >= 0 else min_j e       (what survives BN+LeakyReLU+max)
 *   arg[i,o]  = the slot j that attained it                 (for the scatter backward)
 *   esum[i,o] = sum_j e                                      (optional, may be NULL; backward)
 *   stats[0..Co-1] += sum_ij e, stats[Co..2Co-1] += sum_ij e^2, stats[2Co] += B*N*k
 *             (fp64 [2Co+1], optional, may be NULL; caller zero-fills.  The edge count
 *             rides along so that ONE all-reduce of this buffer makes all three global.)
 * Co % 4 == 0, Co <= 2048. */
int ecb200_edge_gather(const float* Y, const int32_t* idx, const float* gamma, int B, int N,
                       int k, int Co, float* sel, uint8_t* arg, float* esum, double* stats,
                       void* stream);

/* BatchNorm2d statistics -> per-channel affine (nn.BatchNorm2d semantics, dgcnn.py:56):
 * training: count = stats[2Co], mean = S1/count, var = S2/count - mean^2 (biased); `stats` is
 *           GLOBAL (already all-reduced under SyncBatchNorm).
 * eval:     mean/var = running_mean/running_var (read only).
 * out: mean, invstd, a = gamma*invstd, b = beta - a*mean  (all [Co] fp32). */
int ecb200_bn_finalize(const double* stats, const float* gamma, const float* beta,
                       const float* running_mean, const float* running_var, int training,
                       float eps, int Co, float* mean, float* invstd, float* a, float* b,
                       void* stream);

/* The in-place side effect of a training-mode BatchNorm2d forward:
 * running_mean <- (1-f)*running_mean + f*mean, running_var <- (1-f)*running_var + f*var_unbiased,
 * num_batches_tracked += 1; f = momentum, or 1/num_batches_tracked (after the increment) when
 * momentum < 0 (nn.BatchNorm2d(momentum=None)).  Any of the three pointers may be NULL. */
int ecb200_bn_update_running(const double* stats, int Co, float momentum, float* running_mean,
                             float* running_var, int64_t* num_batches_tracked, void* stream);

/* out[b,o,n] = leaky_relu(a[o]*sel[m,o] + b[o], slope)   ([M,Co] -> [B,Co,N]); the same values
 * optionally also point-major, out_pm[m*ld_pm + o] (rows of channels: a column slice of the
 * 512-channel concat in front of conv5, dgcnn.py:100).  Either of out / out_pm may be NULL. */
int ecb200_edge_apply(const float* sel, const float* a, const float* b, float slope, int B,
                      int N, int Co, float* out, float* out_pm, long long ld_pm, void* stream);
/* The same, and max_s amax[s] = max |out| (ECB200_AMAX_SLOTS floats, zero-filled by the callee first): the next layer's operand
 * scale for ecb200_split_f16, so the activations need no separate reduction pass. */
int ecb200_edge_apply_amax(const float* sel, const float* a, const float* b, float slope, int B,
                           int N, int Co, float* out, float* out_pm, long long ld_pm, float* amax,
                           void* stream);

/* ---- EdgeConv backward ------------------------------------------------------------ */

/* g[m,o] = (gout[b,o,n] + gout_pm[m*ld_pm+o]) * (a*sel+b > 0 ? 1 : slope): the gradients of the
 * channel-major output and of its point-major copy (either may be NULL);
 * bstats[0..Co-1] += sum g (d beta), bstats[Co..2Co-1] += sum g*(sel-mean)*invstd (d gamma);
 * fp64, caller zero-fills. */
int ecb200_bwd_prep(const float* gout, const float* gout_pm, long long ld_pm, const float* sel,
                    const float* a, const float* b, const float* mean, const float* invstd,
                    float slope, int B, int N, int Co, float* g, double* bstats, void* stream);

/* dgamma/dbeta (fp32) from the LOCAL sums; c1 = a*dbeta_g/count, c2 = a*dgamma_g*invstd/count
 * from the GLOBAL sums (equal to local on one GPU); count = *count_dev, the global edge count
 * the forward left in stats[2Co].  training == 0 -> c1 = c2 = 0 (count_dev may be NULL). */
int ecb200_bwd_finalize(const double* bstats_local, const double* bstats_global,
                        const double* count_dev,
                        const float* a, const float* invstd, int training, int Co,
                        float* dgamma, float* dbeta, float* c1, float* c2, void* stream);

/* Reverse (destination-major) kNN graph for the dense BatchNorm terms:
 * rowptr[M+1] (global offsets), src[M*k] (global source point ids).
 * `cursor` is an int32 [M] scratch array. */
int ecb200_reverse_graph(const int32_t* idx, int B, int N, int k, int32_t* rowptr,
                         int32_t* src, int32_t* cursor, void* stream);

/* dU = the dense BatchNorm-backward terms through the reverse graph (training; zeros in eval).
 * Plain mode (dU_in == NULL): written to dY[:, :Co] (fp32 [M,2Co]); run BEFORE ecb200_bwd_scatter,
 * which adds the sparse part.  Fused mode: dU_in [M,Co] already holds the sparse part
 * (ecb200_bwd_scatter ran first); the sum is written as tf32 hi/lo halves into
 * dYhi/dYlo[:, :Co] -- the operands of ecb200_gemm_dx_tc / _dw_tc, no fp32 dY, no split pass. */
int ecb200_bwd_dense(const float* Y, const int32_t* rowptr, const int32_t* src,
                     const float* mean, const float* c1, const float* c2, int training, int B,
                     int N, int Co, float* dY, const float* dU_in, float* dYhi, float* dYlo,
                     void* stream);

/* dV and the sparse scatter of a*g into dU rows through the arg slots.
 * Plain mode (dU_acc == NULL): dV -> dY[:, Co:], atomics -> dY[:, :Co].
 * Fused mode: dV -> tf32 hi/lo in dYhi/dYlo[:, Co:], atomics -> dU_acc [M,Co] (caller zero-fills). */
int ecb200_bwd_scatter(const float* g, const float* esum, const uint8_t* arg,
                       const int32_t* idx, const float* a, const float* mean, const float* c1,
                       const float* c2, int B, int N, int k, int Co, float* dY, float* dU_acc,
                       float* dYhi, float* dYlo, void* stream);

/* dx[B,C,N] = dY . Wcat ;  dWcat[2Co,C] = dY^T . x^T (callee zero-fills) */
int ecb200_gemm_dx(const float* dY, const float* Wcat, int B, int C, int N, int Co2, float* dx,
                   void* stream);
int ecb200_gemm_dw(const float* dY, const float* x, int B, int C, int N, int Co2, float* dWcat,
                   void* stream);
/* The two backward GEMMs on the tensor cores (gemm_tc.cu; tcgen05 kind::tf32, 3xTF32, TMA-fed,
 * FP32 accumulators in TMEM).  Operands are tf32 hi/lo halves:
 *   dYhi/dYlo [M,2Co]   = ecb200_split_rows_tf32(dY)
 *   WTh/WTl   [C,2Co]   = ecb200_transpose_split_tf32(Wcat)   (Wcat^T, K = 2Co contiguous)
 *   xhi/xlo   [M,C]     = the point-major operands of the forward (ecb200_split_tf32)
 * gemm_dx_tc streams K = 2Co through a TMA ring and writes dx in the reference's [B,C,N]
 * layout; gemm_dw_tc reduces over the M points with both operands consumed MN-major (no
 * transposed copies), slabs of points per CTA, partial tiles combined with fp32 atomics
 * (callee zero-fills dWcat).  C in {32,64,128}, 2Co a multiple of 32. */
int ecb200_transpose_split_tf32(const float* W, int R, int C, float* hiT, float* loT, void* stream);
int ecb200_gemm_dx_tc(const float* dYhi, const float* dYlo, const float* WTh, const float* WTl,
                      int B, int C, int N, int Co2, float* dx, void* stream);
int ecb200_gemm_dw_tc(const float* dYhi, const float* dYlo, const float* xhi, const float* xlo,
                      long long M, int C, int Co2, float* dWcat, void* stream);
/* dW[Co,2C] from dWcat (inverse of ecb200_pack_weight, including subtract_center) */
int ecb200_unpack_weight_grad(const float* dWcat, int Co, int C, int subtract_center, float* dW,
                              void* stream);

/* ---- "next" row f-2: conv5's BatchNorm2d + LeakyReLU (dgcnn.py:75-78, :102) fused with the global
 * max | average pooling over the points of the classification head ---------------------------
 * z [M, E] = conv5's raw output, point-major (cuDNN's output for a channels-last input); the
 * activated [B, E, N] tensor is never written. */

/* stats[0..E) += sum_m z, stats[E..2E) += sum_m z^2, stats[2E] += M  (fp64, caller zero-fills;
 * the same [sum | sum of squares | count] layout ecb200_bn_finalize / _update_running consume) */
int ecb200_colstats(const float* z, long long M, int E, double* stats, void* stream);
/* pooled[b, o] = max_n LeakyReLU(a[o] z + b[o]),  pooled[b, E+o] = mean_n LeakyReLU(a[o] z + b[o]),
 * arg[b, o] = the n that attains the max (smallest n among equals).  pooled [B, 2E], arg [B, E]. */
int ecb200_embed_pool(const float* z, const float* a, const float* b, float slope, int B, int N,
                      int E, float* pooled, int32_t* arg, void* stream);
/* backward, pass 1: with dact[m,o] = (gpool[b,E+o]/N + [n == arg[b,o]] gpool[b,o]) * LeakyReLU'(a z + b):
 * bstats[0..E) += sum dact (d beta), bstats[E..2E) += sum dact*(z-mean)*invstd (d gamma); fp64. */
int ecb200_embed_pool_bwd_stats(const float* z, const float* gpool, const int32_t* arg,
                                const float* a, const float* b, const float* mean,
                                const float* invstd, float slope, int B, int N, int E,
                                double* bstats, void* stream);
/* backward, pass 2: dz[m,o] = a*dact - c1 - c2*(z - mean)  (c1, c2 from ecb200_bwd_finalize) */
int ecb200_embed_pool_bwd_dz(const float* z, const float* gpool, const int32_t* arg, const float* a,
                             const float* b, const float* mean, const float* c1, const float* c2,
                             float slope, int B, int N, int E, float* dz, void* stream);

/* ---- multi-GPU: the SyncBatchNorm statistics exchange fused into one kernel over NVLink peer
 * memory (replaces the all_gather / all_reduce of torch's SyncBatchNorm, main_partseg_dist.py:189).
 * vals[0..n) (fp64, device) <- element-wise sum over the ranks, in rank order, in place.
 * peer_bufs: DEVICE array of `world` pointers, entry p = rank p's symmetric buffer of
 * ecb200_peer_buffer_bytes(world) bytes, zero-filled once, mapped into this process (e.g. through
 * torch.distributed._symmetric_memory); seq_counter: one device uint64 per buffer, zero-filled
 * once, private to the rank.  Every rank must issue the same sequence of exchanges.  No host
 * synchronisation; capturable in a CUDA graph. */
#define ECB200_PEER_MAX_VALUES 4160   /* >= 2*2048+1: [sum | sum sq | count] of up to 2048 channels */
size_t ecb200_peer_buffer_bytes(int world);
int ecb200_peer_allreduce(double* vals, int n, void* const* peer_bufs, int rank, int world,
                          unsigned long long* seq_counter, void* stream);

/* The exchange fused with the per-channel step that consumes it (one launch instead of two):
 *   ..._bn_finalize : stats[2Co+1] all-reduced in place, then exactly ecb200_bn_finalize(training)
 *   ..._bwd_finalize: bstats_global[2Co] (a copy of bstats_local) all-reduced in place, then exactly
 *                     ecb200_bwd_finalize(training) */
int ecb200_peer_allreduce_bn_finalize(double* stats, int Co, void* const* peer_bufs, int rank, int world,
                                      unsigned long long* seq_counter, const float* gamma,
                                      const float* beta, float eps, float* mean, float* invstd, float* a,
                                      float* b, void* stream);
int ecb200_peer_allreduce_bwd_finalize(const double* bstats_local, double* bstats_global, int Co,
                                       void* const* peer_bufs, int rank, int world,
                                       unsigned long long* seq_counter, const double* count_dev,
                                       const float* a, const float* invstd, float* dgamma, float* dbeta,
                                       float* c1, float* c2, void* stream);

/* conv5 of the backbone (models/dgcnn.py:74-78 applied at :102) as a per-point GEMM on the tensor cores
 * with its BatchNorm statistics in the epilogue: Z[M,E] = X[M,K] . W[E,K]^T (X = the channels-last
 * 512-channel concat of dgcnn.py:100), stats[2E+1] (fp64, zeroed by the caller; may be NULL) +=
 * [sum z | sum z^2 | M] in the layout ecb200_bn_finalize consumes.  xlo = wlo = NULL: plain TF32 on the
 * raw fp32 operands (the precision class of the library convolution under cudnn.allow_tf32, PyTorch's
 * default); with the tf32 halves (ecb200_split_rows_tf32) of both operands: 3xTF32, fp32-equivalent.
 * K a multiple of 32, E a multiple of 128. */
int ecb200_embed_gemm(const float* xhi, const float* xlo, const float* whi, const float* wlo,
                      long long M, int K, int E, float* Z, double* stats, void* stream);

/* ---- two-conv edge block, fused forward (row f-1) ------------------------------------------
 * Replaces, for inference, models/layers.py:45-52 of the reference (PositionEmbedding:
 * get_graph_feature -> conv1 -> conv2 -> max over k) and upstream's two-conv EdgeConv blocks.
 *   Y      [M, 2*C1]  the per-point GEMM of the FIRST conv, [U | V] (ecb200_point_gemm)
 *   a1, b1 [C1]       BatchNorm1 folded to an affine (ecb200_bn_finalize)
 *   w2hi/lo [C2, C1]  tf32 halves of the second conv's weight (ecb200_split_rows_tf32)
 *   sel2   [M, C2]    max_j z_ij where gamma2 >= 0, min_j z_ij otherwise, z = W2.LeakyReLU(a1 e1 + b1)
 *   stats2 [2*C2+1]   fp64, += [sum z | sum z^2 | edge count] (may be NULL); zero it first
 * C1 in {32, 64}, C2 in {32, 64, 128}, k <= 128.  Finish with ecb200_bn_finalize + ecb200_edge_apply. */
int ecb200_two_conv_fwd(const float* Y, const int32_t* idx, const float* a1, const float* b1,
                        float slope1, const float* w2hi, const float* w2lo, const float* gamma2,
                        int B, int N, int k, int C1, int C2, float* sel2, double* stats2, void* stream);

#if defined(__GNUC__)
/* Backward of BatchNorm over the rows of a small [M,F] activation (the cls head's bn6 / bn7 under
 * SyncBatchNorm), two launches around the statistics exchange:
 *   ecb200_rows_bn_bwd_stats: total[2F] fp64 = [sum_m g | sum_m g*xhat] (the vector that is all-reduced),
 *                             dbeta / dgamma [F] = the same sums in fp32 (local parameter gradients)
 *   ecb200_rows_bn_bwd_dx:    dx = a * (g - total[f]/n - xhat * total[F+f]/n), n = *count (device fp64) */
int ecb200_rows_bn_bwd_stats(const float* g, const float* x, const float* mean, const float* invstd,
                             int M, int F, double* total, float* dgamma, float* dbeta, void* stream);
int ecb200_rows_bn_bwd_dx(const float* g, const float* x, const float* mean, const float* invstd,
                          const float* a, const double* total, const double* count, int M, int F,
                          float* dx, void* stream);

/* ---- "next" row f-3: compute_hog_1x1 (models/model_partseg.py:15-92) on the device --------------
 * x [B,3,N] (channel-major, as the reference holds it), idx [B,N,k] = knn(x, k) (cloud-local, any
 * order) -> hist [B,N,18]: per point the L2-normalised 9-bin zenith and azimuth histograms (interleaved,
 * [9][2]) of its neighbours' principal directions, weighted by the square root of the leading
 * singular value.  Replaces the reference's device -> host copy + np.linalg.svd + host -> device copy.
 * Bug-compatible with the reference's un-offset gathers (:28-30, :51-54: every cloud reads the first
 * 3N floats of the batch and cloud 0's directions); the SIGN of each direction, which the reference
 * inherits from LAPACK and which changes the zenith bin, is fixed to v_z >= 0 (ties: v_y, then v_x).
 * dir_ws: N float4 of scratch (16-byte aligned). */
int ecb200_hog_1x1(const float* x, const int32_t* idx, int B, int N, int k, float* dir_ws,
                   float* hist, void* stream);

#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* EDGECONV_B200_H_ */
