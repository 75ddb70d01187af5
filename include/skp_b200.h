/*
 * skp_b200.h -- C ABI of libskp_b200.so: hand-written sm_100a CUDA for the StableKeypoints hot path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers + sizes + a cudaStream_t (as void*),
 * allocates nothing, never throws and never synchronises; it returns 0 on success or a negative
 * skp_status.  `skp_last_error()` gives a human-readable message for the calling thread.
 * All tensors are dense row-major fp32 unless a leading dimension (ld*) is given; indices are int64
 * (torch.long) so they can be shared with the host code without conversion.
 *
 * Each function names the reference code it replaces (paths relative to
 * /root/reference/unsupervised_keypoints).  The reference itself has no FFI: these are the calls
 * its Python would bind via ctypes.
 */
#ifndef SKP_B200_H
#define SKP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  SKP_OK = 0,
  SKP_ERR_INVALID = -1,   /* bad argument (null pointer, non-positive size, unsupported shape) */
  SKP_ERR_LAUNCH = -2,    /* cudaGetLastError() after a launch */
  SKP_ERR_UNSUPPORTED = -3,
  SKP_ERR_DRIVER = -4     /* TMA descriptor encode / driver entry point failure */
} skp_status;

#define SKP_MAX_LAYERS 8
#define SKP_GN_REPL 8

int skp_version(void);
const char* skp_last_error(void);
/* Number of kernels launched by this library in this process (bench.py "gpu_launches"). */
int64_t skp_launch_count(void);

/* ------------------------------------------------------------------ dense projections
 * C[M,N] = alpha * A[M,K] . B[N,K]^T (+ bias[N]) (+ residual[M,N]); fp32 in/out.
 * Replaces the to_q / to_k / to_v / to_out Linear calls of ptp_utils.py:483-491,541 and their
 * input-gradients (frozen weights: only dgrad exists, optimize_token.py:71-76).
 * _simt : fp32 FMA tiles (exact fp32).  _tc : tcgen05 tensor cores, split-bf16 (hi+lo, 3 MMAs, fp32
 * accumulate in TMEM), operands pre-split by skp_split_bf16 into K-major bf16 pairs and fed by TMA. */
int skp_gemm_nt_simt(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc,
                     int M, int N, int K, float alpha, const float* bias, const float* residual,
                     int64_t ldr, void* stream);
/* x[rows, cols] (ld) fp32 -> hi[rows, cols_pad], lo[rows, cols_pad] bf16 with x ~= hi + lo;
 * cols_pad >= cols is a multiple of 64 and the pad is zero-filled. */
int skp_split_bf16(const float* x, int64_t ld, int rows, int cols, int cols_pad, void* hi, void* lo,
                   void* stream);
/* Split-K for the small-M layers that cannot fill 148 SMs.  splits == 0 (the product's setting): the library plans the
 * K-splits; the splits of one output tile are the CTAs of one thread-block cluster (<= 8) that reduce their partial tiles
 * through distributed shared memory in rank order inside the same kernel -- no workspace (splitk_ws may be NULL), no second
 * launch, bit-reproducible.  splits == 1: no split.  splits > 1: the older path for A/B measurements, partial sums in
 * splitk_ws[splits*M*N] then one reduce+epilogue kernel; skp_gemm_nt_tc_plan returns the planner's split count. */
int skp_gemm_nt_tc_plan(int M, int N, int Kpad);
/* Tuning hook: force the N tile (64/96/128/160/256; 0 = planner) of skp_gemm_nt_tc / skp_conv3x3_tc (scripts/gemm_sweep.py). */
void skp_gemm_tc_force_bn(int bn);
/* Persistent form of the two kernels above (one CTA per SM walking several output tiles, two accumulators in tensor memory so
 * that the epilogue of a tile overlaps the main loop of the next): 0 (default) never, 1 un-split problems with more tiles
 * than SMs, 2 every un-split problem with several tiles per CTA (tests).  Env: SKP_GEMM_PERSIST.  Opt-in: faster per launch
 * and bit-identical in isolation, not yet stable inside the multi-stream step (profiles/r02_gemm_persist.md). */
void skp_gemm_tc_persist(int mode);
int skp_gemm_nt_tc(const void* A_hi, const void* A_lo, const void* B_hi, const void* B_lo, int Kpad,
                   float* C, int64_t ldc, int M, int N, float alpha, const float* bias,
                   const float* residual, int64_t ldr, int splits, float* splitk_ws, void* stream);
/* 3x3 im2col of a channels-last activation x[H*W, C] (row stride ldx) written directly as the split-bf16 K-major
 * A operand [Ho*Wo, Kpad] (column = tap*C + c; zero outside the image and in the K padding).  Together with
 * skp_gemm_nt_tc this is the frozen 3x3 convolution (and, with flipped weights, its input gradient) of the
 * UNet / VAE resnets the path runs through (diffusers ResnetBlock2D). */
int skp_im2col3x3_split(const float* x, int64_t ldx, int H, int W, int C, int Ho, int Wo, int stride, int pad,
                        int Kpad, void* hi, void* lo, void* stream);
/* Implicit-GEMM 3x3 convolution, stride 1, zero padding 1, on the same tcgen05 kernel: X_hi/X_lo are the split-bf16
 * channels-last activation [H, W, Cin] (Cin % 64 == 0, W % 8 == 0), B the filter as [Cout, 9*Cin] (column = tap*Cin + c).
 * Each k-block's A tile is a 3-D TMA box of the activation shifted by the tap; the image border is the TMA's
 * out-of-bounds zero fill, so no im2col buffer exists.  Output C[H*W, Cout] (+bias +residual), split-K as above. */
int skp_conv3x3_tc(const void* X_hi, const void* X_lo, int H, int W, int Cin, const void* B_hi, const void* B_lo,
                   float* C, int64_t ldc, int Cout, float alpha, const float* bias, const float* residual,
                   int64_t ldr, int splits, float* splitk_ws, void* stream);

/* ------------------------------------------------------------------ GroupNorm (+SiLU) of the frozen trunk
 * Channels-last activations x[rows = H*W, C] (row stride ldx), `groups` groups of C/groups channels (diffusers
 * ResnetBlock2D / Transformer2DModel norms).  sums[SKP_GN_REPL][2*groups] (fp64: sum, sum of
 * squares; CTA b accumulates into replica b % SKP_GN_REPL, consumers add the replicas) is written by skp_gn_stats and
 * consumed by the others.  The normalised tensor is never stored on its own:
 * skp_gn_apply emits fp32 y[rows, C] and/or the split-bf16 K-major operand [rows, Kpad] of a following 1x1
 * projection; skp_gn_im2col3x3_split emits the split-bf16 3x3 im2col operand of a following convolution.
 * skp_gn_bwd: dx = d(loss)/dx from g = d(loss)/dy (gamma, beta frozen); bsums[SKP_GN_REPL][2*groups] is fp64 workspace. */
int skp_gn_stats(const float* x, int64_t ldx, int rows, int C, int groups, double* sums, void* stream);
/* skp_gn_stats + skp_gn_apply in one call; small activations take ONE launch (one CTA per group: no atomics, no memset). */
int skp_gn_fwd(const float* x, int64_t ldx, int rows, int C, int groups, float eps, const float* gamma,
               const float* beta, int silu, float* y, int64_t ldy, void* hi, void* lo, int Kpad, double* sums,
               void* stream);
int skp_gn_apply(const float* x, int64_t ldx, int rows, int C, int groups, const double* sums, float eps,
                 const float* gamma, const float* beta, int silu, float* y, int64_t ldy, void* hi, void* lo,
                 int Kpad, void* stream);
int skp_gn_im2col3x3_split(const float* x, int64_t ldx, int H, int W, int C, int groups, const double* sums,
                           float eps, const float* gamma, const float* beta, int silu, int Ho, int Wo,
                           int stride, int pad, int Kpad, void* hi, void* lo, void* stream);
int skp_gn_bwd(const float* x, int64_t ldx, const float* g, int64_t ldg, int rows, int C, int groups,
               const double* sums, float eps, const float* gamma, const float* beta, int silu, double* bsums,
               float* dx, int64_t ldd, void* stream);

/* ------------------------------------------------------------------ LayerNorm / GEGLU of the transformer blocks
 * diffusers BasicTransformerBlock: norm1/2/3 feed attn1.qkv, attn2.to_q (ptp_utils.py:483) and
 * ff.net.0.proj; GEGLU = a * gelu(gate) (exact erf) feeds ff.net.2.  Each is fused with the split-bf16 operand write of
 * the projection that follows: hi/lo[rows, Kpad] bf16 (Kpad >= C, multiple of 64, zero padded).  C % 4 == 0.
 * stats[rows, 2] receives (mean, rstd) per row for skp_ln_bwd: dx = d/dx of LayerNorm given g = d/dy (gamma frozen). */
int skp_ln_split_fwd(const float* x, int64_t ldx, int rows, int C, const float* gamma, const float* beta, float eps,
                     void* hi, void* lo, int Kpad, float* stats, void* stream);
int skp_ln_bwd(const float* x, int64_t ldx, const float* g, int64_t ldg, int rows, int C, const float* gamma,
               const float* stats, float* dx, int64_t lddx, void* stream);
/* proj[rows, 2H] = (a | gate): hi/lo = split(a * gelu(gate));  bwd: dproj[rows, 2H] from g = d/d(a * gelu(gate)). */
int skp_geglu_split_fwd(const float* proj, int64_t ld, int rows, int H, void* hi, void* lo, int Kpad, void* stream);
int skp_geglu_bwd(const float* proj, int64_t ld, const float* g, int64_t ldg, int rows, int H, float* dproj, int64_t ldd,
                  void* stream);

/* hi/lo[rows, Kpad] = split-bf16( softmax(x[rows, cols], -1) ), cols % 4 == 0: the probability operand of the dense
 * (two-GEMM) attention used for the VAE mid-block AttentionBlock (one head of 512 channels; diffusers AutoencoderKL
 * encoder, reached from ptp_utils.py:299-302). */
int skp_softmax_split_fwd(const float* x, int64_t ldx, int rows, int cols, void* hi, void* lo, int Kpad, void* stream);

/* ------------------------------------------------------------------ cross-attention core
 * ptp_utils.py:493-506: sim = q k^T * scale; attn = softmax(sim, -1); out = attn v, per head.
 * q,o: [S, heads*d] (ld = heads*d); k,v: [N, heads*d] with leading dims ldk/ldv (slices of the batched
 * K|V projection).  logits[heads, S, N] receives the scaled scores (kept for backward and for the
 * capture kernels). */
int skp_cross_attn_fwd(const float* q, const float* k, int64_t ldk, const float* v, int64_t ldv,
                       float* o, float* logits, int S, int N, int heads, int d, float scale, void* stream);
/* Backward of the above.  d_logits_extra (nullable) is added to d(sim) -- it carries the gradient that
 * arrives through the captured maps.  dk/dv must be zero-initialised [N, heads*d] (atomically
 * accumulated); ds_ws is a workspace of heads*S*(N+2) floats (d(sim) + per-row softmax statistics). */
int skp_cross_attn_bwd(const float* d_o, const float* q, const float* k, int64_t ldk, const float* v,
                       int64_t ldv, const float* logits, const float* d_logits_extra, float* ds_ws,
                       float* dq, float* dk, float* dv, int S, int N, int heads, int d, float scale,
                       void* stream);

/* ------------------------------------------------------------------ self-attention core (attn1 of every transformer block)
 * The patched CrossAttention.forward runs the same lines for context=None (ptp_utils.py:480-506, is_cross == False):
 * out = softmax(q k^T * scale) v per head with k,v projected from the layer's own tokens.  Flash-style (no [S,S]
 * tensor in HBM), every contraction a split-bf16 (hi+lo, 3 MMAs, fp32 accumulate) tensor-core product so the
 * captured maps downstream keep their 1e-3 budget.  q,k,v,o: [S, heads*d] fp32 with leading dims (q,k,v may be column
 * slices of one [S, 3*heads*d] projection); d even, d <= 160.  planes is a workspace of 6*heads*S*DP bf16 with
 * DP = skp_self_attn_dp(d) (the split operands; kept by the caller for the backward); lse[heads, S] receives the
 * per-row log2-sum-exp. */
int skp_self_attn_dp(int d);
int skp_self_attn_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                      float* o, int64_t ldo, float* lse, void* planes, int S, int heads, int d, float scale,
                      void* stream);
/* Backward: dq,dk,dv [S, heads*d] (leading dims given; fully written).  do_planes: workspace of 2*heads*S*DP bf16;
 * dvec: workspace of heads*S floats (rowsum(dO * O)). */
int skp_self_attn_bwd(const float* d_o, int64_t lddo, const float* o, int64_t ldo, const float* lse,
                      const void* planes, void* do_planes, float* dvec, float* dq, int64_t lddq, float* dk,
                      int64_t lddk, float* dv, int64_t lddv, int S, int heads, int d, float scale, void* stream);

/* Long-sequence self-attention forward on tcgen05 (QK^T and PV as tcgen05.mma with TMEM accumulators, operands by TMA,
 * P handed from the softmax warps to the tensor core through 128B-swizzled shared memory; same split-bf16 numerics).
 * Needs S % 128 == 0 and even d <= 64; workspace = skp_self_attn_tc_workspace(S, heads, d) bytes (0 = shape not
 * eligible), 128-byte aligned.  lse as in skp_self_attn_fwd; the backward is skp_self_attn_bwd on planes made by
 * skp_self_attn_split (6*heads*S*DP bf16). */
int64_t skp_self_attn_tc_workspace(int S, int heads, int d);
int skp_self_attn_tc_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                         float* o, int64_t ldo, float* lse, void* workspace, int S, int heads, int d, float scale,
                         void* stream);
int skp_self_attn_split(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                        void* planes, int S, int heads, int d, float scale, void* stream);

/* Long-sequence self-attention BACKWARD on tcgen05 (skp_attn_tc_bwd.cu; the autograd pass through ptp_utils.py:493-506):
 * one launch whose CTAs own either 128 keys (dK, dV accumulated in TMEM over all query tiles) or 128 queries (dQ); scores
 * and dP recomputed per tile as tcgen05.mma, P / dS handed back to the tensor core through 128B-swizzled shared memory;
 * split-bf16 products, no atomics.  Needs S % 128 == 0 and even d <= 96; takes o and the base-2 lse of either forward
 * (skp_self_attn_tc_fwd or skp_self_attn_fwd) together with the fp32 q / k / v it re-splits; workspace = skp_self_attn_tc_bwd_workspace bytes,
 * 128-byte aligned (0 = shape not eligible). */
int64_t skp_self_attn_tc_bwd_workspace(int S, int heads, int d);
int skp_self_attn_tc_bwd(const float* d_o, int64_t lddo, const float* o, int64_t ldo, const float* lse, const float* q,
                         int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, void* workspace, float* dq,
                         int64_t lddq, float* dk, int64_t lddk, float* dv, int64_t lddv, int S, int heads, int d, float scale,
                         void* stream);

/* Cross-attention (attn2, ptp_utils.py:480-506) and short-sequence self-attention forward on tcgen05 / TMEM / TMA
 * (skp_xattn_tc.cu): any Sq / Skv, even head dims up to 160 (64-column K chunks), keys in tiles of 64 with the accumulator
 * resident in tensor memory.  logits != NULL (captured layers, ptp_utils.py:508-538): the scaled logits [heads, Sq, Skv] are
 * written by the same kernel.  lse is in base 2 (scores pre-multiplied by log2 e), like skp_cross_attn_tc_fwd.
 * skp_cross_attn_split makes the operand planes skp_cross_attn_tc_bwd needs from q / k / v. */
int64_t skp_xattn_tc_workspace(int Sq, int Skv, int heads, int d);
int skp_xattn_tc_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv, float* o,
                     int64_t ldo, float* lse, float* logits, void* workspace, int Sq, int Skv, int heads, int d,
                     float scale, void* stream);
int skp_cross_attn_split(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                         void* q_planes, void* kv_planes, int S, int N, int heads, int d, float scale, void* stream);
/* Cross-attention core on the same split-bf16 tensor-core kernels (S queries, N << S keys): the tensor-core
 * replacement of skp_cross_attn_fwd/bwd.  logits (nullable) receives the scaled scores [heads, S, N] of a captured
 * layer; lse[heads, S]; q_planes: 2*heads*S*DP bf16, kv_planes: 4*heads*N*DP bf16 (kept for the backward).
 * Backward: d_logits_extra (nullable, [heads, S, N]) is added to d(sim); dk/dv [N, heads*d] MUST be zero-initialised
 * (the query axis is split over CTAs and accumulated atomically); do_planes: 2*heads*S*DP bf16; dvec: heads*S floats. */
int skp_cross_attn_tc_fwd(const float* q, int64_t ldq, const float* k, int64_t ldk, const float* v, int64_t ldv,
                          float* o, int64_t ldo, float* lse, float* logits, void* q_planes, void* kv_planes, int S,
                          int N, int heads, int d, float scale, void* stream);
int skp_cross_attn_tc_bwd(const float* d_o, int64_t lddo, const float* o, int64_t ldo, const float* lse,
                          const void* q_planes, const void* kv_planes, void* do_planes, float* dvec,
                          const float* d_logits_extra, float* dq, int64_t lddq, float* dk, int64_t lddk, float* dv,
                          int64_t lddv, int S, int N, int heads, int d, float scale, void* stream);

/* Kernel selection switch for tests and A/B measurements.  row_fwd (attn-store forward): 2 (default) register formulation
 * (skp_capture_store.cu), 1 row formulation (skp_capture_row.cu), 0 the tile kernel they fall back to.  row_bwd (fused
 * capture+collect backward): 1 (default) row formulation, 0 tile kernel.  -1 leaves a setting unchanged. */
void skp_capture_select(int row_fwd, int row_bwd);
/* The tcgen05 formulation of the attn-store / fused capture+collect forward (skp_capture_tc.cu: the horizontal bicubic pass
 * is an M128 x N=tokens x K=s GEMM per output row, the softmax runs thread-per-pixel on the TMEM lanes, the row leaves
 * through the copy engine) serves N <= 128 tokens and s <= 32.  on = 0: never; 1 (default): where it is the faster kernel
 * (the fused capture+collect always, the attn-store for N > 96); 2: wherever the shape is eligible (tests, A/B runs). */
void skp_capture_tc(int on);
int skp_capture_tc_ok(const int* s, int n_layers, int N, int R, int store);
/* Workspace bytes (row maxima of the low-res logits, written by a pre-pass) for the two entry points below. */
int64_t skp_capture_tc_workspace(const int* s, int n_layers, int heads);
/* skp_capture_store_fwd / skp_capture_mean_fwd on the tcgen05 kernel; shapes it does not take (N > 128, s > 32, s % 4 != 0,
 * more than two distinct sides, a NULL workspace) run the SIMT kernels instead, so the result contract is the same. */
int skp_capture_store_tc_fwd(const float* logits, float* probs, int heads, int s, int N, int R, float* workspace,
                             void* stream);
int skp_capture_mean_tc_fwd(const float* const* logits, const int* s, int n_layers, float* maps, int heads, int N,
                            int R, float* workspace, void* stream);
/* Debug: device buffer of 3*64*8 int64 that CTA 0 of the tcgen05 kernel fills with clock64() stamps per pipeline role
 * (scripts/capture_tc_trace.py); NULL (default) switches the stamps off. */
void skp_capture_tc_trace(void* buf);
/* ------------------------------------------------------------------ attention-store ("capture")
 * ptp_utils.py:508-538: bicubic (align_corners=False, A=-0.75, clamped taps) upsample of the layer
 * input to R x R, to_q, q' k^T * scale, softmax over the TOKEN axis, stored as [heads, R*R, N].
 * Because to_q is bias-free and bicubic resampling is linear, q'k^T == bicubic(q k^T): the kernels take
 * the layer's own low-res scaled logits [heads, s, s, N] (SURVEY.md 8a note).
 * _store : materialises probs[heads, R*R, N] (what AttentionStore.step_store["attn"] holds).
 * _store_bwd : d_logits[heads,s,s,N] += bicubic^T( softmax-backward(probs, d_probs) ).  */
int skp_capture_store_fwd(const float* logits, float* probs, int heads, int s, int N, int R, void* stream);
int skp_capture_store_bwd(const float* logits, const float* d_probs, float* d_logits, int heads, int s,
                          int N, int R, void* stream);
/* Fused capture + collect_maps (optimize.py:50-75 with upsample_res=-1, indices=None): over the given
 * layers (each with its own s), maps[N, R, R] = mean over (layer, head) of the captured probabilities,
 * never materialising them.  _bwd recomputes and accumulates into d_logits[l] (zero-initialised). */
int skp_capture_mean_fwd(const float* const* logits, const int* s, int n_layers, float* maps, int heads,
                         int N, int R, void* stream);
/* workspace: skp_capture_mean_bwd_workspace bytes (per-row partial gradients [heads, R, s, N] of the largest layer; the
 * transposed bicubic is then two gathers -- no atomics, bit-reproducible); NULL runs the tile kernel (shared atomics). */
int64_t skp_capture_mean_bwd_workspace(const int* s, int n_layers, int heads, int N, int R);
int skp_capture_mean_bwd(const float* const* logits, const int* s, int n_layers, const float* d_maps,
                         float* const* d_logits, int heads, int N, int R, float* workspace, void* stream);

/* ------------------------------------------------------------------ collect_maps (optimize.py:27-79)
 * stored[l] : [BH, R*R, N].  out[T, R2, R2] = bilinear_{R->R2}( mean_{l,bh} stored[l][bh, :, idx[t]] ),
 * T = n_idx if idx != NULL else N; R2 == R means no resize.  (Mean-then-resize equals the reference's
 * resize-then-mean: both are linear.)  tmp is a [T, R, R] workspace (may alias out when R2 == R). */
int skp_collect_maps_fwd(const float* const* stored, int n_layers, int BH, int R, int N,
                         const int64_t* idx, int n_idx, int R2, float* tmp, float* out, void* stream);
/* d_stored[l][bh, pix, idx[t]] = bilinear^T(d_out)[t, pix] / (n_layers*BH); other tokens get 0.
 * d_stored buffers are fully written (no pre-zeroing needed). */
int skp_collect_maps_bwd(const float* d_out, int n_layers, int BH, int R, int N, const int64_t* idx,
                         int n_idx, int R2, float* tmp, float* const* d_stored, void* stream);

/* ------------------------------------------------------------------ arg-max / selection
 * eval.py:39-60 find_max_pixel: first-occurrence arg-max of each [H*W] map; flat index out. */
int skp_argmax_rows(const float* maps, int T, int P, int64_t* flat_idx, void* stream);
/* eval.py:62-111 find_k_max_pixels with num>1: arg-max, then multiply by 0 inside radius 0.05*H
 * (integer pixel coords vs the +0.5 centre), repeated; flat_idx[num, T].  work is a [T,H,W] scratch. */
int skp_k_argmax(const float* maps, int T, int H, int W, int num, float* work, int64_t* flat_idx,
                 void* stream);
/* ptp_utils.py:97-108: KL( normalised(Gaussian@argmax + eps) || softmax_pixels(map + eps) ) per token.
 * peaks[num, T] are flat arg-max indices (num_subjects of them; the target is their mean). */
int skp_gaussian_kl_scores(const float* maps, int T, int H, int W, const int64_t* peaks, int num,
                           float sigma, float eps, float* kl, void* stream);
/* ptp_utils.py:165-187 entropy_sort (--top_k_strategy entropy): entropy of softmax-over-pixels of each [P] map. */
int skp_entropy_scores(const float* maps, int T, int P, float* ent, void* stream);
/* ptp_utils.py:110-112: ascending arg-sort of T scores (stable), first top_k indices. */
int skp_argsort_topk(const float* scores, int T, int top_k, int64_t* out_idx, void* stream);
/* ptp_utils.py:115-159 furthest_point_sampling on arg-max locations (flat indices of the maps it is
 * given, normalised by H): furthest pair then greedy max-min distance, strict '>' tie-breaking.
 * n_out receives the number of indices written (== top_k unless candidates run out). */
int skp_furthest_point_sampling(const int64_t* peaks_flat, int H, int W, const int64_t* candidates,
                                int n_cand, int top_k, int64_t* out_idx, int32_t* n_out, void* stream);

/* ------------------------------------------------------------------ losses
 * optimize.py:166-206 sharpening_loss on maps[sel[k]] : mean over (K,H,W) of (map - G)^2 with
 * G = mean over num peaks of exp(-|p - peak|^2 / (2 sigma^2)) on the +0.5 grid (optimize_token.py:203-241).
 * peaks[num, K] are flat arg-max indices of the SELECTED maps.  loss is a device scalar (overwritten). */
int skp_sharpen_loss_fwd(const float* maps, int H, int W, const int64_t* sel, int K, const int64_t* peaks,
                         int num, float sigma, float* loss, void* stream);
/* d_maps[sel[k], :] += g * 2 (map - G) / (K H W), g read from the device scalar d_loss times weight. */
int skp_sharpen_loss_bwd(const float* maps, int H, int W, const int64_t* sel, int K, const int64_t* peaks,
                         int num, float sigma, const float* d_loss, float weight, float* d_maps,
                         void* stream);
/* optimize.py:157-163 + invertable_transform.py:72-92: mean (maps[sel] - unwarp(maps_t[sel]))^2 where
 * unwarp = bilinear grid_sample(zeros padding, align_corners=False) on affine_grid(theta_inv[2x3]).
 * theta_inv is a 6-float DEVICE array. */
int skp_equivariance_loss_fwd(const float* maps, const float* maps_t, int H, int W, const int64_t* sel,
                              int K, const float* theta_inv, float* loss, void* stream);
int skp_equivariance_loss_bwd(const float* maps, const float* maps_t, int H, int W, const int64_t* sel,
                              int K, const float* theta_inv, const float* d_loss, float weight,
                              float* d_maps, float* d_maps_t, void* stream);
/* invertable_transform.py:65-68 / :87-90: out[b,c] = grid_sample(img[b,c], affine_grid(theta[b])) for
 * [B,C,H,W]; theta is a [B,2,3] DEVICE array. */
int skp_affine_warp(const float* img, int B, int C, int H, int W, const float* theta, float* out,
                    void* stream);
/* Transposed sampler: d_img (zero-initialised) += scatter of d_out through the same bilinear taps. */
int skp_affine_warp_bwd(const float* d_out, int B, int C, int H, int W, const float* theta, float* d_img,
                        void* stream);
/* eval.py:113-155 pixel_from_weighted_avg: zero (in place) beyond `distance` px of the arg-max,
 * normalise by sum+1e-6, expectation of (row, col) + 0.5 -> out[T,2].  peaks: flat arg-max per map. */
int skp_soft_argmax(float* heatmaps, int T, int H, int W, const int64_t* peaks, float distance,
                    float* out, void* stream);

/* Eval-time augmentation ensemble (eval.py:250-262): one pass does sum += unwarp(maps), num += unwarp(ones) for
 * maps[K,H,W] under the inverse affine theta_inv (6-float DEVICE array); skp_ensemble_finalize: out = sum/num with
 * 0/0 -> 0 (eval.py:333-336). */
int skp_unwarp_accumulate(const float* maps, int K, int H, int W, const float* theta_inv, float* sum_samples,
                          float* num_samples, void* stream);
int skp_ensemble_finalize(const float* sum_samples, const float* num_samples, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------ optimiser (optimize.py:320,424)
 * torch.optim.Adam defaults (no weight decay / amsgrad); grad is scaled by grad_scale first (the
 * 1/world_size of the data-parallel mean, optimize.py:405-406). */
int skp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, int step,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Same update with the step count kept on the device (incremented by the call), so that an optimizer step
 * captured in a CUDA graph stays correct on replay. */
int skp_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      int* step_dev, float lr, float beta1, float beta2, float eps, float grad_scale,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKP_B200_H */
