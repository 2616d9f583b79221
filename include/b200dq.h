/*
 * b200dq.h - C ABI of libb200dq.so: the sm_100a kernels behind the DQ-VAE stage-1 hot path.
 *
 * The reference (CrossmodalGroup/DynamicVectorQuantization) has no FFI: its operator interface is
 * "a Python class at a dotted path" (utils/utils.py:41-51).  The overlay classes in
 * dynamicvectorquantization_b200/overlay/ keep that interface and reach the GPU exclusively
 * through the entry points below (ctypes binding: dynamicvectorquantization_b200/_cabi.py; the
 * stub a reference maintainer would add is shown in INTEGRATION.md).
 *
 * Conventions
 *   - every function returns 0 on success, a cudaError_t (>0) or a negative argument/descriptor
 *     error otherwise; nothing throws, allocates device memory, or synchronises;
 *   - all pointers are device pointers owned by the caller (PyTorch); work is enqueued on `stream`;
 *   - activations are NHWC bf16 ("[N,H,W,C]", C contiguous); parameters and statistics fp32;
 *   - scratch buffers are passed in by the caller (sizes stated per function).
 */
#ifndef B200DQ_H_
#define B200DQ_H_

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

int b2dq_version(void);

/* ------------------------------------------------------------------ vector quantisation
 * Replaces modules/vector_quantization/quantize2_mask.py:
 *   VQEmbedding.compute_distances/find_nearest_embedding (:29-55), the one-hot scatter + matmul of
 *   _update_buffers (:77-84), embed (:123,130-132) and the masked loss of VectorQuantize2.forward
 *   (:172-179) in one launch; _update_buffers' EMA/restart (:90-105) and _update_embedding (:107-115)
 *   in b2dq_vq_ema_finalize; the straight-through + commitment gradient (:172-182) in b2dq_vq_bwd.
 */

/* weight_f32 [K(+1),C] -> cb_bf16 [K,C], cb_sqnorm [round_up(K,256)] (+inf padded). */
int b2dq_vq_prepare_codebook(const float* weight_f32, void* cb_bf16, float* cb_sqnorm, int K, int C,
                             cudaStream_t stream);

/* Nearest-code search over x_bf16 [N,C] (C in {64,128,192,256}); lowest index wins ties.
 *   x_f32      optional fp32 copy of x used for loss / EMA sums (else the bf16 values are used)
 *   row_mask   optional [N] weights of the commitment loss (codebook_mask, EncoderDual.py:147-149)
 *   codes      [N] int64;  xq_bf16 / xq_f32: optional gathered rows of weight_f32 (pre-update)
 *   loss_acc   optional [1], += sum_rows mask * sum_c (e - x)^2
 *   counts [K] / sums [K,C]: optional, += per-code row count / row sum (training mode)
 *   max_ctas   0 = one CTA per SM
 *   workspace  optional scratch of b2dq_vq_search_workspace_bytes(N, K) bytes (8 B aligned).  With it and
 *              K >= 2048, the row tiles left over after the full waves of the grid - all of them for the small
 *              N of the residual quantizer depth loop (quantize_rqvae.py:237-271) or stage-2 sampling - are cut
 *              stream-K style into runs of codebook tiles shared between CTAs (64-bit atomicMin of ordered
 *              distance | index, the last CTA of a row tile gathers); results are identical. */
int b2dq_vq_search_plan(int N, int K, int num_ctas, int allow_split, int* out5 /* host: grid, full rounds,
                        tail row tiles, codebook tiles per tail run, split flag */);
int b2dq_vq_search_workspace_bytes(int N, int K);   /* 0: K < 2048, a shared row tile would not pay */
int b2dq_vq_search_gather(const void* x_bf16, const float* x_f32, const void* cb_bf16,
                          const float* cb_sqnorm, const float* weight_f32, const float* row_mask,
                          long long* codes, void* xq_bf16, float* xq_f32, float* loss_acc,
                          float* counts, float* sums, int N, int C, int K, int max_ctas,
                          void* workspace, int workspace_bytes, cudaStream_t stream);

/* EMA update of cluster_size_ema [K] / embed_ema [K,C], dead-code restart from restart_rows [K,C]
 * (rows the reference draws with randperm, :97), and weight[:K] = embed_ema / smoothed size.
 * dead_scratch [K] bytes, n_scratch [1] float. */
int b2dq_vq_ema_finalize(const float* counts, const float* sums, const float* restart_rows,
                         float* cluster_size_ema, float* embed_ema, float* weight_f32,
                         unsigned char* dead_scratch, float* n_scratch, int K, int C, float decay,
                         float eps, int restart, cudaStream_t stream);

/* g_x = g_xq + g_loss[0] * coef * mask * (x - xq);  coef = 2*beta/(N*C).  All [n_rows,C] bf16. */
int b2dq_vq_bwd(const void* g_xq, const void* x, const void* xq, const float* row_mask,
                const float* g_loss, float coef, void* g_x, long long n_rows, int C,
                cudaStream_t stream);

/* ------------------------------------------------------------------ convolutions as tap GEMMs
 * Replaces the cuDNN/cuBLAS calls behind nn.Conv2d at modules/diffusionmodules/model.py:43-47
 * (Upsample), :62-72 (Downsample, pad (0,1,0,1) + stride 2), :88-115 (ResnetBlock), :146-165
 * (AttnBlock 1x1), EncoderDual.py:41,72,83, DecoderPositional.py:62,91 and
 * dqvae_dual_feat.py:34-35 - forward and data gradient.
 *
 *   D[pixel, co] = alpha * sum_t sum_ci A[pixel + tap_t, ci] * B[co, tap_bk[t] + ci] + bias + residual
 *
 * A is addressed through a 5-D view (c, w, p, h, n) of a bf16 tensor; a tap is a coordinate offset
 * (tap_c, tap_w, tap_p, tap_h) in that view (out-of-range coordinates read zeros).  The output
 * pixel (n, oh, ow) is written at out + n*oN + oh*oH + ow*oW (+ channel), so parity classes of a
 * transposed stride-2 convolution can be written in place.
 */
typedef struct b2dq_tapgemm_desc {
  const void* a_ptr;
  long long a_dims[5];
  long long a_strides[5]; /* elements; a_strides[0] ignored (contiguous) */
  const void* b_ptr;      /* [b_batch][b_rows][b_k] bf16, K contiguous */
  long long b_rows, b_k, b_batch, b_batch_stride;
  int num_taps, kchunks;  /* kchunks = channels per tap / 64 */
  int tap_c[16], tap_w[16], tap_p[16], tap_h[16], tap_bk[16];
  int TW, TH, TN;         /* output tile, TW*TH*TN == 128 */
  int Wout, Hout, NB, Cout;
  void* out;
  long long oN, oH, oW;
  const float* bias;      /* [Cout] or NULL */
  const void* residual;   /* bf16, indexed with rN/rH/rW, or NULL */
  long long rN, rH, rW;
  float alpha;
  int out_f32;            /* 0: bf16 output, 1: fp32 output */
  int block_n;            /* 0 = auto (16/64/128/256) */
  int m_tiles_per_cta;    /* 0 = auto, 1 or 2 (two 128-pixel tiles share each weight tile) */
  int relu;               /* 1: clamp the result at zero (conv + ReLU of the VGG16 stack, lpips.py:88-97);
                           * 2: LeakyReLU(0.2) (PatchGAN stem, modules/discriminator/model.py:37) */
} b2dq_tapgemm_desc;

int b2dq_tapgemm(const b2dq_tapgemm_desc* desc, cudaStream_t stream);

/* Persistent variant for the wide 128-output-channel 3x3 stride-1 layers (W % 128 == 0, Cout == 128,
 * Cin % 64 == 0): one activation strip serves the three horizontal taps, two output tiles share
 * each weight tile, epilogue overlapped through double-buffered TMEM.  Same maths as b2dq_tapgemm
 * with the 3x3 tap table; dgrad != 0 mirrors the taps (data gradient).  b: [128, 9*Cin] bf16. */
int b2dq_pconv3x3(const void* a_bf16, const void* b_bf16, void* out_bf16, const float* bias,
                  const void* residual_bf16, float* gn_part, int NB, int H, int W, int Cin, int dgrad,
                  int max_ctas, cudaStream_t stream);
/* The same persistent kernel for any tap group that can be read from row strips: nr row taps (A-view channel base /
 * parity plane / row offset each) x ns column taps (pixel offsets col_dw[s], at most 2 apart), e.g. the 2x2 parity
 * classes of the folded up-convolution (model.py:49-53) and of the stride-2 data gradient (model.py:62-72), whose
 * four-tap K loop is too short for the one-shot b2dq_tapgemm.  Output tile grid NB x H x W (W % 128 == 0), 128 output
 * channels, pixel (n, h, w) written at out + n*oN + h*oH + w*oW (strides multiples of 8 elements). */
typedef struct b2dq_pconv_taps_desc {
  const void* a_ptr;
  long long a_dims[5];
  long long a_strides[5];  /* (c, w, p, h, n) view of the input, elements */
  const void* b_ptr;       /* [128][b_k] bf16 */
  long long b_k;
  int kchunks;             /* channels per tap / 64 */
  int nr, ns;              /* row taps (1..9), column taps (1..3) */
  int row_c[9], row_p[9], row_dh[9];
  int col_dw[3];
  int wcol[27];            /* weight column base of tap (r, s) at [r*3 + s] */
  int NB, H, W;
  void* out;
  long long oN, oH, oW;
  const float* bias;       /* [128] or NULL */
} b2dq_pconv_taps_desc;
int b2dq_pconv_taps(const b2dq_pconv_taps_desc* desc, int max_ctas, cudaStream_t stream);
/* gn_part (optional, [NB*H*W/128][32][2] floats): per-tile GroupNorm(32) partial sums of the OUTPUT, so the
 * next GroupNorm needs no statistics pass; b2dq_gn_finalize_tiles turns them into stats [NB][32][2]. */
int b2dq_gn_finalize_tiles(const float* gn_part, float* stats, int N, int H, int W, float eps,
                           cudaStream_t stream);

/* ------------------------------------------------------------------ batched / split-K GEMM
 * Weight gradients of the convolutions above (autograd of nn.Conv2d) and the attention
 * contractions of AttnBlock.forward (model.py:176-188) with their gradients.
 *   out[z][t][m][n] = alpha * sum_k A[m,k] * B_t[n,k]      z = batch*splits + split
 * Each operand is K-major ([rows][K]) or MN-major ([K][rows]); for MN-major operands the
 * contraction index is a 64-element box {KW,KH,KN} of a 5-D view and tap t shifts the B box.
 */
typedef struct b2dq_mm_desc {
  const void* a_ptr; long long a_dims[5]; long long a_strides[5];
  const void* b_ptr; long long b_dims[5]; long long b_strides[5];
  int a_mn, b_mn;
  int ntaps;              /* total B shifts (filter taps), 1..12 */
  int taps_per_cta;       /* accumulators per CTA, 1..3 (0 = auto); tap groups are spread over grid.x */
  int tap_c[12], tap_w[12], tap_p[12], tap_h[12];
  int KW, KH, KN;
  int ktiles_w, ktiles_h, kblocks;
  int splits, batches;
  int M, N;
  void* out; long long oZ, oT, oM;
  float alpha;
  int out_f32;
  int block_n;            /* 0 = auto, 128 or 256 */
  int b_strip;            /* 1: the 3 taps of a CTA are 1-pixel shifts (3x3 filter row): B is loaded once
                             per k-block as a 66-pixel strip (needs a_mn = b_mn = 1, KW = 64, KH = KN = 1) */
  float* colsum;          /* optional [splits][M] fp32 scratch: sum_k A[m,k] per split (the bias gradient when
                             A = dY), computed on the tensor core as A x ones; 3-tap MN-major-A launches only */
} b2dq_mm_desc;

int b2dq_mmgemm(const b2dq_mm_desc* desc, cudaStream_t stream);

/* out[m] = sum_splits part[split][m] (ordered) */
int b2dq_colsum_reduce(const float* part, float* out, int splits, int M, cudaStream_t stream);

/* partial [splits][taps][cout][cin] fp32 -> dw [cout][cin][taps] fp32 (OIHW); accumulate != 0: += */
int b2dq_wgrad_reduce(const float* partial, float* dw, int splits, int taps, int cout, int cin,
                      int accumulate, cudaStream_t stream);
/* the same plus the bias gradient db[cout] = sum_splits colsum[split][cout] (b2dq_colsum_reduce) in one launch */
int b2dq_wgrad_reduce_bias(const float* partial, float* dw, int splits, int taps, int cout, int cin,
                           const float* colsum, float* db, cudaStream_t stream);

/* Perceptual-loss feature stack (modules/losses/lpips.py:78-113, torchvision VGG16): 2x2/2 max pooling of
 * NHWC bf16 [N,H,W,C] (H, W even, C % 8 == 0) and its gradient (first maximum in window scan order gets
 * the gradient, like ATen), and the gradient of a ReLU given its OUTPUT y: dx = dy * (y > 0). */
int b2dq_maxpool2x2(const void* x_bf16, void* y_bf16, int N, int H, int W, int C, cudaStream_t stream);
int b2dq_maxpool2x2_bwd(const void* dy_bf16, const void* x_bf16, void* dx_bf16, int N, int H, int W, int C,
                        cudaStream_t stream);
int b2dq_relu_bwd(const void* dy_bf16, const void* y_bf16, void* dx_bf16, long long n, cudaStream_t stream);
/* gradient of LeakyReLU(slope) from its OUTPUT y (nn.LeakyReLU(0.2, True) of modules/discriminator/model.py:37-62) */
int b2dq_lrelu_bwd(const void* dy_bf16, const void* y_bf16, void* dx_bf16, long long n, float slope,
                   cudaStream_t stream);

/* LPIPS level head (lpips.py:44-55,116-122): f0, f1 = NHWC bf16 features [N,HW,C] (C in {64,128,256,512}),
 * w = lin weights [C].  fwd: part[n*chunks + j] = partial sums over pixel chunks of
 * sum_c w_c drop_c (f0_c/(|f0|+1e-10) - f1_c/(|f1|+1e-10))^2 (divide the per-image sum by HW for the spatial mean;
 * chunks = b2dq_lpips_head_chunks).  bwd: g[n] = gradient w.r.t. the spatial mean; writes the gradient w.r.t.
 * f0 and/or f1 (bf16, null = skip).  seed: device pointer to the dropout seed of this call, or null for no
 * dropout; p_drop = drop probability (nn.Dropout in front of the lin head, live in training mode). */
int b2dq_lpips_head_chunks(int N, int HW);
int b2dq_lpips_head_fwd(const void* f0_bf16, const void* f1_bf16, const float* w, float* part, int N, int HW, int C,
                        const unsigned long long* seed, float p_drop, cudaStream_t stream);
int b2dq_lpips_head_bwd(const void* f0_bf16, const void* f1_bf16, const float* w, const float* g, void* df0_bf16,
                        void* df1_bf16, int N, int HW, int C, const unsigned long long* seed, float p_drop,
                        cudaStream_t stream);

/* weight [Cout,Cin,R,S] fp32 (the nn.Conv2d parameter, model.py:43-47 etc.) -> the bf16 GEMM packings
 * fwd [Cout, R*S*Cin] and/or dgrad [Cin, R*S*Cout] in one pass (null = skip that packing). */
int b2dq_pack_weights(const float* weight, void* fwd, void* dgrad, int Cout, int Cin, int R, int S,
                      cudaStream_t stream);
/* Nearest x2 + 3x3 convolution (Upsample.forward, model.py:49-53) folded into four 2x2 convolutions of the
 * low-resolution input, one per output parity class (ph, pw): weight [Cout,Cin,3,3] fp32 -> fwd [Cout, 16*Cin] and
 * dgrad [Cin, 16*Cout] bf16 (null = skip), column block ((ph*2+pw)*2+a)*2+b holds the sum of the taps that land on
 * low-resolution slot (a, b) of that class.  b2dq_upconv_wgrad_reduce is the transpose for the weight gradient:
 * partial fp32 [4 classes][splits][4 slots][Cout][Cin] (the split-K partials of the four class GEMMs) ->
 * dw [Cout,Cin,3,3] fp32, splits added in index order. */
int b2dq_upconv_pack(const float* weight, void* fwd, void* dgrad, int Cout, int Cin, cudaStream_t stream);
int b2dq_upconv_wgrad_reduce(const float* partial, float* dw, int splits, int Cout, int Cin, cudaStream_t stream);
/* The same for many weights in one launch (all convolution weights are repacked after every optimizer step).
 * items_dev: device array of n_items records of six 64-bit words {weight pointer, fwd pointer or 0, dgrad pointer
 * or 0, Cout | Cin << 32, R*S | tiles_ci << 32, tile_start} with tiles_ci = ceil(Cin / 32) and tile_start the running
 * sum of ceil(Cout / 32) * tiles_ci over the preceding records; total_tiles = that sum over all records;
 * max_rs = the largest R*S among them (<= 16). */
int b2dq_pack_weights_multi(const void* items_dev, int n_items, long long total_tiles, int max_rs,
                            cudaStream_t stream);

/* out[c] = sum_rows dy[row][c]; deterministic two-stage sum, part = b2dq_bias_grad_blocks(rows)*C floats */
int b2dq_bias_grad_blocks(long long rows);
int b2dq_bias_grad(const void* dy_bf16, float* out, float* part, long long rows, int C, cudaStream_t stream);

/* ------------------------------------------------------------------ GroupNorm(32, eps) + swish
 * model.py:29-35 (Normalize, nonlinearity) as used at :119-127,170.
 * stats [N,G,2] = (mean, rstd);  chunks = b2dq_gn_chunks(N, HW);  ws: [N*chunks*G*2] floats;
 * part: [N*chunks*C*2] floats;  ws_nc: [N*C*2] floats;  dgb: [2*C] floats.  All sums are combined in a
 * fixed order (no atomics): results are bit-reproducible run to run. */
int b2dq_gn_chunks(int N, int HW);
int b2dq_gn_stats(const void* x, float* stats, float* ws, int N, int HW, int C, int G, float eps,
                  cudaStream_t stream);
int b2dq_gn_apply(const void* x, const float* stats, const float* gamma, const float* beta, void* y,
                  int N, int HW, int C, int G, int swish, cudaStream_t stream);
int b2dq_gn_bwd_stats(const void* dy, const void* x, const float* stats, const float* gamma,
                      const float* beta, float* part, float* ws_nc, int N, int HW, int C, int G,
                      int swish, cudaStream_t stream);
/* add (optional, bf16, same shape as dx): dx += add - the gradient arriving through a residual branch */
int b2dq_gn_bwd_apply(const void* dy, const void* x, const float* stats, const float* gamma,
                      const float* beta, const float* ws_nc, void* dx, float* dgb, const void* add,
                      int N, int HW, int C, int G, int swish, cudaStream_t stream);

/* dgb = NULL in b2dq_gn_bwd_apply skips the (dgamma, dbeta) reduction; b2dq_gn_bwd_param does it alone
 * (used when the backward is run in L2-sized image groups). */
int b2dq_gn_bwd_param(const float* ws_nc, float* dgb, int N, int C, cudaStream_t stream);

/* GroupNorm(+swish) backward as ONE persistent kernel (csrc/norm_fused.cu): 2 reads + 1 write of HBM instead of
 * the 5 passes of b2dq_gn_bwd_stats + b2dq_gn_bwd_apply.  Teams of CTAs own one image at a time, meet at a
 * per-image barrier in `ws` and re-read their slice of (dy, x) from L2.  Replaces the autograd of
 * Normalize + nonlinearity (modules/diffusionmodules/model.py:29-35).
 * b2dq_gn_bwd_fused_workspace_bytes: bytes of `ws` for an [N,HW,C] tensor (0 = shape not supported: use the pair
 * above).  dgb [2*C] = (dgamma, dbeta), overwritten.  add: optional, summed into dx.  Returns 0, a cudaError_t,
 * -1 (unsupported shape) or -2 (workspace too small).  b2dq_gn_bwd_fused_plan: out4 = {teams (images in flight),
 * CTAs per team, rows per CTA, grid}. */
int b2dq_gn_bwd_fused_workspace_bytes(int N, int HW, int C, int G);
int b2dq_gn_bwd_fused(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                      void* dx, float* dgb, const void* add, void* ws, long long ws_bytes, int N, int HW, int C,
                      int G, int swish, cudaStream_t stream);
int b2dq_gn_bwd_fused_plan(int N, int HW, int C, int* out4);
/* Forward with the same scheme: statistics + apply in one kernel (1 read + 1 write of HBM, second read from L2).
 * y = act(GroupNorm(x)) (swish: 0 none, 1 swish); stats [N][G][2] = (mean, rstd) is written for the backward. */
int b2dq_gn_fwd_fused_workspace_bytes(int N, int HW, int C, int G);
int b2dq_gn_fwd_fused(const void* x, const float* gamma, const float* beta, void* y, float* stats, void* ws,
                      long long ws_bytes, int N, int HW, int C, int G, float eps, int swish, cudaStream_t stream);

/* ------------------------------------------------------------------ layout / elementwise */
int b2dq_nchw_f32_to_nhwc_bf16(const float* src, void* dst, int N, int C, int HW, cudaStream_t stream);
int b2dq_nchw_f32_to_nhwc_f32(const float* src, float* dst, int N, int C, int HW, cudaStream_t stream);
int b2dq_nhwc_bf16_to_nchw_f32(const void* src, float* dst, int N, int C, int HW, cudaStream_t stream);
int b2dq_nhwc_f32_to_nchw_f32(const float* src, float* dst, int N, int C, int HW, cudaStream_t stream);
/* nearest x2 (model.py:50) and its gradient; in [N,H,W,C] bf16 */
int b2dq_upsample2x(const void* in, void* out, int N, int H, int W, int C, cudaStream_t stream);
int b2dq_upsample2x_bwd(const void* gout, void* gin, int N, int H, int W, int C, cudaStream_t stream);
/* row softmax (model.py:182): logits fp32 (in_f32) or bf16 -> probabilities bf16; and gradient */
int b2dq_softmax_rows(const void* s, void* p, long long rows, int T, int in_f32, cudaStream_t stream);
int b2dq_softmax_bwd_rows(const void* p, const void* dp, void* ds, long long rows, int T, float scale,
                          cudaStream_t stream);
int b2dq_add_bf16(const void* a, const void* b, void* out, long long n, cudaStream_t stream);
int b2dq_cast_f32_to_bf16(const float* a, void* out, long long n, cudaStream_t stream);
/* 3x3 window gather for the 3-channel edge convolutions: dst [N,H,W,64] */
int b2dq_im2col3x3_small(const void* src, void* dst, int N, int H, int W, int Cs, int flip,
                         cudaStream_t stream);
/* KxK window gather for the few-channel edge convolutions of the PatchGAN (modules/discriminator/model.py:37,66):
 * dst[n,oh,ow,(r*K+s)*Cs+c] = src[n, oh*stride + sgn*r + off, ow*stride + sgn*s + off, c], zero outside src and in
 * the unused columns; K*K*Cs <= 64.  Forward window: sgn=+1, off=-pad; data-gradient window: sgn=-1, off=+pad. */
int b2dq_im2col_window(const void* src, void* dst, int N, int Hs, int Ws, int Ho, int Wo, int Cs, int K, int stride,
                       int sgn, int off, cudaStream_t stream);


/* ------------------------------------------------------------------ operator-level entry points
 * One call per operator (csrc/oplevel.cu): the tile shapes, tap tables, split-K factors and the kernel choice
 * that dynamicvectorquantization_b200/kernels.py derives in Python are derived here in C from the operator
 * geometry, and the descriptor-level kernels above are enqueued.  Same inputs, bit-identical outputs on both
 * routes.  A binding in another host language needs only these.
 *
 * nn.Conv2d(Cin, Cout, ksize, stride) as the model uses it (modules/diffusionmodules/model.py:43-47,62-72,
 * 88-115,146-165): ksize 1 or 3; stride 1 with padding ksize/2, or stride 2 (ksize 3, H and W even) with
 * Downsample's (0,1,0,1) zero padding.  x [N,H,W,Cin], y [N,H/stride,W/stride,Cout] NHWC bf16.
 *   wpack        [Cout, ksize*ksize*Cin] bf16 (b2dq_pack_weights fwd), Cin % 64 == 0
 *   wpack_dgrad  [Cin, ksize*ksize*Cout] bf16 (b2dq_pack_weights dgrad), Cout % 64 == 0
 *   act          epilogue activation: 0 none, 1 ReLU, 2 LeakyReLU(0.2)
 *   gn_stats     optional [N,32,2] (mean, rstd) of GroupNorm(32, 1e-6) over y, produced by the convolution's own
 *                epilogue; only where b2dq_conv2d_fwd_workspace_bytes(g) > 0 (else pass NULL: -3 otherwise)
 *   dw           [Cout,Cin,ksize,ksize] fp32 (OIHW), overwritten;  db optional [Cout] fp32
 *   ws           scratch of the matching *_workspace_bytes (256 B aligned); -2 if too small */
typedef struct b2dq_conv2d_geom {
  int N, H, W, Cin, Cout, ksize, stride;
} b2dq_conv2d_geom;

int b2dq_conv2d_out_hw(const b2dq_conv2d_geom* g, int* out2 /* host: Hout, Wout */);
int b2dq_conv2d_fwd_workspace_bytes(const b2dq_conv2d_geom* g);
int b2dq_conv2d_fwd(const b2dq_conv2d_geom* g, const void* x, const void* wpack, const float* bias,
                    const void* residual, void* y, int act, float* gn_stats, void* ws, long long ws_bytes,
                    cudaStream_t stream);
int b2dq_conv2d_dgrad(const b2dq_conv2d_geom* g, const void* dy, const void* wpack_dgrad, void* dx,
                      cudaStream_t stream);
int b2dq_conv2d_wgrad_workspace_bytes(const b2dq_conv2d_geom* g, int want_bias);
int b2dq_conv2d_wgrad(const b2dq_conv2d_geom* g, const void* x, const void* dy, float* dw, float* db, void* ws,
                      long long ws_bytes, cudaStream_t stream);

/* GroupNorm(G, eps) + activation (0 none, 1 swish) of x [N,HW,C] bf16 and its backward (Normalize + nonlinearity,
 * model.py:29-35): the persistent fused kernels where the shape allows, else the statistics + apply pairs.
 * stats [N,G,2] is written by the forward and read by the backward; dgb [2,C] = (dgamma, dbeta); add: optional
 * bf16 tensor summed into dx.  backward = 0 / 1 selects which workspace size is returned. */
int b2dq_groupnorm_workspace_bytes(int N, int HW, int C, int G, int backward);
int b2dq_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, void* ws,
                       long long ws_bytes, int N, int HW, int C, int G, float eps, int act, cudaStream_t stream);
int b2dq_groupnorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                       void* dx, float* dgb, const void* add, void* ws, long long ws_bytes, int N, int HW, int C,
                       int G, int act, cudaStream_t stream);

/* The model's own pair - GroupNorm(32, eps 1e-6) + swish (ws sized with G = 32) - under the names of SURVEY 8b */
int b2dq_groupnorm_swish_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, void* ws,
                             long long ws_bytes, int N, int HW, int C, cudaStream_t stream);
int b2dq_groupnorm_swish_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const float* beta,
                             void* dx, float* dgb, const void* add, void* ws, long long ws_bytes, int N, int HW, int C,
                             cudaStream_t stream);

/* Attention core of AttnBlock.forward (model.py:176-188): out = softmax(scale * q k^T) v per image.
 * qkv [N,T,3C] bf16 holds q | k | v side by side (one [C -> 3C] 1x1 convolution); out [N,T,C] bf16; probs [N,T,T]
 * bf16 is kept for the backward, which writes dq | dk | dv into dqkv [N,T,3C].  scale = C^-0.5 in the model. */
int b2dq_attention_workspace_bytes(int N, int T, int C, int backward);
int b2dq_attention_fwd(const void* qkv, void* out, void* probs, void* ws, long long ws_bytes, int N, int T, int C,
                       float scale, cudaStream_t stream);
int b2dq_attention_bwd(const void* qkv, const void* probs, const void* dout, void* dqkv, void* ws,
                       long long ws_bytes, int N, int T, int C, float scale, cudaStream_t stream);

/* ------------------------------------------------------------------ entropy router input
 * Per-patch grey-level entropy (models/stage1_dynamic/dqvae_dual_entropy.py:25-63): x NCHW fp32
 * [B,3,H,W] in [-1,1] -> out [B, H/patch, W/patch] fp32.  Soft histogram over `nbins` (= 32) bin
 * centres `bins` (device, fp32) with Gaussian width sigma, eps 1e-40, -sum p log p. */
int b2dq_patch_entropy(const float* x_nchw, const float* bins, float* out, int B, int H, int W, int patch,
                       int nbins, float sigma, cudaStream_t stream);

/* ------------------------------------------------------------------ stage-2 tokenisation: dual-grain permuter
 * modules/dynamic_modules/permuter.py:50-109 (forward) and :111-132 (forward_back).  All tensors int64 on
 * the device.  forward: indices [B,F,F] codes, grain [B,Hc,Hc] (0 coarse / 1 fine) -> content / position /
 * segment rows of length coarse_len resp. fine_len (= longest sequence of the batch + 1 for the eos; the
 * caller sizes them), eos then pad filled.  codes6 is a HOST array {content_pad, content_eos,
 * coarse_position_pad, coarse_position_eos, fine_position_pad, fine_position_eos}.  region_first selects the
 * fine ordering (:78-96).  backward: the inverse map -> target [B,F,F]; duplicate positions resolve to the
 * last sequence element before the eos, positions outside the map are ignored (the reference raises). */
int b2dq_permuter_forward(const long long* indices, const long long* grain, long long* coarse_content,
                          long long* coarse_position, long long* coarse_segment, long long* fine_content,
                          long long* fine_position, long long* fine_segment, int B, int coarse_hw, int fine_hw,
                          int coarse_len, int fine_len, int region_first, const long long* codes6,
                          cudaStream_t stream);
int b2dq_permuter_backward(const long long* coarse_content, const long long* fine_content,
                           const long long* coarse_position, const long long* fine_position, long long* target,
                           int B, int coarse_hw, int fine_hw, int coarse_len, int fine_len,
                           long long coarse_position_eos, long long fine_position_eos, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* B200DQ_H_ */
