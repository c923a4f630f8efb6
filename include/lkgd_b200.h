/* lkgd_b200 - C ABI of the B200-native (sm_100a) kernels behind the LKGD / Stable-Video-Diffusion denoise
 * hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host".
 *
 * The reference (caoql98/LKGD) has no FFI: its extension point is Python module injection
 * (pipeline/pipeline_stable_video_diffusion_controlnet.py:122-143, run_models/run_inference.py:279-281).
 * Each entry point below replaces the ATen / cuDNN / cuBLAS calls that the cited reference lines (or the
 * diffusers==0.27.2 block they import) issue; the Python facade in lkgd_b200/ binds them with ctypes
 * (see INTEGRATION.md).
 *
 * Conventions: all launches go to the caller's `stream` (a cudaStream_t passed as void*); no hidden device
 * allocation; the caller owns every buffer; activations are channels-last bf16: a tensor [B,F,H,W,C] is a
 * row-major matrix [M = B*F*H*W, C].  Return value: 0 on success, <0 = LKGD_E*.
 */
#ifndef LKGD_B200_H_
#define LKGD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LKGD_ABI_VERSION 7

#if defined(__GNUC__)
#define LKGD_API __attribute__((visibility("default")))
#else
#define LKGD_API
#endif

#define LKGD_OK 0
#define LKGD_ESHAPE (-1) /* unsupported / inconsistent shape        */
#define LKGD_EALIGN (-2) /* pointer or pitch not 16-byte aligned    */
#define LKGD_EARCH (-3)  /* device is not sm_100                    */
#define LKGD_EWS (-4)    /* workspace too small                     */
#define LKGD_ECUDA (-5)  /* CUDA runtime / driver error (see lkgd_last_cuda_error) */

LKGD_API int lkgd_abi_version(void);
LKGD_API const char* lkgd_strerror(int code);
LKGD_API const char* lkgd_last_cuda_error(void);
/* 0 if device `dev` is a compute-capability-10.x part, LKGD_EARCH otherwise. */
LKGD_API int lkgd_device_check(int dev);
/* number of kernels this library has launched since load (bench.py's gpu_launches). */
LKGD_API uint64_t lkgd_launch_count(void);

/* ------------------------------------------------------------------------------------------------------
 * GEMM / implicit-GEMM convolution on tcgen05 tensor cores (TMA -> smem -> UMMA, fp32 accum in TMEM).
 *   out[m, n] = s0 * act( sum_k A[m,k] * Bw[n,k] + bias[n] + rowvec[g(m), n] ) + s1*res1[m,n] + s2*res2[m,n]
 * Replaces: nn.Linear projections / GEGLU feed-forward (diffusers attention.py, restated in the reference at
 * patch/patch.py:390-686), Conv2d 3x3 + Conv3d (3,1,1) of SpatioTemporalResBlock, down/up-sampler convs,
 * conv_in / conv_out (models/unet_spatio_temporal_condition_controlnet.py:127-132,240-245,431,500),
 * LoRA update (models/lora_layer.py:417-443) as a second K segment, AlphaBlender mix and residual adds as
 * epilogue terms.
 * A-operand modes:
 *   LINEAR  : A is [M, K0] row-major bf16 (pitch lda elements).
 *   CONV3X3 : A is NHWC [NIMG, Hin, Win, C0]; output pixel (ho,wo) gathers 3x3 taps with pad 1, stride 1 or 2;
 *             Bw is [N, 9*C0] with k = (ky*3+kx)*C0 + c.
 *   TCONV3  : A is [B, F, HW, C0]; output (b,f,p) gathers frames f-1,f,f+1 (zero pad); Bw is [N, 3*C0].
 * Optional second K segment (LoRA / shortcut / concat source): A1 [M, K1] (same addressing mode, centre tap),
 * Bw1 [N, K1].
 * Residuals: res1 / res2 rows must be 16-byte addressable (pointer, pitch and column count); res2 requires res1
 * and the same element type.
 * GEGLU: Bw rows are pre-interleaved per 256-row tile (128 value rows then their 128 gate rows); the output
 * has N/2 columns: out = (acc_h + b_h) * gelu_erf(acc_g + b_g).
 */
enum { LKGD_A_LINEAR = 0, LKGD_A_CONV3X3 = 1, LKGD_A_TCONV3 = 2 };
/* GELU: exact (erf) form; QUICK_GELU: x * sigmoid(1.702 x) (the CLIP vision tower's MLP, transformers activations.py) */
enum { LKGD_ACT_NONE = 0, LKGD_ACT_SILU = 1, LKGD_ACT_GEGLU = 2, LKGD_ACT_GELU = 3, LKGD_ACT_QUICK_GELU = 4 };
/* row -> rowvec index g(m) with HW = rows per frame, F frames, B = batch:
 *   NONE; FRAME: m/HW; FRAMEPOS: (m/HW)%F; BATCH: m/(HW*F);
 *   TCTX_0272: ((m/(HW*F))*HW + m%HW) % B  (diffusers 0.27.2 temporal-context quirk)
 *   BATCH_TCTX: (m/(HW*F))*B + TCTX_0272(m): a [B*B, N] table - the row's OWN sample picks the matrix (per-sample masked
 *               LoRA adapters on the temporal attn2 projections, patch/patch.py:57-92), the 0.27.2 rule picks the context */
enum { LKGD_RV_NONE = 0, LKGD_RV_FRAME = 1, LKGD_RV_FRAMEPOS = 2, LKGD_RV_BATCH = 3, LKGD_RV_TCTX_0272 = 4,
       LKGD_RV_BATCH_TCTX = 5 };

typedef struct lkgd_gemm_args {
  int32_t a_mode;     /* LKGD_A_*                                             */
  int32_t M, N;       /* output rows, Bw rows (GEGLU: N counts value+gate rows) */
  int32_t K0;         /* channels per tap of segment 0 (LINEAR: K)            */
  int32_t K1;         /* channels of segment 1, 0 = none                      */
  const void* A;      /* bf16                                                 */
  int32_t lda;        /* LINEAR: row pitch of A in elements (>= K0)           */
  const void* A1;     /* bf16 [M, K1] (or same layout as A with C = K1)       */
  int32_t lda1;
  const void* Bw;     /* bf16 [N, taps*K0], row pitch ldb                     */
  int32_t ldb;
  const void* Bw1;    /* bf16 [N, K1], row pitch ldb1                         */
  int32_t ldb1;
  /* conv geometry (CONV3X3: NIMG,Hin,Win,stride; TCONV3: B=NIMG, F, HW) */
  int32_t NIMG, Hin, Win, stride;
  int32_t F, HW;
  /* epilogue */
  const float* bias;  /* [N] or NULL                                          */
  const float* rowvec;/* [G, N] fp32 or NULL                                  */
  int32_t rv_mode, rv_HW, rv_F, rv_B;
  int32_t act;        /* LKGD_ACT_*                                           */
  float s0;
  const void* res1;   /* bf16 (or fp32, see res1_f32) [M, ldr1] or NULL      */
  int32_t ldr1;
  float s1;
  const void* res2;
  int32_t ldr2;
  float s2;
  void* out;          /* bf16 or fp32 [M, ldo]                                */
  int32_t ldo;
  int32_t out_f32;    /* 1: fp32 output                                       */
  int32_t n_store;    /* store only columns < n_store (0 = all)               */
  int32_t res1_f32;   /* 1: res1 is fp32 (the residual stream is kept in fp32) */
  int32_t res2_f32;
  int32_t rv_ld;      /* row pitch of rowvec in floats (0 = contiguous: N, or N/2 for GEGLU); multiple of 4 */
  /* Fused GroupNorm statistics of the stored output (fp32 outputs only): per (frame image, channel) sum and sum of
   * squares are ADDED to gn_stats [M / gn_rows][N][2] doubles (caller zeroes it), gn_rows = rows per frame image (H*W).
   * LINEAR mode needs gn_rows % 128 == 0.  The consumer is lkgd_groupnorm_from_stats: the GroupNorm that follows a
   * conv / projection (diffusers ResnetBlock2D.norm2, TemporalResnetBlock.norm1/2, the next block's norm1) no longer
   * re-reads the tensor for its statistics. */
  double* gn_stats;   /* NULL = off                                           */
  int32_t gn_rows;
  /* Optional bf16 copy of an fp32 output, [M, ldo2] (NULL = off; n_store % 8 == 0, ldo2 % 8 == 0): the tensor stays in the
   * fp32 residual stream AND is the A operand of the next GEMM without a separate narrowing pass - the ControlNet's skip
   * tensors feeding its zero convs (models/controlnet_sdv.py:558-571). */
  void* out2;
  int32_t ldo2;
  /* CONV3X3 with stride 2 only. 0: padding 1 on every side (diffusers Downsample2D(padding=1), the UNet). 1: the input is
   * padded on the bottom / right only, F.pad(x, (0, 1, 0, 1)) + Conv2d(stride=2, padding=0): diffusers
   * Downsample2D(padding=0) of the VAE encoder; output (Hin - 2) / 2 + 1 rows. */
  int32_t pad_br;
} lkgd_gemm_args;

LKGD_API int lkgd_gemm(const lkgd_gemm_args* args, void* stream);
/* Reference-quality SIMT implementation of the same contract (tests only: on-GPU checker at large sizes). */
LKGD_API int lkgd_gemm_simt_check(const lkgd_gemm_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) over channels-last data.  x is [NS, R, C1] (+ optional second source [NS, R, C2] that is
 * concatenated on the channel axis: torch.cat([hidden, skip], 1) of the up blocks), statistics per
 * (sample, group) over R * (C/G) elements.  Spatial resblock: NS = B*F, R = H*W.  Temporal resblock
 * (5-D GroupNorm, statistics ACROSS frames): NS = B, R = F*H*W.  Output bf16 [NS, R, C1+C2].
 * Replaces nn.GroupNorm + SiLU (diffusers resnet.py ResnetBlock2D / TemporalResnetBlock; the reference's
 * conv_norm_out + conv_act, ...controlnet.py:237-238,498-499; TransformerSpatioTemporalModel.norm).
 * x_f32 = 1: the sources are fp32 (the residual stream is kept in fp32); the output is always bf16 (a GEMM operand).
 * workspace: lkgd_groupnorm_workspace(NS, C) bytes.
 */
LKGD_API size_t lkgd_groupnorm_workspace(int32_t NS, int32_t C);
LKGD_API int lkgd_groupnorm(const void* x1, int32_t C1, const void* x2, int32_t C2, int32_t NS, int32_t R, int32_t groups,
                   const float* gamma, const float* beta, float eps, int32_t silu, int32_t x_f32, void* out,
                   void* workspace, size_t ws_bytes, void* stream);
/* Same normalisation with the statistics already accumulated by the producing lkgd_gemm launches (gn_stats):
 * stats1 / stats2 are [NS * frames_per_sample][C1 or C2][2] doubles (per frame image, channel); frames_per_sample = 1
 * for the spatial GroupNorms (NS = B*F) and F for the temporal ones (NS = B, statistics across frames).  One pass over
 * the tensor instead of two; the workspace is left exactly as lkgd_groupnorm leaves it (per-(sample, channel) sums first),
 * so lkgd_groupnorm_bwd can take it as fwd_sums.  raw_out (bf16 [NS, R, C1+C2], may be NULL) additionally receives the UN-normalised,
 * concatenated input narrowed to bf16: the operand of the resblock's 1x1 conv_shortcut, for free in the same pass. */
LKGD_API int lkgd_groupnorm_from_stats(const void* x1, int32_t C1, const double* stats1, const void* x2, int32_t C2,
                   const double* stats2, int32_t NS, int32_t R, int32_t frames_per_sample, int32_t groups,
                   const float* gamma, const float* beta, float eps, int32_t silu, int32_t x_f32, void* out,
                   void* raw_out, void* workspace, size_t ws_bytes, void* stream);

/* LayerNorm over the last axis of a [M, C] bf16 matrix (C <= 2048, C % 8 == 0) with optional fused
 *   s = x + addvec[g(m)]   (fp32 addvec [G, C], row pitch addvec_ld floats (0 = C); frame positional embedding or
 *                           KV-length-1 cross-attention term)
 * sum_out (same dtype as x: bf16, or fp32 when x_f32 = 1; may alias x, may be NULL) receives s; out (bf16) receives
 * LN(s)*gamma+beta.
 * Replaces nn.LayerNorm norm1/norm2/norm3/norm_in (patch/patch.py:415-416,529-530,555-556,599,610,664,670). */
LKGD_API int lkgd_layernorm(const void* x, int32_t M, int32_t C, const float* gamma, const float* beta, float eps,
                   const float* addvec, int32_t addvec_ld, int32_t rv_mode, int32_t rv_HW, int32_t rv_F, int32_t rv_B,
                   int32_t x_f32, void* sum_out, void* out, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Spatial self-attention (and general cross-attention), flash-style on tcgen05: per (image, head)
 *   O = softmax(Q K^T * scale) V,  non-causal, no mask.
 * q/k/v point at the first head's first element; consecutive heads are `d` elements apart inside a token row;
 * token rows are ld{q,k,v} elements apart; images are Nq (resp. Nk) rows apart.  d % 8 == 0, d <= 128 (d <= 64: three CTAs
 * per SM; 64 < d <= 128: two 64-channel sub-tiles, two CTAs per SM - the reference-default UNet heads (5,10,10,20) give d = 128,
 * the CLIP ViT-H image encoder d = 80).
 * Replaces F.scaled_dot_product_attention in diffusers AttnProcessor2_0 for transformer_blocks.*.attn1/attn2. */
LKGD_API int lkgd_attention(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv, void* out,
                   int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq, int32_t Nk, float scale,
                   void* stream);
/* SIMT checker with the same contract (tests only). */
LKGD_API int lkgd_attention_simt_check(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                              void* out, int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq,
                              int32_t Nk, float scale, void* stream);

/* Temporal self-attention over the frame axis without a transpose: qkv is the fused projection
 * [B, F, HW, 3*C] (q | k | v, C = heads*d); sequence (b, p, head) attends over its F frames (F <= 32).
 * out is [B, F, HW, C].  Replaces temporal_transformer_blocks.*.attn1 (patch/patch.py:592-597,659-661). */
LKGD_API int lkgd_attention_temporal(const void* qkv, void* out, int32_t B, int32_t F, int32_t HW, int32_t heads, int32_t d,
                            float scale, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Small fp32 helpers of the conditioning path (M <= 64 rows).
 *   y[m, n] = act_out( sum_k act_in(x[m,k]) * W[n,k] + b[n] )     W fp32 [N, K]
 * act codes: 0 none, 1 SiLU, 3 LeakyReLU(0.1).
 * Replaces TimestepEmbedding MLPs, resnet time_emb_proj, KV-length-1 cross-attention to_v / to_out
 * (...controlnet.py:406-419) and the LKGD fuse MLPs (unet_spatio_temporal_condition.py:595). */
LKGD_API int lkgd_small_linear(const float* x, int32_t ldx, const float* W, const float* b, float* y, int32_t ldy, int32_t M,
                      int32_t N, int32_t K, int32_t act_in, int32_t act_out, void* stream);
/* Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): out[m] = [cos(t_m f_k) | sin(t_m f_k)]. */
LKGD_API int lkgd_timestep_embedding(const float* t, int32_t M, int32_t dim, float* out, void* stream);

/* y += alpha * x on fp32 (embedding sums). */
LKGD_API int lkgd_axpy_f32(const float* x, float alpha, float* y, int64_t n, void* stream);
/* y = alpha * x on fp32 (scheduler.scale_model_input, utils/scheduling_euler_discrete_karras_fix.py:284-285). */
LKGD_API int lkgd_scale_f32(const float* x, float alpha, float* y, int64_t n, void* stream);
/* mode 0: (a,b) = (re,im) -> o0 = |z|, o1 = atan2(im,re);  mode 1: (a,b) = (mag,pha) -> o0 = mag cos, o1 = mag sin.
 * The torch.abs / torch.angle / cos / sin of the LKGD spectral fuse (unet_spatio_temporal_condition.py:559-580). */
LKGD_API int lkgd_polar(const float* a, const float* b, float* o0, float* o1, int32_t n, int32_t mode, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Layout / glue kernels (HBM-bound, 128-bit vectorised).
 */
/* out[n,f,h,w,0:C0] = src0[n % N0, f, :, h, w] * scale0 ; out[..., C0:C0+C1] = src1[n % N1, ...]; rest 0.
 * src are fp32 NCHW-per-frame [N?, F, C?, H, W]; out is bf16 [N, F, H, W, Cpad].  Fuses the CFG duplication,
 * scheduler.scale_model_input and the image-latent concat (pipeline...controlnet.py:579-584).
 * scale0_dev (may be NULL): DEVICE scalar that replaces scale0 - the per-step 1/sqrt(sigma^2+1) of a denoise step
 * captured once in a CUDA graph and replayed for every sigma. */
LKGD_API int lkgd_pack_input(const float* src0, int32_t N0, int32_t C0, float scale0, const float* scale0_dev,
                    const float* src1, int32_t N1, int32_t C1, void* out, int32_t N, int32_t F, int32_t H, int32_t W,
                    int32_t Cpad, void* stream);
/* src fp32 [N*F, H, W, ld] channels-last -> dst fp32 [N, F, C, H, W]. */
LKGD_API int lkgd_unpack_output(const float* src, int32_t ld, float* dst, int32_t NF, int32_t C, int32_t H, int32_t W,
                       void* stream);
/* generic NCHW fp32 [N, C, H, W] <-> NHWC bf16 [N, H, W, C] converters (ControlNet residual exchange). */
LKGD_API int lkgd_nchw_to_nhwc(const float* src, void* dst, int32_t N, int32_t C, int32_t H, int32_t W, void* stream);
LKGD_API int lkgd_nhwc_to_nchw(const void* src, float* dst, int32_t N, int32_t C, int32_t H, int32_t W, void* stream);
/* nearest 2x upsample, channels-last [N,H,W,C] (bf16, or fp32 when src_f32) -> bf16 [N,2H,2W,C]
 * (Upsample2D before its conv; the fp32 residual stream is narrowed to the conv's bf16 operand on the way). */
LKGD_API int lkgd_upsample2x(const void* src, int32_t src_f32, void* dst, int32_t N, int32_t H, int32_t W, int32_t C,
                    void* stream);
/* fp32 -> bf16 narrowing of a residual-stream tensor that a GEMM reads raw (shortcut / downsampler convs). */
LKGD_API int lkgd_cast_bf16(const float* src, void* dst, int64_t n, void* stream);
/* channel concat of two channels-last matrices (bf16, or fp32 when src_f32) into bf16:
 * dst[m] = [a[m, 0:Ca] | b[m, 0:Cb]]. */
LKGD_API int lkgd_concat_channels(const void* a, int32_t Ca, const void* b, int32_t Cb, int32_t src_f32, void* dst,
                         int64_t M, void* stream);
/* y[i] = alpha * x[i] + beta * y[i]; x / y each bf16 or fp32 (ControlNet residual injection with the F6
 * multipliers: ...controlnet.py:453-462,472-473). n % 8 == 0. */
LKGD_API int lkgd_axpby(const void* x, int32_t x_f32, float alpha, void* y, int32_t y_f32, float beta, int64_t n,
               void* stream);

/* Thin 3x3 convolutions (pad 1, stride 1) of the ControlNet condition encoder at pixel resolution
 * (models/controlnet_sdv.py:64-119, `ControlNetConditioningEmbeddingSVD`): HBM-bound layers with 2-32 channels.
 *   lkgd_cond_conv_in : x fp32 planar [N, Cc, H, W] (Cc <= 4; the reference's `controlnet_cond` flattened over batch and
 *                       frames), weight fp32 [16, Cc, 3, 3], bias fp32 [16] -> out bf16 channels-last [N, H, W, 16] = SiLU(conv)
 *                       (`conv_in` + `F.silu`, :104-105; also replaces the fp32 -> bf16 channels-last pack).
 *   lkgd_thin_conv3x3 : x bf16 [N, H, W, Cin], weight bf16 [9, Cout, Cin] (tap-major: weight.permute(2,3,0,1)), bias fp32
 *                       [Cout] -> out bf16 [N, H, W, Cout], optional SiLU; (Cin, Cout) in {(16,16), (16,32), (32,32)}
 *                       (the stride-1 `blocks[2i]` convs + `F.silu`, :107-109). */
LKGD_API int lkgd_cond_conv_in(const float* x, int32_t N, int32_t Cc, int32_t H, int32_t W, const float* weight,
                      const float* bias, void* out, void* stream);
LKGD_API int lkgd_thin_conv3x3(const void* x, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, const void* weight,
                      const float* bias, int32_t silu, void* out, void* stream);

/* Non-overlapping patch unfold for the ViT patch embedding (CLIPVisionEmbeddings.patch_embedding, a Conv2d with
 * kernel = stride = P, no bias): x fp32 [N, C, H, W] -> out bf16 [N * (H/P) * (W/P), Kpad] with column (c * P + py) * P + px
 * (the Conv2d weight's own flattening), zero beyond C*P*P; the embedding itself is then one lkgd_gemm. */
LKGD_API int lkgd_patchify(const float* x, int32_t N, int32_t C, int32_t H, int32_t W, int32_t P, void* out, int32_t Kpad,
                  void* stream);

/* ------------------------------------------------------------------------------------------------------
 * VAE (diffusers AutoencoderKLTemporalDecoder, un-vendored; call sites pipeline/pipeline_stable_video_diffusion_controlnet.py
 * :216-237 `_encode_vae_image`, :268-295 `decode_latents`).  Convolutions, GroupNorms and projections run on lkgd_gemm /
 * lkgd_groupnorm; these two cover what those do not:
 *   lkgd_softmax_rows  : out[m, n] = softmax_n(scale * x[m, n]); x fp32 [M, N] (row pitch ldx), out bf16 (row pitch ldo),
 *                        N % 4 == 0, N <= 16384.  The mid-block `Attention` has ONE 512-wide head (heads = C / 512), which does
 *                        not fit the flash kernel's TMEM budget: it runs as Q K^T (lkgd_gemm, fp32 out) -> this -> P V^T.
 *   lkgd_time_conv_out : the decoder's final Conv3d(C, C, (3,1,1), padding (1,0,0)) over the frame axis, fused with the
 *                        channels-last -> planar unpack: x fp32 rows [NB*F*HW, ldx] (first C columns), weight fp32 [C, C, 3]
 *                        (= Conv3d weight[..., 0, 0]), bias fp32 [C] -> out fp32 [NB*F, C, H*W]; C in {1, 3, 4}. */
LKGD_API int lkgd_softmax_rows(const float* x, int64_t ldx, int64_t M, int32_t N, float scale, void* out, int64_t ldo,
                      void* stream);
LKGD_API int lkgd_time_conv_out(const float* x, int32_t ldx, const float* weight, const float* bias, float* out, int32_t NB,
                       int32_t F, int64_t HW, int32_t C, void* stream);

/* out[m, :] = srcs[g(m)][m, :] over bf16 [M, C] matrices (C % 8 == 0; srcs = HOST array of n_src <= 8 device pointers, g as
 * for the row vectors above).  Temporal cross-attention with KV length > 1 under the diffusers 0.27.2 context order: row m
 * of the temporal batch attends to context g(m) = TCTX_0272 (transformer_temporal.py `time_context` broadcast), so
 * the attention is evaluated once per context and this kernel keeps, per row, the result of its own context. */
LKGD_API int lkgd_select_rows(const void* const* srcs, int32_t n_src, void* out, int64_t M, int32_t C, int32_t rv_mode,
                     int32_t rv_HW, int32_t rv_F, int32_t rv_B, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Fused classifier-free-guidance combine + Euler (Karras sigmas, v-prediction) step, fp32.
 *   v      = u + g[f] * (c - u)           u = pred[s], c = pred[S + s]   (pred channels-last fp32 [2S*F,H,W,ld])
 *   x0     = v * (-sigma / sqrt(sigma^2+1)) + x / (sigma^2+1)
 *   x_next = x + (x - x0) / sigma * (sigma_next - sigma)
 * x / x_next are fp32 [S, F, C, H, W] (reference layout).  cfg == 0: v = pred.  ld == 0: pred is laid out like x
 * ([2S or S, F, C, H, W], the reference's own layout) instead of channels-last rows.
 * v_out / x0_out (may be NULL) receive the guided prediction and x0 (`pred_original_sample` of the reference's
 * scheduler output, :506,:527).  x_next may alias x.  sigmas_dev (may be NULL): DEVICE [sigma, sigma_next] replacing the
 * two host scalars (CUDA-graph replay).
 * Replaces pipeline...controlnet.py:614-616 and utils/scheduling_euler_discrete_karras_fix.py:481-520. */
LKGD_API int lkgd_cfg_euler_step(const float* pred, int32_t ld, int32_t cfg, const float* guidance, const float* x,
                        float* x_next, float* v_out, float* x0_out, int32_t S, int32_t F, int32_t C, int32_t H,
                        int32_t W, float sigma, float sigma_next, const float* sigmas_dev, void* stream);

/* The same step with the two halves of the CFG batch at SEPARATE bases (each fp32 channels-last rows [S*F, H, W, ld]):
 * the CFG-pair split (two GPUs share one sample) keeps every rank's prediction in peer-mapped memory and
 * passes its own buffer for one half and the partner's mapping for the other - the kernel reads the partner's half over
 * NVLink, so exchange, guidance combine (pipeline...controlnet.py:614-616) and Euler step are ONE kernel and no
 * all-gather runs (ABI v7). */
LKGD_API int lkgd_cfg_euler_step_pair(const float* pred_uncond, const float* pred_cond, int32_t ld, const float* guidance,
                             const float* x, float* x_next, float* v_out, float* x0_out, int32_t S, int32_t F,
                             int32_t C, int32_t H, int32_t W, float sigma, float sigma_next, const float* sigmas_dev,
                             void* stream);

/* Bidirectional "direct fusion" Euler step (pipeline/pipeline_stable_video_diffusion_trans_controlnet.py:639-667,
 * SURVEY 8f N3): v and x are fp32 [2S, F, C, H, W] (forward samples, then their time-reversed partners), weights fp32
 * [F] = linspace(1, 0, F).  x0 = v * (-sigma / sqrt(sigma^2+1)) + x / (sigma^2+1) per half; the forward half keeps
 * w[f] x0_fwd[f] + (1-w[f]) x0_bwd[F-1-f], the backward half its frame flip; x_next = x + (x - x0) / sigma *
 * (sigma_next - sigma).  One launch instead of ~15 elementwise / flip / cat ATen kernels. */
LKGD_API int lkgd_fusion_euler_step(const float* v, const float* x, const float* weights, float* x_next, int32_t S,
                   int32_t F, int32_t C, int32_t H, int32_t W, float sigma, float sigma_next, void* stream);

/* ======================================================================================================
 * Training step (LoRA fine-tuning): the backward of the path above.  Reference: train_models/train_svd_lora.py
 * :1445-1689 - EDM preconditioning :1503-1530, loss :1651-1672, accelerator.backward :1683 (PyTorch autograd of the
 * diffusers blocks), clip_grad_norm_ :1684-1686, AdamW :1225-1231, DDP gradient all-reduce :1300-1302.
 * Data gradients of the GEMMs / convolutions reuse lkgd_gemm with transposed (and, for convolutions, tap-flipped)
 * weights; the entries below are the remaining backward kernels.
 * ====================================================================================================== */

/* Spatial attention forward that also returns the log-sum-exp the backward needs:
 * lse[img, head, row] = log2( sum_k 2^(s_k * scale * log2 e) )  (fp32, [n_img, heads, Nq]). */
LKGD_API int lkgd_attention_lse(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                void* out, int32_t ldo, int32_t n_img, int32_t heads, int32_t d, int32_t Nq, int32_t Nk,
                                float scale, float* lse, void* stream);
/* Backward of O = softmax(Q K^T scale) V per (image, head), self-attention (Nq = Nk = N), d in {16,32,64,128}.
 * o / dO share the pitch ldo; dq / dk / dv may be column slices of one fused [rows, 3C] gradient.
 * workspace: lkgd_attention_bwd_workspace(n_img, heads, N) bytes. */
LKGD_API size_t lkgd_attention_bwd_workspace(int32_t n_img, int32_t heads, int32_t N);
LKGD_API int lkgd_attention_bwd(const void* q, int32_t ldq, const void* k, int32_t ldk, const void* v, int32_t ldv,
                                const void* o, const void* dO, int32_t ldo, const float* lse, void* dq, int32_t lddq,
                                void* dk, int32_t lddk, void* dv, int32_t lddv, int32_t n_img, int32_t heads, int32_t d,
                                int32_t N, float scale, void* workspace, size_t ws_bytes, void* stream);
/* Backward of lkgd_attention_temporal: qkv [B,F,HW,3C], dO [B,F,HW,C] -> dqkv [B,F,HW,3C] (all bf16). */
LKGD_API int lkgd_attention_temporal_bwd(const void* qkv, const void* dO, void* dqkv, int32_t B, int32_t F, int32_t HW,
                                         int32_t heads, int32_t d, float scale, void* stream);

/* GroupNorm(+SiLU) backward.  x1/x2/NS/R/groups/gamma/beta/eps/silu/x_f32 as in lkgd_groupnorm; fwd_sums is the
 * workspace the forward call filled (per-(sample, channel) sum and sum of squares); dy is bf16 [NS*R, C].
 *   dx = GN'(dy) [+ add]      add: optional [NS*R, C] (bf16, or fp32 when add_f32)
 * written to  out1 (fp32 [NS*R, C1], += when acc1), out2 (fp32 [NS*R, C2], += when acc2; the concatenated skip's
 * share) and / or out_bf16 (bf16 [NS*R, C], the value stored in out1/out2 after accumulation) - each may be NULL. */
LKGD_API size_t lkgd_groupnorm_bwd_workspace(int32_t NS, int32_t C);
LKGD_API int lkgd_groupnorm_bwd(const void* x1, int32_t C1, const void* x2, int32_t C2, int32_t NS, int32_t R,
                                int32_t groups, const float* gamma, const float* beta, float eps, int32_t silu,
                                int32_t x_f32, const void* dy, const void* fwd_sums, const void* add, int32_t add_f32,
                                float* out1, int32_t acc1, float* out2, int32_t acc2, void* out_bf16, void* workspace,
                                size_t ws_bytes, void* stream);
/* LayerNorm backward: x fp32 [M, C] (the normalised input), dy [M, C] (bf16, or fp32 when dy_f32);
 * G (fp32 [M, C]) = (accumulate ? G : 0) + LN'(dy); g_bf16 (optional) receives the new G as bf16. */
LKGD_API int lkgd_layernorm_bwd(const float* x, const void* dy, int32_t dy_f32, int32_t M, int32_t C, const float* gamma,
                                float eps, float* G, int32_t accumulate, void* g_bf16, void* stream);
/* GEGLU on the tile-interleaved projection output pre [M, 2H] (lkgd_gemm GEGLU weight order, act NONE):
 * fwd: out[m, j] = value * gelu_erf(gate);  bwd: dpre from dout [M, H]. */
LKGD_API int lkgd_geglu_fwd(const void* pre, void* out, int64_t M, int32_t H, void* stream);
LKGD_API int lkgd_geglu_bwd(const void* pre, const void* dout, void* dpre, int64_t M, int32_t H, void* stream);
/* out[g(m), c] += G[m, c]: gradient of the per-context vectors added by LayerNorm's addvec (rv modes as above;
 * n_groups <= 8; out must be zero-initialised by the caller, fp32 [n_groups, C] with row pitch ldo). */
LKGD_API int lkgd_colsum_grouped(const float* G, int64_t M, int32_t C, int32_t rv_mode, int32_t rv_HW, int32_t rv_F,
                                 int32_t rv_B, int32_t n_groups, float* out, int64_t ldo, void* stream);
/* Nearest-2x upsample backward: out fp32 [N,H,W,C] = 2x2 block sums of in [N,2H,2W,C] (bf16 / fp32). */
LKGD_API int lkgd_downsum2x(const void* in, int32_t in_f32, float* out, int32_t N, int32_t H, int32_t W, int32_t C,
                            void* stream);
/* Stride-2 conv data gradient helper: out bf16 [N,Hin,Win,C], out[2ho,2wo] = in[ho,wo] (in [N,Ho,Wo,C]), else 0. */
LKGD_API int lkgd_zero_stuff2x(const void* in, int32_t in_f32, void* out, int32_t N, int32_t Hin, int32_t Win, int32_t C,
                               void* stream);
/* out[i, j] += alpha * sum_m X[m, i] * Y[m, j]: LoRA weight gradients (models/lora_layer.py:437 under autograd).
 * X / Y rows are read in 8-column pieces: when I or J is not a multiple of 8 the rows must be readable up to the next
 * multiple (a column slice of a wider, padded matrix). */
LKGD_API int lkgd_gemm_tn(const void* X, int64_t ldx, int32_t I, const void* Y, int64_t ldy, int32_t J, int64_t M,
                          float alpha, float* out, int64_t ldo, void* stream);
/* EDM training wrapper (train_svd_lora.py:1503-1530): noisy = latents + noise * sigma[b] (fp32 [B,F,C,H,W]);
 * x_in (bf16 rows [B*F*H*W, Cpad]) = [noisy / sqrt(sigma^2+1) | cond[b] (fp32 [B,C,H,W], repeated over frames) | 0]. */
LKGD_API int lkgd_edm_precondition(const float* latents, const float* noise, const float* sigma, const float* cond,
                                   float* noisy, void* x_in, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W,
                                   int32_t Cpad, void* stream);
/* Weighted-MSE loss of the v-prediction wrapper and its gradient (train_svd_lora.py:1651-1672):
 *   den = pred * c_out + c_skip * noisy;  loss = mean_b mean_{f,c,h,w} (1+s^2)/s^2 (den - target)^2   (double, device)
 *   dpred (bf16 rows [B*F*H*W, Cpad], zero padded) = grad_scale * dloss/dpred.  pred: fp32 rows, pitch ld. */
LKGD_API int lkgd_edm_loss(const float* pred, int32_t ld, const float* noisy, const float* target, const float* sigma,
                           double* loss, void* dpred, int32_t B, int32_t F, int32_t C, int32_t H, int32_t W, int32_t Cpad,
                           float grad_scale, void* stream);
/* *out = sum x^2 (double, device): clip_grad_norm_ (train_svd_lora.py:1684-1686). */
LKGD_API int lkgd_sumsq(const float* x, int64_t n, double* out, void* stream);
/* torch.optim.AdamW step over flat fp32 buffers; the gradient is scaled by grad_scale (1/world after the all-reduce)
 * and clipped to max_norm using *sumsq (both optional: sumsq NULL or max_norm <= 0 disables clipping). */
LKGD_API int lkgd_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                        float eps, float weight_decay, int32_t step, float grad_scale, const double* sumsq,
                        float max_norm, void* stream);
/* dst[r*ldd + c] = bf16(alpha * src[r*lds + c*src_cs]) (src_cs = 1: plain; = pitch of a transposed view) (repacking fp32 master LoRA weights into the GEMM operands,
 * scaled bf16 copies of gradient tensors). */
LKGD_API int lkgd_cast2d_bf16(const float* src, int64_t lds, int64_t src_cs, void* dst, int64_t ldd, int32_t rows,
                              int32_t cols, float alpha, void* stream);
/* The same cast for a TABLE of jobs in ONE launch (ABI v7): after every optimizer step the trainer rewrites four small
 * bf16 operands per adapter (A, A^T, s B, s B^T - the repack after train_svd_lora.py:1687 optimizer.step()), 192
 * launches of a few microseconds each for the 48 adapters of the reference configuration.  ``jobs`` is a DEVICE array of
 * n_jobs descriptors (built once: the master parameters and the operands never move); max_elems >= rows * cols of every
 * job. */
typedef struct lkgd_cast2d_job {
  const float* src;
  int64_t lds;      /* row pitch of src in floats                        */
  int64_t src_cs;   /* column stride of src (1, or a pitch: transposed view) */
  void* dst;        /* bf16                                              */
  int64_t ldd;      /* row pitch of dst in elements                      */
  int32_t rows, cols;
  float alpha;
  int32_t reserved;
} lkgd_cast2d_job;
LKGD_API int lkgd_cast2d_bf16_batch(const lkgd_cast2d_job* jobs, int32_t n_jobs, int32_t max_elems, void* stream);


/* ---- backward of the fp32 conditioning helpers (latent-knowledge block under autograd,
 * models/unet_spatio_temporal_condition.py:536-595; trainable 'quaternion' parameters, train_svd_lora.py:1068-1073).
 * lkgd_small_linear backward with dy' = dy * act_out'(y) (act_out 0 or 3 = LeakyReLU(0.1), y = the forward output):
 *   dx[m,k] (=|+=) sum_n dy'[m,n] W[n,k]     (dx NULL: skipped)
 *   dW[n,k] += sum_m dy'[m,n] x[m,k];  db[n] += sum_m dy'[m,n]     (each may be NULL) */
LKGD_API int lkgd_small_linear_bwd(const float* dy, int32_t lddy, const float* y, int32_t ldy, int32_t act_out,
                                   const float* x, int32_t ldx, const float* W, float* dx, int32_t lddx,
                                   int32_t dx_accumulate, float* dW, float* db, int32_t M, int32_t N, int32_t K,
                                   void* stream);
/* backward of lkgd_polar: mode 0 (a,b) = (re,im), (d0,d1) = (d mag, d pha) -> (o0,o1) = (d re, d im);
 * mode 1 (a,b) = (mag,pha), (d0,d1) = (d re, d im) -> (o0,o1) = (d mag, d pha). */
LKGD_API int lkgd_polar_bwd(const float* a, const float* b, const float* d0, const float* d1, float* o0, float* o1,
                            int32_t n, int32_t mode, void* stream);
/* Conv1d(4G -> G, kernel 1, groups G) weight gradient: dw[j, m] += sum_b dy[b, j] * x[b, 4j + m]. */
LKGD_API int lkgd_grouped1x1_bwd_w(const float* dy, int32_t lddy, const float* x, int32_t ldx, float* dw, int32_t B,
                                   int32_t G, void* stream);
/* Folds the dense gradient dWt [out, in] of a quaternion linear layer back onto its r / i / j / k components
 * ([in/4, out/4] each, +=). */
LKGD_API int lkgd_hamilton_bwd(const float* dWt, int32_t in, int32_t out, float* dr, float* di, float* dj, float* dk,
                               void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LKGD_B200_H_ */
