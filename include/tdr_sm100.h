/* tdr_sm100.h -- C ABI of libtdr_sm100.so: the B200 (sm_100a) kernels beneath the restoration-network hot path
 * of mrluin/TextualDegRemoval.
 *
 * The reference has no FFI of its own: every op below replaces a stock PyTorch call inside
 * /root/reference/models/archs/network_restormer_guided_arch.py (cited per entry point as file:line, "R:" prefix)
 * or network_nafnet_guided_arch.py ("N:" prefix).  The Python modules in textualdegremoval_b200/archs bind these
 * symbols with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless named host_*; the caller owns every
 *     buffer (including workspaces, sized by the *_workspace_bytes helpers); the library never allocates or syncs.
 *   - activations are NHWC ("pixel rows x channels"): bf16 for GEMM operands, fp32 for the residual stream.
 *     `ld` arguments are row strides in ELEMENTS (channels of the underlying buffer), so channel slices and
 *     concatenations are expressed by pointer offset + ld, never by copies.
 *   - every function returns TDR_OK (0) or a negative TDR_E* code; tdr_last_error() gives the message for the
 *     calling thread.  Work is enqueued on `stream` (a cudaStream_t passed as void*) and is asynchronous.
 */
#ifndef TDR_SM100_H_
#define TDR_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDR_OK 0
#define TDR_EINVAL (-1) /* bad argument (shape, alignment, null pointer) */
#define TDR_ECUDA (-2)  /* CUDA runtime / driver error */
#define TDR_ENOSUP (-3) /* configuration not supported by the sm_100a kernels */

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

const char* tdr_last_error(void);
int tdr_version(void);
/* 0 if the current device is sm_100 (B200), TDR_ENOSUP otherwise. */
int tdr_check_device(void);
/* Programmatic dependent launch of the hot kernels (their prologues overlap the previous kernel's tail).  Off by default
 * (measured neutral); TDR_PDL=1 in the environment or tdr_set_pdl(1) turns it on.  Sets the switch, returns the previous
 * value.  Results do not depend on it. */
int tdr_set_pdl(int on);

/* ---------------------------------------------------------------------------------------------------------------
 * Implicit-GEMM convolution (tcgen05 + TMA).  Replaces nn.Conv2d for: 1x1 convs R:229,234,252,254,613,619;
 * dense 3x3 convs R:38-39,106-116 (stride 1/2, bias, ReLU, residual), R:376 + PixelUnshuffle, R:387 + PixelShuffle;
 * `attn @ v` + project_out R:272-276 (per-sample weights); MASA correlations R:683-692, R:661-669 (per-sample
 * filters, dilation, per-pixel scale, window origins).
 *   out = g * alpha * act( acc * rowscale[pixel] + bias[co] ) + g * res1_scale * res1 + res2,   g = scale_ptr ? *scale_ptr : 1
 * With H == 1 the op is a plain row-major GEMM over [W rows, Ci] (ViT / mapper linears, q.k^T, p.v).
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct tdr_conv_gemm_desc {
  const void* in; /* bf16 [B_img, H, W, in_ld], channels [0, Ci) used */
  long long in_ld;
  int B, H, W, Ci;
  const void* weight; /* bf16 [T][Co][w_ld], T = taps (ky-major) or B*taps when w_batched */
  long long w_ld;
  int Co, KH, KW, stride, pad, dil;
  int w_batched;
  long long w_batch_stride; /* elements between per-sample weight sets; 0 = taps * Co * w_ld (packed) */
  const int* origin; /* optional int32 [B][3] = (image, y0, x0): sample b reads image `image` shifted by (y0, x0);
                        H, W then describe the per-sample window used to size the output */
  int n_images;      /* number of images in `in` when origin != NULL (else ignored) */
  int img_h, img_w;  /* full image size when origin != NULL */
  const float* bias;     /* fp32 [Co] or NULL */
  const float* rowscale; /* fp32 [B*OH*OW] or NULL */
  float alpha;
  const float* scale_ptr; /* optional DEVICE scalar g (e.g. TransformerResFusionBlock.alpha R:343,353) */
  int act;                /* 0 none, 1 ReLU, 2 exact (erf) GELU */
  const float* res1; /* fp32, same addressing as the output, or NULL */
  long long res1_ld;
  float res1_scale;
  const void* res2; /* fp32 (or bf16 when res2_bf16) or NULL */
  long long res2_ld;
  int res2_bf16;
  float* out_f32; /* either or both outputs */
  long long out_f32_ld;
  void* out_bf16;
  long long out_bf16_ld;
  int store_mode; /* 0 plain; 1 PixelUnshuffle(2) R:377; 2 PixelShuffle(2) R:388 */
  int impl;       /* 0 = tcgen05 (product path); 1 = SIMT restatement (tests/debug only) */
  /* Optional fused LayerNorm of the OUTPUT rows: ln_out = LN(out) in bf16, i.e. the norm1 / norm2 that follows on the
   * residual stream (TransformerBlock R:318-331) folded into the conv that produces it, saving the norm kernel's read
   * of the fp32 stream.  ln_mode as tdr_rownorm (0 off, 1 WithBias R:189-205, 2 BiasFree R:172-186).  Supported with
   * impl 0, store_mode 0, fp32 output only, an fp32 res2, no res1, Co <= 96, 16 B-aligned ln_out rows; anything else
   * is rejected with TDR_EINVAL (tdr_conv_gemm_ln_supported tells beforehand). */
  int ln_mode;
  float ln_eps;
  const float* ln_weight; /* fp32 [Co] */
  const float* ln_bias;   /* fp32 [Co] (WithBias) or NULL */
  void* ln_out_bf16;
  long long ln_out_ld;
  /* 16-bit operand format of the MASA feature encoder R:100-134 (its features feed two arg-max searches, so they need
   * more mantissa than bf16 carries -- tools/precision_study.py): in_fp16 != 0: `in` AND `weight` hold IEEE fp16
   * (tcgen05 kind::f16 takes either format at the same rate); out_fp16 != 0: the 16-bit outputs (`out_bf16`,
   * `ln_out_bf16`) are written as fp16, saturating at +-65504 (not together with a bf16 res2).
   * The INFERENCE forward of the restoration nets runs every GEMM on fp16 operands too: bf16's 8 significand bits put
   * the 512x512 guided-Restormer output at mean |delta| 1.05e-3 / 54.3 dB / 0.032 dB PSNR-delta from the fp32
   * reference (the reference's own bf16 path: 1.7e-3 / 53.8 dB), fp16's 11 bits meet the 1e-3 / 0.01 dB bar
   * (tools/operand_format_study.py).  Accumulation, residual stream, LayerNorm, softmax stay fp32; the training
   * forward and every gradient stay bf16 (range). */
  int in_fp16;
  int out_fp16;
} tdr_conv_gemm_desc;
int tdr_conv_gemm(const tdr_conv_gemm_desc* d, cudaStream_t stream);
/* 1 when the fused output LayerNorm of the descriptor (see ln_mode) can run, 0 otherwise; no launch. */
int tdr_conv_gemm_ln_supported(const tdr_conv_gemm_desc* d);
/* ABI self-description for binding checks: writes sizeof(desc) and the offsets of weight, origin, bias, scale_ptr, res1,
 * res2, out_f32, out_bf16, impl, ln_mode, ln_weight, ln_out_bf16 into out[0..12]. */
void tdr_conv_gemm_desc_layout(int* out);

/* Small-channel direct 3x3 convs (SIMT): Ci <= 8 inputs (patch_embed R:362, masa_enc.conv_L1 R:106) read fp32 NHWC;
 * or Co <= 4 outputs (output conv R:640, + input image residual R:962) read bf16 NHWC.
 * small_ci `relu`: bit 0 = ReLU, bit 1 = the 16-bit output is IEEE fp16 instead of bf16. */
int tdr_conv3x3_small_ci(const float* in, int B, int H, int W, int Ci, const float* weight /* [Co][Ci][3][3] */,
                         const float* bias, int Co, int relu, float* out_f32, long long out_f32_ld, void* out_bf16,
                         long long out_bf16_ld, cudaStream_t stream);
int tdr_conv3x3_small_co(const void* in_bf16, long long in_ld, int B, int H, int W, int Ci,
                         const float* weight /* [Co][Ci][3][3] */, const float* bias, int Co,
                         const float* res /* fp32 NHWC [.., Co] or NULL */, float* out /* fp32 NHWC [.., Co] */,
                         cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Row-wise channel LayerNorm / cast: fp32 [rows, C] -> bf16 [rows, C].
 *   mode 0: cast only.  mode 1: WithBias LN R:189-205 (also LayerNorm2d nafnet_arch_utils.py:264-300 with eps 1e-6).
 *   mode 2: BiasFree LN R:172-186 (x / sqrt(var + eps) * w, mean not subtracted).
 * ------------------------------------------------------------------------------------------------------------- */
/* act: 0 none, 3 LeakyReLU(0.01) applied after the affine (mapper MLPs, main_train_tr_mapping.py:52-60); | 16: the
 * 16-bit output is IEEE fp16 instead of bf16.
 * Either or both of out_bf16 / out_f32 (e.g. the final DINO norm, models/dino/vision_transformers.py:262). */
int tdr_rownorm(const float* in, long long in_ld, long long rows, int C, int mode, const float* weight,
                const float* bias, float eps, int act, void* out_bf16, long long out_ld, float* out_f32,
                long long out_f32_ld, cudaStream_t stream);

/* fp32 [rows, C] -> bf16 and / or IEEE fp16 (saturating) copies in one pass (either may be NULL): the MASA encoder's
 * fp32 residual stream R:38-49 feeds the next conv as fp16 and the transfer kernels as bf16.
 *   out_fp16 = fp16(x * *scale16)   and, when rescale_in != 0, the fp32 rows are rewritten as x * *scale16 too;
 *   out_bf16 = bf16(x * *scale_bf16)            (scale pointers: DEVICE scalars or NULL = 1).
 * Level scaling of the MASA feature encoder.  The encoder is 17-21 ReLU convs with residual adds and no normalisation,
 * so its activations can grow geometrically with depth and leave fp16's range (65504).  Conv, bias, ReLU and the
 * residual add are positively homogeneous, so every level runs in units of a power-of-two scale s_L chosen on the
 * device from the level's first activation (no host sync, bit-exact: scaling by 2^k does not change mantissas):
 *   tdr_masa_level_scale: state_cur = {s_L, 1/s_L, 1/r, max} with r = 2^(ceil(log2 max|x|) - 6) if max|x| > 64 else 1,
 *                         s_L = s_prev * r  (x = the level's first activation in units of s_prev; state_cur[3] must be
 *                         zero on entry; state_prev NULL = {1, 1, 1});
 *   tdr_scale_vec       : out[i] = v[i] * *scale  (the level's conv biases in scaled units). */
int tdr_cast_rows(float* in, long long in_ld, long long rows, int C, void* out_bf16, long long out_bf16_ld,
                  void* out_fp16, long long out_fp16_ld, const float* scale16, const float* scale_bf16, int rescale_in,
                  cudaStream_t stream);
int tdr_masa_level_scale(const float* x, long long ld, long long rows, int C, const float* state_prev, float* state_cur,
                         cudaStream_t stream);
int tdr_scale_vec(const float* v, long long n, const float* scale, float* out, cudaStream_t stream);

/* IEEE fp16 rows -> bf16 rows (the backward's weight-gradient GEMMs take bf16 operands only: tcgen05 kind::f16 rejects
 * mixed fp16 x bf16 operand pairs, and gradients need bf16's exponent range). */
int tdr_cvt_f16_bf16(const void* in_fp16, long long in_ld, long long rows, int C, void* out_bf16, long long out_ld,
                     cudaStream_t stream);

/* Depthwise 3x3, pad 1 (R:231, R:253; N:conv2) on bf16 NHWC.  weight fp32 [9][C] (tap-major), bias fp32 [C] or NULL.
 * gate = 0: out[.., C].  gate = 1 (GDFN R:238-239): C = 2*Ch, out[.., Ch] = gelu(dw(x)[:Ch]) * dw(x)[Ch:].
 * gate = 2 (SimpleGate N:170-175): out = dw(x)[:Ch] * dw(x)[Ch:].
 * gate | 16: input and output are IEEE fp16 instead of bf16 (same kernels, other 16-bit conversions). */
int tdr_dwconv3x3(const void* in_bf16, long long in_ld, int B, int H, int W, int C, const float* weight,
                  const float* bias, int gate, void* out_bf16, long long out_ld, cudaStream_t stream);

/* Fused tail of the gated-dconv feed-forward R:236-240 (+ the residual add R:329 / the Res-fusion epilogue R:345-353):
 *   out = g*alpha * ( W_out . ( gelu(dw3x3(hidden)[:hp]) * dw3x3(hidden)[hp:] ) + bias ) + g*res1_scale*res1 + res2
 * in ONE kernel: the gated tensor never leaves the SM (depthwise stencil + exact GELU gate on CUDA cores -> SWIZZLE_128B
 * shared-memory tile -> tcgen05.mma against TMA-loaded slices of W_out, accumulator in TMEM), which removes its 4*hp B
 * per pixel of HBM traffic and one launch per transformer block.  hidden: 16-bit NHWC [B,H,W,hidden_ld >= 2*hp]
 * (bf16, or IEEE fp16 when fp16 != 0 -- then w_out is fp16 too); dw_weight fp32 [9][2*hp] tap-major, dw_bias fp32
 * [2*hp] or NULL; w_out 16-bit [C][w_ld >= hp]; out / res1 / res2 fp32 rows (out may alias res2).
 * Supported: 16 <= C <= 192, C % 16 == 0, hp % 8 == 0 (tdr_gdfn_tail_supported tells; callers keep the two-kernel
 * path otherwise). */
typedef struct tdr_gdfn_tail_desc {
  const void* hidden;
  long long hidden_ld;
  int B, H, W, hp, C;
  const float* dw_weight;
  const float* dw_bias;
  const void* w_out;
  long long w_ld;
  const float* bias;
  float alpha;
  const float* scale_ptr;
  const float* res1;
  long long res1_ld;
  float res1_scale;
  const float* res2;
  long long res2_ld;
  float* out;
  long long out_ld;
  int fp16;
} tdr_gdfn_tail_desc;
int tdr_gdfn_tail(const tdr_gdfn_tail_desc* d, cudaStream_t stream);
int tdr_gdfn_tail_supported(const tdr_gdfn_tail_desc* d);

/* NAFNet pieces (network_nafnet_guided_arch.py, "N:").
 *   tdr_gate_mul     : SimpleGate N:170-175 on bf16 rows, out[r, c] = x[r, c] * x[r, C + c].
 *   tdr_naf_sca_fold : Simplified channel attention N:192-196 folded into conv3 N:229: s = W_sca * avgpool(g) + b_sca,
 *                      Weff[b][co][ci] = rowscale[co] * W3[co][ci] * s[b][ci]  (bf16 [B][Co][weff_ld]); then
 *                      `x * sca(x)` -> conv3 -> `* beta` is one tdr_conv_gemm(g, Weff, w_batched=1). */
int tdr_gate_mul(const void* x_bf16, long long ld, long long rows, int C /* output channels */, void* out_bf16,
                 long long out_ld, int fp16 /* in / out are IEEE fp16 */, cudaStream_t stream);
size_t tdr_naf_sca_workspace_bytes(int B, long long P, int C);
int tdr_naf_sca_fold(const void* g_bf16, long long ld, int B, long long P, int C, const float* w_sca /* [C][C] */,
                     const float* b_sca, const float* w3 /* fp32 [Co][C] */, int Co, const float* rowscale /* [Co] */,
                     void* weff_bf16, long long weff_ld, float* workspace,
                     float* mean_out /* optional [B][C]: avgpool(g) */, float* s_out /* optional [B][C]: sca vector */,
                     void* weff_t_bf16 /* optional Weff[b]^T [B][C][weff_t_ld] (dgrad operand, always bf16) */,
                     long long weff_t_ld, int fp16 /* g and Weff are IEEE fp16 */, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MDTA channel attention R:262-276.
 *   tdr_mdta_gram : per (sample, head) Gram q^T k plus sum-of-squares of q and k over all pixels (tcgen05, split over
 *                   pixel chunks into `partials`).  qkv = bf16 [B, P, ld] with q|k|v at channel offsets 0|C|2C.
 *   tdr_mdta_weff : reduce partials, attn = softmax(G / (|q||k|) * temperature) R:266-270, and fold it into
 *                   project_out: Weff[b] = W_out * blockdiag(attn)  -> bf16 [B][C][weff_ld]; then `attn @ v` +
 *                   project_out is tdr_conv_gemm(v, Weff, w_batched=1).  attn_ws: fp32 [B,heads,c,c] workspace that
 *                   receives the attention maps (two launches: softmax, fold).
 * ------------------------------------------------------------------------------------------------------------- */
size_t tdr_mdta_partials_bytes(int B, long long P, int C, int heads);
int tdr_mdta_gram(const void* qkv_bf16, long long ld, int B, long long P, int C, int heads, float* partials,
                  int fp16 /* qkv holds IEEE fp16 */, cudaStream_t stream);
int tdr_mdta_weff(const float* partials, int B, long long P, int C, int heads, const float* temperature /* [heads] */,
                  const float* w_out /* fp32 [C][C] */, void* weff_bf16, long long weff_ld, float* attn_ws,
                  void* weff_t_bf16 /* optional: Weff[b]^T, same ld (the dgrad operand of the training step) */,
                  float* shat_out /* optional fp32 [B][heads][c*c + 2c]: normalised Gram | |q| | |k| (for tdr_mdta_bwd) */,
                  int fp16 /* Weff is written as IEEE fp16 (weff_t stays bf16) */,
                  const float* topk_w /* optional fp32 [4]: top-k sparse attention of the DRSformer family
                                         (network_drsformer_guided_arch.py:296-327) -- attn = sum_i topk_w[i] * softmax over the
                                         int(c/2), int(2c/3), int(3c/4), int(4c/5) largest entries of each row */,
                  cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * ViT encoder / mapper glue (DINOv2 ViT-B/14 models/dino/*.py "D:", CLIP ViT-H/14 via transformers, mappers "M:" =
 * scripts/train/main_train_tr_mapping.py:40-122).  Dense contractions go through tdr_conv_gemm (H == 1 view).
 * ------------------------------------------------------------------------------------------------------------- */
/* NCHW fp32 images -> bf16 patches [B, (H/p)*(W/p), ld], K index = (c*p + ky)*p + kx  (D: patch_embed.py:76) */
int tdr_vit_patchify(const float* img, int B, int C, int H, int W, int patch, void* out_bf16, long long ld,
                     cudaStream_t stream);
/* x[b,0] = cls + pos[0]; x[b,1+t] = patch_tokens[b,t] + pos[1+t]  (D: vision_transformers.py:215-216) */
int tdr_vit_assemble_tokens(const float* patch_tokens, const float* cls, const float* pos, int B, int N, int D, float* x,
                            cudaStream_t stream);
/* p = softmax(scale * s[row, 0:n]) as bf16, pad columns zeroed (D: attention.py:61-64) */
int tdr_softmax_rows(const float* s, long long ld, long long rows, int n, float scale, void* out_bf16, long long ld_out,
                     cudaStream_t stream);
/* vt[b, h, d, t] = qkv[b, t, voff + h*hd + d], t padded with zeros to n_pad (K-major operand of p.v) */
int tdr_vit_transpose_v(const void* qkv_bf16, long long ld, int B, int N, int heads, int hd, int voff, void* vt_bf16,
                        int n_pad, cudaStream_t stream);
/* Fused multi-head self-attention of the frozen ViTs (models/dino/attention.py:52-68; transformers CLIPAttention as called
 * from scripts/train/main_train_tr_mapping.py:780): out[b, i, h*hd:(h+1)*hd] = softmax_j(scale q_i.k_j) v_j, one launch per
 * layer.  qkv: bf16 rows [B, N, ld] holding q | k | v at columns 0 | D | 2D (D = heads * hd); out: bf16 rows [B, N, out_ld].
 * fp32 scores / online softmax / accumulation (tcgen05 + TMEM), nothing of size N x N touches HBM.  hd in {16, 32, 64, 80}. */
int tdr_vit_attention_supported(int hd);
int tdr_vit_attention(const void* qkv_bf16, long long ld, int B, int N, int heads, int hd, float scale, void* out_bf16,
                      long long out_ld, cudaStream_t stream);
/* Generic grouped / depthwise K x K stencil over NHWC 16-bit rows (DRSformer MSFN :216-256 and MEFC experts :454-520):
 * out16[b,y,x,co] = act(bias[co] + sum_{j<ipg} sum_taps weight[co][j][ky][kx] * in16[b, y+(ky-K/2)dil, x+(kx-K/2)dil, idx[co*ipg+j]])
 * with zero padding; ipg in {1, 2}, K in {1, 3, 5, 7}; act 1 = ReLU; pool != 0: AvgPool2d(3, 1, 1, count_include_pad=False).
 * The index table expresses depthwise (idx[co] = co), grouped 2 -> 1 convs and the reference's chunk / cat re-orderings. */
int tdr_grouped_stencil(const void* in16, long long in_ld, int B, int H, int W, int Co, int ipg, const int* idx,
                        const float* weight, const float* bias, int K, int dil, int act, int pool, void* out16,
                        long long out_ld, int fp16, cudaStream_t stream);
/* MEFC (network_drsformer_guided_arch.py:371-549): the gate softmax_ops(Linear(ReLU(Linear(avgpool(x))))) of OALayer /
 * subnet.forward as [B, O = steps * num_ops] fp32, and the per-sample 1x1 weights of OperationLayer._out with the gate of
 * a step folded into their column blocks (out16[b][co][k*C + ci] = weight[co][k*C + ci] * gate[b][k]). */
int tdr_mefc_gate(const float* emb, long long emb_ld, int B, int C, const float* w1, const float* b1, int H1, const float* w2,
                  const float* b2, int O, int num_ops, float* out, cudaStream_t stream);
int tdr_mefc_mix_weights(const float* weight, int Co, int Ci, int C, const float* gate, long long gate_ld, int B, void* out16,
                         long long ld, int fp16, cudaStream_t stream);
/* PromptGenBlock.forward (models/archs/network_promptir_guided_arch.py:424-440) between the spatial mean and the 3x3 conv:
 * out[b, k] = softmax_k(linear_layer(emb[b]))  (L <= 8 prompts), and
 * out16[b, y, x, d] = bilinear(sum_k wts[b, k] * prompt_param[k, d])(y, x), F.interpolate(mode="bilinear") semantics,
 * written NHWC in the 16-bit operand format of the conv that follows (fp16 != 0: IEEE fp16, else bf16). */
int tdr_prompt_weights(const float* emb, long long emb_ld, const float* weight, const float* bias, int B, int C, int L,
                       float* out, cudaStream_t stream);
int tdr_prompt_mix_resize(const float* prompt, int L, int D, int S, const float* wts, int B, int H, int W, void* out16,
                          long long out_ld, int fp16, cudaStream_t stream);
/* Reference-crop selection (models/image_restoration_ref_model.py:215-247): crop (origin[k] = (image, y0, x0), size
 * crop_h x crop_w) + bilinear resize (align_corners=False) of NCHW fp32 images in one pass, and the cosine similarity
 * between one query feature row per sample and n candidate rows. */
int tdr_crop_resize(const float* img, int C, int H, int W, const int* origin, int ncrops, int crop_h, int crop_w,
                    int out_h, int out_w, float* out, cudaStream_t stream);
int tdr_cosine_rows(const float* fl /* [B, F] */, const float* fr /* [B*n, F] */, int B, int n, long long F,
                    float* cosv /* [B, n] */, cudaStream_t stream);
/* out[b, c] (+)= mean_t x[b, t0 + t, c], t < n  (M: `.mean(dim=1, keepdim=True)` :77) */
int tdr_mean_tokens(const float* x, long long ld, int B, int tokens_per_b, int t0, int n, int C, float* out,
                    long long out_ld, int accumulate, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Optimizer tail of the DDP step on flat fp32 buffers (models/image_restoration_ref_model.py:276-284,
 * models/base_model.py:54-62).  No host synchronisation: the clip coefficient stays on the device.
 * ------------------------------------------------------------------------------------------------------------- */
int tdr_sumsq_partial_count(void); /* number of floats tdr_sumsq_partial writes */
int tdr_sumsq_partial(const float* g, long long n, float* partial, cudaStream_t stream);
/* out2[0] = min(1, max_norm / (grad_scale * sqrt(sum partial) + 1e-6)), out2[1] = total norm (clip_grad_norm_) */
int tdr_clip_coef(const float* partial, int n, float max_norm, float grad_scale, float* out2, cudaStream_t stream);
/* torch.optim.AdamW semantics on grad_scale * clip_coef[0] * g (clip_coef may be NULL); step >= 1 */
int tdr_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                   float eps, float weight_decay, int step, float grad_scale, const float* clip_coef,
                   cudaStream_t stream);
int tdr_ema_update(float* ema, const float* p, long long n, float decay, cudaStream_t stream);
/* L1 pixel loss (reference losses L1Loss, reduction 'mean', image_restoration_ref_model.py:268-274) and its gradient in
 * one pass: loss[0] = w * mean|x - gt| (device scalar), dx = w / n * sign(x - gt).  partial: scratch of
 * tdr_sumsq_partial_count() floats. */
int tdr_l1_loss_grad(const float* x, const float* gt, long long n, float loss_weight, float* dx, float* loss,
                     float* partial, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Backward kernels (training step).  The reference obtains every gradient from autograd over stock PyTorch ops
 * (`l_total.backward()` models/image_restoration_ref_model.py:268-275); the data gradients of convolutions are
 * tdr_conv_gemm / tdr_dwconv3x3 calls with transposed + flipped weights, everything else is below.
 * Reductions are two-stage and deterministic; `accumulate` != 0 adds into the destination (gradient accumulation).
 * ------------------------------------------------------------------------------------------------------------- */
/* Weight gradient of a dense convolution (tcgen05, contraction over pixels):
 *   dW[co][ci][ky][kx] = sum_{b,y,x} dy[b,y,x,co] * x[b, y*stride + ky*dil - pad, x*stride + kx*dil - pad, ci]
 * out[b*out_stride_b + map(co)*out_stride_co + map(ci)*out_stride_ci + tap*out_stride_tap] (+)= scale * dW; with
 * per_sample != 0 one result per batch sample (the per-sample Weff of MDTA R:272-276), else summed over the batch.
 * co_map / ci_map: optional DEVICE int32 maps from the (padded) kernel channel to the parameter channel, -1 = skip. */
typedef struct tdr_wgrad_desc {
  const void* dy; /* bf16 [B, OH, OW, dy_ld], channels [0, Co) */
  long long dy_ld;
  const void* x; /* bf16 [B, H, W, x_ld], channels [0, Ci) */
  long long x_ld;
  int B, H, W, Ci, Co, KH, KW, stride, pad, dil;
  int per_sample;
  float* out;
  long long out_stride_b, out_stride_co, out_stride_ci, out_stride_tap;
  const int* co_map;
  const int* ci_map;
  int accumulate;
  float scale;
  float* workspace; /* tdr_wgrad_workspace_bytes(d) */
  size_t workspace_bytes;
  const float* scale_ptr; /* optional DEVICE scalar multiplied into `scale` (level scale of the MASA encoder's tape) */
} tdr_wgrad_desc;
size_t tdr_wgrad_workspace_bytes(const tdr_wgrad_desc* d);
int tdr_wgrad(const tdr_wgrad_desc* d, cudaStream_t stream);
/* workspace (bytes) large enough for tdr_colsum / tdr_dwconv3x3_wgrad / tdr_rownorm_bwd / tdr_dot_f32 on C channels */
size_t tdr_reduce_workspace_bytes(int C);
/* out[map(c) * out_stride] (+)= sum_rows x[r, c]   (bias gradients) */
int tdr_colsum(const void* x_bf16, long long ld, long long rows, int C, float* out, long long out_stride,
               const int* c_map, int accumulate, float* workspace, cudaStream_t stream);
/* depthwise 3x3 (pad 1) weight and bias gradient: dw fp32 [C_param][9] (= nn.Conv2d weight [C,1,3,3]), db [C_param] or NULL */
int tdr_dwconv3x3_wgrad(const void* dy_bf16, long long dy_ld, const void* x_bf16, long long x_ld, int B, int H, int W,
                        int C, float* dw, float* db, const int* c_map, int accumulate, float* workspace,
                        cudaStream_t stream);
/* Backward of tdr_rownorm (same modes): dx[r,:] = (add ? add[r,:] : 0) + d/dx( norm(x)[r,:] . dy[r,:] ), fp32;
 * dweight / dbias (+)= column sums (either may be NULL; workspace needed when dweight != NULL).  dx may alias add. */
int tdr_rownorm_bwd(const float* x, long long x_ld, const void* dy_bf16, long long dy_ld, long long rows, int C, int mode,
                    const float* weight, float eps, const float* add, long long add_ld, float* dx, long long dx_ld,
                    void* dx_bf16 /* optional bf16 copy of dx (operand of the next dgrad / wgrad) */, long long dx_bf16_ld,
                    float* dweight, float* dbias, int accumulate, float* workspace, cudaStream_t stream);
/* Gate backward.  y = pre-gate depthwise output [rows, 2*Ch] (a | b), dg = gradient of the gated product [rows, Ch]:
 * gate 1 (GDFN, gelu(a)*b R:238-239): dy = [dg*b*gelu'(a) | dg*gelu(a)];  gate 2 (SimpleGate): dy = [dg*b | dg*a].
 * dy may alias y. */
int tdr_gate_bwd(const void* y_bf16, long long y_ld, const void* dg_bf16, long long dg_ld, long long rows, int Ch,
                 int gate, void* dy_bf16, long long dy_ld,
                 const float* dg_add /* optional fp32 [rows / rows_per_sample][Ch] added to dg (SCA pool gradient) */,
                 long long rows_per_sample, cudaStream_t stream);
/* Training forward of the gated depthwise conv: as tdr_dwconv3x3(gate 1|2) and ALSO stores the pre-gate tensor
 * y = [a | b] (bf16 [B,H,W,>=C]) that tdr_gate_bwd needs, so the backward pass does not recompute the convolution. */
int tdr_dwconv3x3_gated_train(const void* in_bf16, long long in_ld, int B, int H, int W, int C, const float* weight,
                              const float* bias, int gate, void* out_bf16, long long out_ld, void* y_bf16, long long y_ld,
                              cudaStream_t stream);
/* Fused recompute + gate backward: dy = tdr_gate_bwd(tdr_dwconv3x3(in, gate 0), dg) without writing the pre-gate tensor:
 * in = the depthwise conv's INPUT [B,H,W,C] (C = 2*Ch), dy bf16 [B,H,W,>=C] receives [d a | d b]. */
int tdr_dwconv3x3_gate_bwd(const void* in_bf16, long long in_ld, int B, int H, int W, int C, const float* weight,
                           const float* bias, int gate /* 1 GELU gate, 2 SimpleGate */, const void* dg_bf16,
                           long long dg_ld, const float* dg_add /* optional [B][C/2] */, void* dy_bf16, long long dy_ld,
                           cudaStream_t stream);
/* NAFBlock scaled convs N:225-237, y = x + scale[co] * (W (g * s_b) + bias) with scale = beta (conv3, s_b = SCA vector)
 * or gamma (conv5, s = NULL).  raw = tdr_wgrad(dy, g) [nb][Co][C] (per sample when s != NULL), colsum_dy = tdr_colsum(dy):
 *   dW += scale * sum_b s_b raw_b,  dscale += sum W s_b raw_b + bias * colsum_dy,  dbias += scale * colsum_dy.
 * tdr_naf_sca_bwd continues through the SCA branch N:192-196: dW_sca, db_sca (+=) and dg_add[b][c] = the per-sample
 * constant that the average pool adds to every pixel's dg (pass it to tdr_gate_bwd). */
int tdr_naf_scaled_conv_bwd(const float* raw, int nb, int Co, int C, const float* W /* fp32 [Co][C] */, const float* bias,
                            const float* scale, const float* colsum_dy, const float* s /* [nb][C] or NULL */, float* dW,
                            float* dbias, float* dscale, cudaStream_t stream);
int tdr_naf_sca_bwd(const float* raw, int B, int Co, int C, const float* w3, const float* scale, const float* mean,
                    const float* w_sca, long long P, float* dw_sca, float* db_sca, float* dg_add,
                    float* workspace /* B*C floats */, cudaStream_t stream);
/* Backward of the MDTA score path R:266-276 given dWeff (per-sample tdr_wgrad of the attn.v.project_out product):
 * dW_out (+)=, dtemperature (+)=, and mqk[b] = the [2C x 2C] matrix with [dq; dk] = mqk[b] . [q; k] per pixel (softmax,
 * temperature and F.normalize backward folded; bf16 [B][2C][mqk_ld], rows beyond 2C / pad columns untouched). */
size_t tdr_mdta_bwd_workspace_bytes(int B, int C, int heads);
int tdr_mdta_bwd(const float* shat /* tdr_mdta_weff shat_out */, const float* attn, int B, long long P, int C, int heads,
                 const float* temperature, const float* w_out, const float* dweff /* fp32 [B][C][C] */, void* mqk_bf16,
                 long long mqk_ld, float* dw_out, float* dtemperature, int accumulate, float* workspace,
                 cudaStream_t stream);
/* out = (scale_ptr ? *scale_ptr : 1) * scale * x + y   (fp32 rows; y may be NULL; R:353 and its backward) */
int tdr_scale_add_f32(const float* x, long long x_ld, const float* y, long long y_ld, long long rows, int C,
                      const float* scale_ptr, float scale, float* out, long long out_ld, cudaStream_t stream);
/* out[0] (+)= sum x*y over [rows, C] fp32 (gradient of the fusion gate alpha R:343) */
int tdr_dot_f32(const float* x, long long x_ld, const float* y, long long y_ld, long long rows, int C, float* out,
                int accumulate, float* workspace, cudaStream_t stream);
/* bf16 NHWC PixelUnshuffle(2) (mode 1) / PixelShuffle(2) (mode 2): the adjoints of store_mode 2 / 1 of tdr_conv_gemm */
int tdr_pixel_shuffle_nhwc(const void* in_bf16, long long in_ld, int B, int H, int W, int C, int mode, void* out_bf16,
                           long long out_ld, cudaStream_t stream);
/* out = y > 0 ? dy : 0 (ReLU backward from the stored activation, R:38-49,106-116).  y may be bf16 or IEEE fp16: only its
 * sign / zero bits are inspected, which the two formats share. */
int tdr_relu_mask(const void* y_bf16, long long y_ld, const void* dy_bf16, long long dy_ld, long long rows, int C,
                  void* out_bf16, long long out_ld, cudaStream_t stream);

/* Backward of the MASA transfer / confidence path (the matches themselves are arg-max constants).
 *   tdr_masa_transfer_bwd: dref[gathered ref pixel] += dout * a / cnt  and  datt[win, q] += bilinear^T(sum_c dout * mean)
 *                          (dref fp32 [B, Hr_s, Wr_s, C] and datt fp32 [nwin, k_y*k_x] are accumulated with fp32 atomics
 *                          and must be zero-initialised by the caller; the only non-deterministic reductions here).
 *   tdr_masa_fine_bwd    : gradient of att = max cosine (R:661-670) w.r.t. the deepest lq and ref features.
 *   tdr_dilate2_nhwc     : zero insertion out[b,2y,2x,:] = in[b,y,x,:] (adjoint of stride 2; out pre-zeroed), so the
 *                          data gradient of the stride-2 convs R:109-116 is a stride-1 tdr_conv_gemm. */
int tdr_masa_transfer_bwd(const float* dout, long long dout_ld, const void* f_ref_bf16, int B, int Hr_s, int Wr_s, int C,
                          const int* origin, const int* index, const float* att, int py, int px, int k_y, int k_x,
                          int d_x, int s, float* dref, float* datt, cudaStream_t stream);
int tdr_masa_fine_bwd(const void* f_lq_bf16, int B, int H, int W, const void* f_ref_bf16, int Hr, int Wr, int C, int k_y,
                      int k_x, int d_x, const int* origin, const int* index, const float* datt, float* dlq, float* dref,
                      cudaStream_t stream);
int tdr_dilate2_nhwc(const void* in_bf16, long long in_ld, int B, int H, int W, int C, void* out_bf16, long long out_ld,
                     int OH, int OW, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Weight packing (host-side `prepared()` of the arch modules; re-run after every optimizer step in training).
 *   tdr_pack_conv_weight: nn.Conv2d weight fp32 [Co][Ci][KH][KW] -> bf16 [T][Co_p][ld] (the tdr_conv_gemm operand) and/or
 *                         its transposed, tap-flipped twin [T][Ci_p][ld_t] (the data-gradient operand).  co_map / ci_map:
 *                         optional DEVICE int32 padded->logical channel maps (-1 = zero row / column); scale: optional
 *                         fp32 [Co] multiplied into output channel co.
 *   tdr_pack_dw_weight  : depthwise weight [C][1][3][3] -> fp32 [9][C_p], its flipped twin, and the padded bias.
 *   tdr_gather_vec      : out[i] = map[i] >= 0 ? v[map[i]] : 0.
 * ------------------------------------------------------------------------------------------------------------- */
int tdr_pack_conv_weight(const float* w, int Co, int Ci, int KH, int KW, const int* co_map, int Co_p, const int* ci_map,
                         int Ci_p, const float* scale, void* out_bf16, long long ld, void* out_t_bf16, long long ld_t,
                         int fwd_fp16 /* the forward operand `out_bf16` is IEEE fp16; the twin stays bf16 */,
                         cudaStream_t stream);
int tdr_pack_dw_weight(const float* w, const float* bias, int C, const int* c_map, int C_p, float* out, float* out_flip,
                       float* out_bias, cudaStream_t stream);
int tdr_gather_vec(const float* v, const int* map, int n, int n_src, float* out, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Layout / copies.
 * ------------------------------------------------------------------------------------------------------------- */
int tdr_nchw_to_nhwc(const float* src, int B, int C, int H, int W, int pad_h, int pad_w /* zero-padded output size */,
                     float* dst_f32, long long dst_f32_ld, void* dst_bf16, long long dst_bf16_ld, cudaStream_t stream);
/* dst[b,c,y,x] = src[b,y,x,c] (+ res[b,y,x,c], e.g. the `+ inp_img` of R:499,962), cropped to out_h x out_w. */
/* NCHW fp32 image -> dense 16-bit NHWC rows [B, pad_h, pad_w, C16], zero beyond (C, H, W): the 16-byte-row operand that lets
 * the image-boundary 3x3 convs (OverlapPatchEmbed :362-370, Encoder.conv_L1 :106) run on tdr_conv_gemm. */
int tdr_image_to_rows16(const float* src, int B, int C, int H, int W, int pad_h, int pad_w, int C16, void* dst16, int fp16,
                        cudaStream_t stream);
int tdr_nhwc_to_nchw(const float* src, long long src_ld, int B, int C, int H, int W /* source size */, int out_h,
                     int out_w /* crop */, const float* res /* NHWC fp32 or NULL */, long long res_ld, float* dst,
                     cudaStream_t stream);
/* dst[r, 0:C] = src[r, 0:C] (fp32 rows with independent strides); optionally also writes a 16-bit copy (bf16, or IEEE
 * fp16 when dst16_fp16). */
int tdr_copy_rows_f32(const float* src, long long src_ld, long long rows, int C, float* dst, long long dst_ld,
                      void* dst_bf16, long long dst_bf16_ld, int dst16_fp16, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * MASA match-and-transfer R:753-900 (closed form of SURVEY.md appendix A; no unfold/fold materialisation).
 * `k` = lr block size in feature pixels, (py, px) block grid, d = window diameter.
 * The two arg-max searches see the deepest features in fp32: bf16 descriptors alone flip 2.5 % of the fine matches
 * between near-tied candidates (tools/precision_study.py).  The correlations still run as ONE bf16 tcgen05 GEMM over
 * 3C channels: the reference features are split into [hi | lo | hi] bf16 thirds (tdr_masa_split3), the normalised lq
 * descriptors into [hi | hi | lo], so the fp32 accumulator receives hi*hi + lo*hi + hi*lo (2^-16 relative).
 * The transfer / backward kernels read the bf16 copies of the features.
 * ------------------------------------------------------------------------------------------------------------- */
/* n2[r] = sum_c x[r,c]^2, x fp32 */
int tdr_sqnorm_rows(const float* x, long long ld, long long rows, int C, float* n2, cudaStream_t stream);
/* out[r] = [bf16(x) | bf16(x - bf16(x)) | bf16(x)]  (bf16 [rows][out_ld >= 3C]) */
int tdr_masa_split3(const float* x, long long ld, long long rows, int C, void* out_bf16, long long out_ld,
                    cudaStream_t stream);
/* inv[dil_i][b,y,x] = 1 / max(sqrt(sum_{3x3 taps, dilation dil_i, zero pad} n2), 1e-12)   (R:683,690) */
int tdr_masa_ref_invnorm(const float* n2, int B, int H, int W, const int* host_dils, int ndil, float* inv,
                         cudaStream_t stream);
/* coarse filters R:685-689: w[dil_i][b][tap][blk (padded to co_pad)][3C] = normalised dilated 3x3 centre descriptor of
 * lq block blk (replicate-padded halo R:785), f_lq fp32 [B,H,W,C] dense, each row split as [hi | hi | lo]. */
int tdr_masa_coarse_filters(const float* f_lq, int B, int H, int W, int C, int k_y, int k_x, const int* host_dils,
                            int ndil, int co_pad, void* w_bf16, cudaStream_t stream);
/* argmax over ref positions of score[b, pos, blk] (fp32, ld = co_pad) R:694 + window placement R:793-815.
 * origin[b*nblk + blk] = (b, y1, x1). */
int tdr_masa_coarse_argmax(const float* score, int B, int Hr, int Wr, int nblk, int co_pad, int d_y, int d_x,
                           int* idx_out, int* origin, cudaStream_t stream);
/* fine filters R:662-665: w[b*nblk+blk][tap][q (k_y*k_x)][3C] = normalised 3x3 patch at interior position q, split as
 * [hi | hi | lo]; f_lq fp32 dense. */
int tdr_masa_fine_filters(const float* f_lq, int B, int H, int W, int C, int k_y, int k_x, void* w_bf16,
                          cudaStream_t stream);
/* inv[blk, jy, jx] = 1 / max(|3x3 patch of the window at (jy, jx)|, 1e-12) from the per-pixel n2 map R:666 */
int tdr_masa_win_invnorm(const float* n2_ref, int Hr, int Wr, const int* origin, int nwin, int d_y, int d_x,
                         float* inv, cudaStream_t stream);
/* argmax over the d_y*d_x window positions of corr[win, pos, q] (fp32, ld = nq) R:670: index[win, q], att[win, q]. */
int tdr_masa_fine_argmax(const float* corr, int nwin, int npos, int nq, int* index, float* att, cudaStream_t stream);
/* transfer R:698-715 + re-tiling R:877-891 at scale s: out[b, Y, X, 0:C] (fp32, out_ld) from f_ref level
 * [B, Hr*s, Wr*s, C] bf16. */
int tdr_masa_transfer(const void* f_ref_bf16, int B, int Hr_s, int Wr_s, int C, const int* origin, const int* index,
                      const float* att, int py, int px, int k_y, int k_x, int d_x, int s, float* out, long long out_ld,
                      void* out_bf16, long long out_bf16_ld, cudaStream_t stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Input pipeline on the device: what Dataset_PairedImageWithRef.__getitem__ does to a decoded frame
 * (data/restoration_dataset.py:194-253) -- imfrombytes float32 / 255 (utils/utils_image.py:216-217), `padding` =
 * cv2.BORDER_REFLECT at the bottom / right up to the patch size (:243-254), paired_random_crop
 * (data/transforms.py:24-83), random_augmentation mode 0..7 (data/transforms.py:223-275), img2tensor BGR->RGB + HWC->CHW
 * (utils/utils_image.py:102-126), optional normalize (x - mean) / std (:240-244) -- for n samples in one launch.
 * The random decisions (top, left, mode) are the caller's (the reference's `random.randint` calls); results are
 * bit-identical to the reference's numpy / cv2 chain.  out = fp32 [n, channels, out_h, out_w].
 * ------------------------------------------------------------------------------------------------------------- */
typedef struct tdr_patch_desc {
  const void* image; /* DEVICE uint8 [h, w, channels], HWC, BGR when channels == 3 */
  int h, w;
  int top, left; /* crop origin in the (reflect-padded) frame */
  int mode;      /* data_augmentation mode 0..7; modes 2, 3, 6, 7 (rot90 family) need out_h == out_w */
  int reserved;
} tdr_patch_desc;
int tdr_prepare_patches(const tdr_patch_desc* descs_device, const tdr_patch_desc* descs_host /* same content, validated */,
                        int n, int channels, int out_h, int out_w, int bgr2rgb, const float* mean /* host [channels] or NULL */,
                        const float* stdv /* host [channels] or NULL */, float* out, cudaStream_t stream);
/* Training noise of the Gaussian-denoising datasets (data/restoration_dataset.py:474-476): out = img + fl(n * level[b]) with
 * level[b] = sigma_b / 255 per sample (device array; `sigma_type` constant / random / choice is the caller's draw).
 * noise != NULL: n = caller-supplied standard normals (bit-identical to the reference's mul_ / add_ for that draw);
 * noise == NULL: n from a Philox4x32-10 stream keyed by (seed, sample, element) + Box-Muller (no host RNG, no H2D copy). */
int tdr_add_gaussian_noise(const float* img, const float* noise, const float* level_device, int B, long long per_sample,
                           unsigned long long seed, float* out, cudaStream_t stream);

/* Validation PSNR (use_image: true) without a device->host image copy: per image the exact integer
 * sum (q1 - q2)^2 over the crop_border-trimmed window and max(q1), q = tensor2img's uint8 quantisation
 * (utils/utils_image.py:160-186: clamp [0, 1], * 255.0, round half-to-even).  The host finishes with calculate_psnr's
 * float64 formula (metrics/psnr_ssim.py:55-59) and obtains the identical double.  img1 / img2: fp32 [B, C, H, W]. */
int tdr_psnr_u8_sums(const float* img1, const float* img2, int B, int C, int H, int W, int crop_border,
                     unsigned long long* sse /* [B] */, int* max1 /* [B] */, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* TDR_SM100_H_ */
