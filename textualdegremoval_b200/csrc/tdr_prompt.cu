// PromptGenBlock of the reference's PromptIR-guided network
// (/root/reference/models/archs/network_promptir_guided_arch.py:417-440):
//     emb    = x.mean(dim=(-2, -1))                                   -> tdr_mean_tokens (existing)
//     w      = softmax(linear_layer(emb), dim=1)                      -> tdr_prompt_weights
//     prompt = sum_k w[b, k] * prompt_param[k]                        \
//     prompt = F.interpolate(prompt, (H, W), mode="bilinear")         /  tdr_prompt_mix_resize (interpolation is linear, so
//     prompt = conv3x3(prompt)                                            the two commute: one pass, NHWC 16-bit out)
// and the 3x3 conv is tdr_conv_gemm.  Byte work on tiny tensors (prompt_param is 5 x D x S x S fp32, <= 5 MB): SIMT.
#include "tdr_common.cuh"

namespace {

// grid B, 32 * L threads (L <= 8 prompts): warp k computes logit k = <W[k, :], emb[b, :]> + bias[k]; then softmax over k.
__global__ void prompt_weights_kernel(const float* __restrict__ emb, long long emb_ld, const float* __restrict__ w,
                                      const float* __restrict__ bias, int C, int L, float* __restrict__ out) {
  __shared__ float logit[8];
  const int b = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* e = emb + (size_t)b * emb_ld;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) s = fmaf(w[(size_t)k * C + c], e[c], s);
  s = warp_sum(s);
  if (lane == 0) logit[k] = s + (bias ? bias[k] : 0.f);
  __syncthreads();
  if (threadIdx.x < L) {
    float mx = -INFINITY;
    for (int j = 0; j < L; ++j) mx = fmaxf(mx, logit[j]);
    float sum = 0.f;
    for (int j = 0; j < L; ++j) sum += expf(logit[j] - mx);
    out[b * L + threadIdx.x] = expf(logit[threadIdx.x] - mx) / sum;
  }
}

// out[b, y, x, d] (NHWC, 16-bit) = bilinear_{align_corners=False}( sum_k wts[b, k] * prompt[k, d, :, :] )(y, x)
// One thread per (pixel, 8 channels): the 4 source taps of a pixel are shared by its channels.
__global__ void __launch_bounds__(256) prompt_mix_resize_kernel(const float* __restrict__ prompt, int L, int D, int S,
                                                                const float* __restrict__ wts, int B, int H, int W,
                                                                uint16_t* __restrict__ out, long long out_ld, int fp16) {
  const int dv = D >> 3;
  const long long total = (long long)B * H * W * dv;
  const float sy = (float)S / H, sx = (float)S / W;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d0 = (int)(i % dv) * 8;
    const long long pix = i / dv;
    const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
    float fy = (y + 0.5f) * sy - 0.5f, fx = (x + 0.5f) * sx - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int iy1 = iy0 + 1 < S ? iy0 + 1 : S - 1, ix1 = ix0 + 1 < S ? ix0 + 1 : S - 1;
    const float ly = fy - iy0, lx = fx - ix0;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
      for (int k = 0; k < L; ++k) {
        const float wk = wts[b * L + k];
        const float* p = prompt + ((size_t)k * D + d0 + e) * S * S;
        v00 = fmaf(wk, p[iy0 * S + ix0], v00);
        v01 = fmaf(wk, p[iy0 * S + ix1], v01);
        v10 = fmaf(wk, p[iy1 * S + ix0], v10);
        v11 = fmaf(wk, p[iy1 * S + ix1], v11);
      }
      v[e] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
    }
    pack8r(out + pix * out_ld + d0, v, fp16);
  }
}

}  // namespace

extern "C" int tdr_prompt_weights(const float* emb, long long emb_ld, const float* weight, const float* bias, int B, int C,
                                  int L, float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(emb && weight && out && B > 0 && C > 0 && L >= 1 && L <= 8 && emb_ld >= C,
                "tdr_prompt_weights: bad arguments (1 <= prompt_len <= 8)");
  prompt_weights_kernel<<<B, 32 * L, 0, stream>>>(emb, emb_ld, weight, bias, C, L, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_prompt_mix_resize(const float* prompt, int L, int D, int S, const float* wts, int B, int H, int W,
                                     void* out16, long long out_ld, int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(prompt && wts && out16 && L >= 1 && D > 0 && D % 8 == 0 && S > 0 && B > 0 && H > 0 && W > 0,
                "tdr_prompt_mix_resize: bad arguments (prompt_dim must be a multiple of 8)");
  TDR_CHECK_ARG(out_ld >= D && out_ld % 8 == 0 && ((uintptr_t)out16 & 15) == 0, "tdr_prompt_mix_resize: bad output rows");
  const long long total = (long long)B * H * W * (D / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  prompt_mix_resize_kernel<<<(unsigned)blocks, 256, 0, stream>>>(prompt, L, D, S, wts, B, H, W,
                                                                 reinterpret_cast<uint16_t*>(out16), out_ld, fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
