// tdr_prepare_patches: the dataset's per-sample tensor preparation on the device (the caller side of
// the path).  Reference, per sample of Dataset_PairedImageWithRef (data/restoration_dataset.py:194-253):
//   imfrombytes(float32=True)      uint8 BGR HWC -> float32 / 255.            utils/utils_image.py:194-218
//   padding                        cv2.copyMakeBorder(..., BORDER_REFLECT) bottom / right up to gt_size   :243-254
//   paired_random_crop             [top : top + size, left : left + size]     data/transforms.py:24-83
//   random_augmentation            one of 8 flip / rot90 modes                data/transforms.py:223-275
//   img2tensor(bgr2rgb=True)       BGR -> RGB, HWC -> CHW                     utils/utils_image.py:102-126
//   normalize(mean, std)           (x - mean) / std  (optional)               restoration_dataset.py:240-244
// The random decisions (top, left, mode) stay on the host with the reference's own `random` calls; the kernel is pure
// data movement plus two IEEE operations per element, so the result is bit-identical to the reference's.  One thread
// per output pixel (all channels): reads are 3 B gathers from the decoded frame (L2-resident, frames are a few MB),
// writes are coalesced along the output row of each channel plane.
#include "tdr_common.cuh"

namespace {

inline int grid_for(long long items, int per_block, int max_waves) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = (long long)tdr_num_sms() * max_waves;
  if (g > cap) g = cap;
  return g < 1 ? 1 : (int)g;
}

__device__ __forceinline__ int reflect_index(int i, int n) {      // cv2.BORDER_REFLECT: fedcba|abcdefgh|hgfedcb
  if (i < n) return i;
  const int period = 2 * n;
  i %= period;
  return i < n ? i : period - 1 - i;
}

__global__ void __launch_bounds__(256) prepare_patches_kernel(const tdr_patch_desc* __restrict__ descs, int n, int channels,
                                                              int out_h, int out_w, int bgr2rgb, float3 mean, float3 stdv,
                                                              int normalize, float* __restrict__ out) {
  const long long total = (long long)n * out_h * out_w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % out_w);
    const int i = (int)((idx / out_w) % out_h);
    const int s = (int)(idx / ((long long)out_w * out_h));
    const tdr_patch_desc d = descs[s];
    // output (i, j) <- patch (y, x): inverse of data_augmentation (np.rot90 = counter-clockwise, then np.flipud)
    int y, x;
    switch (d.mode) {
      case 1: y = out_h - 1 - i; x = j; break;                    // flipud
      case 2: y = j; x = out_h - 1 - i; break;                    // rot90            (square patches)
      case 3: y = j; x = i; break;                                // rot90 + flipud = transpose
      case 4: y = out_h - 1 - i; x = out_w - 1 - j; break;        // rot180
      case 5: y = i; x = out_w - 1 - j; break;                    // rot180 + flipud = fliplr
      case 6: y = out_w - 1 - j; x = i; break;                    // rot270
      case 7: y = out_w - 1 - j; x = out_h - 1 - i; break;        // rot270 + flipud = anti-transpose
      default: y = i; x = j; break;
    }
    const int sy = reflect_index(d.top + y, d.h), sx = reflect_index(d.left + x, d.w);
    const unsigned char* px = reinterpret_cast<const unsigned char*>(d.image) + ((long long)sy * d.w + sx) * channels;
    const float m[3] = {mean.x, mean.y, mean.z}, sd[3] = {stdv.x, stdv.y, stdv.z};
    for (int c = 0; c < channels; ++c) {
      const int src_c = (channels == 3 && bgr2rgb) ? 2 - c : c;
      float v = __fdiv_rn((float)px[src_c], 255.f);               // img.astype(np.float32) / 255.
      if (normalize) v = __fdiv_rn(__fsub_rn(v, m[c]), sd[c]);    // tensor.sub_(mean).div_(std)
      out[(((long long)s * channels + c) * out_h + i) * out_w + j] = v;
    }
  }
}

}  // namespace

extern "C" int tdr_prepare_patches(const tdr_patch_desc* descs_device, const tdr_patch_desc* descs_host, int n,
                                   int channels, int out_h, int out_w, int bgr2rgb, const float* mean, const float* stdv,
                                   float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(descs_device && descs_host && out, "tdr_prepare_patches: null pointer");
  TDR_CHECK_ARG(n > 0 && out_h > 0 && out_w > 0 && (channels == 1 || channels == 3),
                "tdr_prepare_patches: need n > 0, a non-empty patch and 1 or 3 channels");
  TDR_CHECK_ARG((mean == nullptr) == (stdv == nullptr), "tdr_prepare_patches: mean and std go together");
  for (int s = 0; s < n; ++s) {
    const tdr_patch_desc& d = descs_host[s];
    TDR_CHECK_ARG(d.image && d.h > 0 && d.w > 0, "tdr_prepare_patches: sample %d has no image", s);
    TDR_CHECK_ARG(d.mode >= 0 && d.mode <= 7, "tdr_prepare_patches: sample %d: augmentation mode %d not in 0..7", s, d.mode);
    TDR_CHECK_ARG(d.top >= 0 && d.left >= 0, "tdr_prepare_patches: sample %d: negative crop origin", s);
    const bool transposing = d.mode == 2 || d.mode == 3 || d.mode == 6 || d.mode == 7;
    TDR_CHECK_ARG(!transposing || out_h == out_w, "tdr_prepare_patches: rot90 modes need a square patch");
    // the reference raises when the (reflect-padded) frame is smaller than the patch: padding only fills up to the
    // patch size, so the crop must lie inside max(h, size) x max(w, size)  (data/transforms.py:58-62)
    const int ph = d.h > out_h ? d.h : out_h, pw = d.w > out_w ? d.w : out_w;
    TDR_CHECK_ARG(d.top + out_h <= ph && d.left + out_w <= pw, "tdr_prepare_patches: sample %d: crop outside the frame", s);
  }
  float3 m = make_float3(0.f, 0.f, 0.f), sd = make_float3(1.f, 1.f, 1.f);
  if (mean) {
    m = make_float3(mean[0], channels == 3 ? mean[1] : 0.f, channels == 3 ? mean[2] : 0.f);
    sd = make_float3(stdv[0], channels == 3 ? stdv[1] : 1.f, channels == 3 ? stdv[2] : 1.f);
  }
  const long long total = (long long)n * out_h * out_w;
  prepare_patches_kernel<<<grid_for(total, 256, 8), 256, 0, stream>>>(descs_device, n, channels, out_h, out_w, bgr2rgb, m,
                                                                     sd, mean != nullptr, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

// ------------------------------------------------------------------------------------------------ validation metric
// tdr_psnr_u8_sums: the integer part of the reference's validation PSNR (use_image: true) without moving images to the
// host.  Reference: tensor2img (utils/utils_image.py:129-191: clamp to [0, 1], * 255.0, numpy round = half-to-even,
// uint8) on `result` and `gt`, then calculate_psnr (metrics/psnr_ssim.py:9-63: crop_border, float64 mean of squared
// differences, max_value = 1 if img1.max() <= 1 else 255).  Squared differences of uint8 values are integers, so their
// sum is exact in int64 whatever the order: the kernel returns, per image, sum (qa - qb)^2 over the cropped window and
// max(qa); the host finishes with the reference's own float64 formula and gets the identical double.
namespace {

__device__ __forceinline__ int quant_u8(float x) {
  x = fminf(fmaxf(x, 0.f), 1.f);                  // clamp_(0, 1); (x - 0) / (1 - 0) is exact
  return (int)rintf(__fmul_rn(x, 255.f));         // (img_np * 255.0).round() on float32, half-to-even
}

__global__ void __launch_bounds__(256) psnr_u8_sums_kernel(const float* __restrict__ a, const float* __restrict__ b, int C,
                                                           int H, int W, int crop, unsigned long long* __restrict__ sse,
                                                           int* __restrict__ max_a) {
  const int img = blockIdx.y;
  const int h2 = H - 2 * crop, w2 = W - 2 * crop;
  const long long n = (long long)C * h2 * w2;
  const float* pa = a + (long long)img * C * H * W;
  const float* pb = b + (long long)img * C * H * W;
  unsigned long long acc = 0;
  int mx = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w2);
    const int y = (int)((i / w2) % h2);
    const int c = (int)(i / ((long long)w2 * h2));
    const long long off = ((long long)c * H + y + crop) * W + x + crop;
    const int qa = quant_u8(pa[off]), qb = quant_u8(pb[off]);
    const int d = qa - qb;
    acc += (unsigned long long)(d * d);
    mx = max(mx, qa);
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc += __shfl_xor_sync(0xffffffffu, acc, o);
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {                  // integer atomics: order-independent, deterministic
    atomicAdd(&sse[img], acc);
    atomicMax(&max_a[img], mx);
  }
}

}  // namespace

extern "C" int tdr_psnr_u8_sums(const float* img1, const float* img2, int B, int C, int H, int W, int crop_border,
                                unsigned long long* sse, int* max1, cudaStream_t stream) {
  TDR_CHECK_ARG(img1 && img2 && sse && max1, "tdr_psnr_u8_sums: null pointer");
  TDR_CHECK_ARG(B <= 65535, "tdr_psnr_u8_sums: at most 65535 images per call (got %d)", B);
  TDR_CHECK_ARG(B > 0 && C > 0 && crop_border >= 0 && H > 2 * crop_border && W > 2 * crop_border,
                "tdr_psnr_u8_sums: empty window (B %d C %d H %d W %d crop %d)", B, C, H, W, crop_border);
  TDR_CHECK_CUDA(cudaMemsetAsync(sse, 0, sizeof(unsigned long long) * B, stream));
  TDR_CHECK_CUDA(cudaMemsetAsync(max1, 0, sizeof(int) * B, stream));
  const long long n = (long long)C * (H - 2 * crop_border) * (W - 2 * crop_border);
  int gx = grid_for(n, 256 * 8, 4);
  psnr_u8_sums_kernel<<<dim3(gx, B), 256, 0, stream>>>(img1, img2, C, H, W, crop_border, sse, max1);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}


// ------------------------------------------------------------------------------------------------ Gaussian training noise
// Dataset_GaussianDenoisingWithRef.__getitem__ (data/restoration_dataset.py:474-476):
//     noise_level = sigma / 255;  noise = torch.randn(img.size()).mul_(noise_level);  img_lq.add_(noise)
// on the device.  Two modes: (a) `noise` given (standard normals drawn by the caller, e.g. from the reference's CPU
// generator): out = img + fl(noise * level), the same two roundings as mul_ / add_, i.e. bit-identical to the reference
// for that draw; (b) `noise` null: standard normals from a counter-based Philox4x32-10 stream keyed by (seed, sample,
// element / 4) + Box-Muller, so a batch needs no host RNG, no H2D copy of the noise and is reproducible from the seed.
namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t* out) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void gaussian_noise_kernel(const float* __restrict__ img, const float* __restrict__ noise,
                                      const float* __restrict__ level, long long per_sample, int B,
                                      unsigned long long seed, float* __restrict__ out) {
  const long long quads = (per_sample + 3) / 4;
  const long long total = quads * B;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / quads);
    const long long qd = i % quads;
    const float lv = level[b];
    float z[4];
    if (!noise) {
      uint32_t r[4];
      philox4x32_10((uint32_t)qd, (uint32_t)(qd >> 32), (uint32_t)b, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
      // two Box-Muller pairs from four 32-bit words; u in (0, 1]
      const float u0 = ((float)r[0] + 1.0f) * 2.3283064365386963e-10f, u1 = ((float)r[2] + 1.0f) * 2.3283064365386963e-10f;
      const float a0 = (float)r[1] * 1.4629180792671596e-9f, a1 = (float)r[3] * 1.4629180792671596e-9f;   // 2 pi / 2^32
      const float m0 = sqrtf(-2.f * logf(u0)), m1 = sqrtf(-2.f * logf(u1));
      float s0, c0, s1, c1;
      sincosf(a0, &s0, &c0);
      sincosf(a1, &s1, &c1);
      z[0] = m0 * c0; z[1] = m0 * s0; z[2] = m1 * c1; z[3] = m1 * s1;
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const long long j = qd * 4 + e;
      if (j < per_sample) {
        const long long o = (long long)b * per_sample + j;
        const float n = noise ? noise[o] : z[e];
        out[o] = __fadd_rn(img[o], __fmul_rn(n, lv));         // mul_ then add_: two roundings, no FMA contraction
      }
    }
  }
}

}  // namespace

extern "C" int tdr_add_gaussian_noise(const float* img, const float* noise, const float* level_device, int B,
                                      long long per_sample, unsigned long long seed, float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(img && level_device && out && B > 0 && per_sample > 0, "tdr_add_gaussian_noise: bad arguments");
  const long long total = (per_sample + 3) / 4 * B;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  gaussian_noise_kernel<<<(unsigned)blocks, 256, 0, stream>>>(img, noise, level_device, per_sample, B, seed, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
