// Row / token kernels of the ViT encoders (DINOv2 ViT-B/14: /root/reference/models/dino/*.py; CLIP ViT-H/14: the
// transformers CLIPVisionModel the reference calls at scripts/train/main_train_tr_mapping.py:780) and the mapper MLPs
// (main_train_tr_mapping.py:40-122).  All dense contractions (patch embedding, qkv / proj / fc1 / fc2, q.k^T, p.v, the
// mapper linears) run through tdr_conv_gemm on the flat [1 x tokens x channels] view; these kernels are the glue:
// patch extraction, token assembly (+cls, +pos), fp32 softmax rows, V transposition, token mean.
#include "tdr_common.cuh"

namespace {

// images NCHW fp32 -> patches bf16 [B, gh*gw, ld] with K index = (c*ps + ky)*ps + kx (the flattening of Conv2d weights)
__global__ void patchify_kernel(const float* __restrict__ img, int B, int C, int H, int W, int ps, int K, bf16* out,
                                long long ld) {
  const int gw = W / ps, gh = H / ps;
  const long long total = (long long)B * gh * gw * K;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long p = i / K;
    const int px = (int)(p % gw), py = (int)((p / gw) % gh), b = (int)(p / ((long long)gw * gh));
    const int kx = k % ps, ky = (k / ps) % ps, c = k / (ps * ps);
    out[p * ld + k] = __float2bfloat16(img[(((long long)b * C + c) * H + py * ps + ky) * W + px * ps + kx]);
  }
}

// x[b, 0, :] = cls + pos[0];  x[b, 1 + t, :] = patch[b, t, :] + pos[1 + t]      (fp32, D % 4 == 0)
__global__ void assemble_tokens_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                       const float* __restrict__ pos, int B, int N, int D, float* __restrict__ x) {
  const int nv = D >> 2;
  const long long total = (long long)B * (N + 1) * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nv);
    const long long r = i / nv;
    const int t = (int)(r % (N + 1)), b = (int)(r / (N + 1));
    const float4 pe = *reinterpret_cast<const float4*>(pos + (size_t)t * D + v * 4);
    float4 s = t == 0 ? *reinterpret_cast<const float4*>(cls + v * 4)
                      : *reinterpret_cast<const float4*>(patch + ((size_t)b * N + t - 1) * D + v * 4);
    s.x += pe.x; s.y += pe.y; s.z += pe.z; s.w += pe.w;
    *reinterpret_cast<float4*>(x + r * D + v * 4) = s;
  }
}

// one warp per row: p = softmax(scale * s[row, 0:n]) -> bf16, columns [n, ld_out) zeroed
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ s, long long ld, long long rows, int n,
                                                           float scale, bf16* __restrict__ out, long long ld_out) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* sr = s + row * ld;
  float mx = -INFINITY;
  for (int j = lane; j < n; j += 32) mx = fmaxf(mx, sr[j]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < n; j += 32) sum += __expf((sr[j] - mx) * scale);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  bf16* orow = out + row * ld_out;
  for (int j = lane; j < (int)ld_out; j += 32)
    orow[j] = __float2bfloat16(j < n ? __expf((sr[j] - mx) * scale) * inv : 0.f);
}

// vt[b, h, d, t] = qkv[b, t, voff + h*hd + d]   (tokens contiguous = K-major "weights" of the p.v GEMM), pad zeroed
__global__ void transpose_v_kernel(const bf16* __restrict__ qkv, long long ld, int B, int N, int heads, int hd, int voff,
                                   bf16* __restrict__ vt, int n_pad) {
  __shared__ bf16 tile[32][33];
  const int b = blockIdx.z / heads, h = blockIdx.z % heads;
  const int t0 = blockIdx.x * 32, d0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int t = t0 + r, d = d0 + threadIdx.x;
    tile[r][threadIdx.x] = (t < N && d < hd) ? qkv[((size_t)b * N + t) * ld + voff + h * hd + d] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int d = d0 + r, t = t0 + threadIdx.x;
    if (d < hd && t < n_pad) vt[(((size_t)b * heads + h) * hd + d) * n_pad + t] = tile[threadIdx.x][r];
  }
}

// out[b, c] (+)= mean over t of x[b, t0 + t, c], t < n  (fp32)
__global__ void mean_tokens_kernel(const float* __restrict__ x, long long ld, int tokens_per_b, int t0, int n, int C,
                                   float* __restrict__ out, long long out_ld, int accumulate) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int t = 0; t < n; ++t) s += x[((size_t)b * tokens_per_b + t0 + t) * ld + c];
  s /= (float)n;
  float* o = out + (size_t)b * out_ld + c;
  *o = accumulate ? *o + s : s;
}

inline int grid1(long long items, int per_block) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = (long long)tdr_num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" int tdr_vit_patchify(const float* img, int B, int C, int H, int W, int patch, void* out_bf16, long long ld,
                                cudaStream_t stream) {
  TDR_CHECK_ARG(img && out_bf16 && patch > 0 && H % patch == 0 && W % patch == 0,
                "tdr_vit_patchify: H and W must be multiples of the patch size (%d): got %dx%d", patch, H, W);
  const int K = C * patch * patch;
  TDR_CHECK_ARG(ld >= K, "tdr_vit_patchify: ld too small");
  patchify_kernel<<<grid1((long long)B * (H / patch) * (W / patch) * K, 256), 256, 0, stream>>>(
      img, B, C, H, W, patch, K, reinterpret_cast<bf16*>(out_bf16), ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_vit_assemble_tokens(const float* patch_tokens, const float* cls, const float* pos, int B, int N, int D,
                                       float* x, cudaStream_t stream) {
  TDR_CHECK_ARG(patch_tokens && cls && pos && x && D % 4 == 0, "tdr_vit_assemble_tokens: bad arguments");
  assemble_tokens_kernel<<<grid1((long long)B * (N + 1) * (D / 4), 256), 256, 0, stream>>>(patch_tokens, cls, pos, B, N,
                                                                                           D, x);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_softmax_rows(const float* s, long long ld, long long rows, int n, float scale, void* out_bf16,
                                long long ld_out, cudaStream_t stream) {
  TDR_CHECK_ARG(s && out_bf16 && rows > 0 && n > 0 && ld >= n && ld_out >= n, "tdr_softmax_rows: bad arguments");
  softmax_rows_kernel<<<(int)((rows * 32 + 255) / 256), 256, 0, stream>>>(s, ld, rows, n, scale,
                                                                         reinterpret_cast<bf16*>(out_bf16), ld_out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_vit_transpose_v(const void* qkv_bf16, long long ld, int B, int N, int heads, int hd, int voff,
                                   void* vt_bf16, int n_pad, cudaStream_t stream) {
  TDR_CHECK_ARG(qkv_bf16 && vt_bf16 && n_pad >= N, "tdr_vit_transpose_v: bad arguments");
  dim3 grid((n_pad + 31) / 32, (hd + 31) / 32, B * heads), block(32, 8);
  transpose_v_kernel<<<grid, block, 0, stream>>>(reinterpret_cast<const bf16*>(qkv_bf16), ld, B, N, heads, hd, voff,
                                                 reinterpret_cast<bf16*>(vt_bf16), n_pad);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_mean_tokens(const float* x, long long ld, int B, int tokens_per_b, int t0, int n, int C, float* out,
                               long long out_ld, int accumulate, cudaStream_t stream) {
  TDR_CHECK_ARG(x && out && n > 0 && t0 >= 0 && t0 + n <= tokens_per_b, "tdr_mean_tokens: bad arguments");
  dim3 grid((C + 127) / 128, B);
  mean_tokens_kernel<<<grid, 128, 0, stream>>>(x, ld, tokens_per_b, t0, n, C, out, out_ld, accumulate);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

// ------------------------------------------------------------------------------------------------ crop selection
namespace {
// out[c_idx, ch, oy, ox] = bilinear(align_corners=False) sample of the (ch x cw) crop of img[b] at origin[c_idx]
// (F.unfold + F.interpolate of models/image_restoration_ref_model.py:219-227 in one pass; identity when oh==ch).
__global__ void crop_resize_kernel(const float* __restrict__ img, int C, int H, int W, const int* __restrict__ origin,
                                   int ncrops, int ch, int cw, int oh, int ow, float* __restrict__ out) {
  const long long total = (long long)ncrops * C * oh * ow;
  const float sy = (float)ch / oh, sx = (float)cw / ow;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
    const int c = (int)((i / ((long long)ow * oh)) % C), k = (int)(i / ((long long)ow * oh * C));
    const int b = origin[3 * k], y0 = origin[3 * k + 1], x0 = origin[3 * k + 2];
    float fy = (oy + 0.5f) * sy - 0.5f, fx = (ox + 0.5f) * sx - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int iy0 = (int)fy, ix0 = (int)fx;
    const int iy1 = iy0 + 1 < ch ? iy0 + 1 : ch - 1, ix1 = ix0 + 1 < cw ? ix0 + 1 : cw - 1;
    const float ly = fy - iy0, lx = fx - ix0;
    const float* p = img + ((long long)b * C + c) * H * W;
    const float v00 = p[(long long)(y0 + iy0) * W + x0 + ix0], v01 = p[(long long)(y0 + iy0) * W + x0 + ix1];
    const float v10 = p[(long long)(y0 + iy1) * W + x0 + ix0], v11 = p[(long long)(y0 + iy1) * W + x0 + ix1];
    out[i] = (1.f - ly) * ((1.f - lx) * v00 + lx * v01) + ly * ((1.f - lx) * v10 + lx * v11);
  }
}

// grid (n, B): cos[b, k] = <fl[b], fr[b*n + k]> / (max(|fl|,eps) max(|fr|,eps))  over F features (fp32)
__global__ void __launch_bounds__(256) cosine_kernel(const float* __restrict__ fl, const float* __restrict__ fr, int n,
                                                     long long F, float* __restrict__ cosv) {
  __shared__ float red[3][8];
  const int k = blockIdx.x, b = blockIdx.y;
  const float* a = fl + (size_t)b * F;
  const float* r = fr + ((size_t)b * n + k) * F;
  float dot = 0.f, na = 0.f, nr = 0.f;
  for (long long i = threadIdx.x; i < F; i += blockDim.x) {
    const float x = a[i], y = r[i];
    dot = fmaf(x, y, dot);
    na = fmaf(x, x, na);
    nr = fmaf(y, y, nr);
  }
  dot = warp_sum(dot); na = warp_sum(na); nr = warp_sum(nr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][warp] = dot; red[1][warp] = na; red[2][warp] = nr; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float d = 0.f, x = 0.f, y = 0.f;
    for (int w = 0; w < 8; ++w) { d += red[0][w]; x += red[1][w]; y += red[2][w]; }
    cosv[(size_t)b * n + k] = d / (fmaxf(sqrtf(x), 1e-12f) * fmaxf(sqrtf(y), 1e-12f));
  }
}

}  // namespace

extern "C" int tdr_crop_resize(const float* img, int C, int H, int W, const int* origin, int ncrops, int crop_h,
                               int crop_w, int out_h, int out_w, float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(img && origin && out && ncrops > 0 && crop_h > 0 && crop_w > 0 && out_h > 0 && out_w > 0,
                "tdr_crop_resize: bad arguments");
  crop_resize_kernel<<<grid1((long long)ncrops * C * out_h * out_w, 256), 256, 0, stream>>>(
      img, C, H, W, origin, ncrops, crop_h, crop_w, out_h, out_w, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_cosine_rows(const float* fl, const float* fr, int B, int n, long long F, float* cosv,
                               cudaStream_t stream) {
  TDR_CHECK_ARG(fl && fr && cosv && B > 0 && n > 0 && F > 0, "tdr_cosine_rows: bad arguments");
  dim3 grid(n, B);
  cosine_kernel<<<grid, 256, 0, stream>>>(fl, fr, n, F, cosv);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
