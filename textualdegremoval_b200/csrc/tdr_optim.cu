// Optimizer tail of the DDP training step (reference: models/image_restoration_ref_model.py:276-284 --
// `clip_grad_norm_(net_g.parameters(), 0.01)` then `AdamW.step()` over two LR groups -- and the EMA update
// models/base_model.py:54-62).  Gradients live in ONE flat fp32 buffer per parameter group (the buffer NCCL
// all-reduces in buckets), so the whole tail is three passes over flat memory with no host synchronisation:
//   tdr_sumsq_partial (per-block partial sums of squares, deterministic) -> tdr_clip_coef (device scalar) ->
//   tdr_adamw_step (clip * 1/world applied on the fly, decoupled weight decay, bias correction) [-> tdr_ema_update].
#include "tdr_common.cuh"

namespace {

constexpr int kSumsqBlocks = 1184;   // 148 SMs x 8

__global__ void __launch_bounds__(256) sumsq_partial_kernel(const float* __restrict__ g, long long n,
                                                            float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = g4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) s += g[i] * g[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

// coef = min(1, max_norm / (scale * sqrt(sum partials) + 1e-6))   (torch.nn.utils.clip_grad_norm_ semantics)
__global__ void clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm, float grad_scale,
                                 float* __restrict__ out /* [2]: coef, total_norm */) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    const float norm = grad_scale * (float)sqrt(t);
    const float c = max_norm / (norm + 1e-6f);
    out[0] = c < 1.f ? c : 1.f;
    out[1] = norm;
  }
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n,
                                                    float lr, float beta1, float beta2, float eps, float wd,
                                                    float bc1, float bc2_sqrt, float grad_scale,
                                                    const float* __restrict__ clip_coef) {
  const float gs = grad_scale * (clip_coef ? clip_coef[0] : 1.f);
  const float step_size = lr / bc1;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    pi -= step_size * mi / (sqrtf(vi) / bc2_sqrt + eps);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

__global__ void ema_kernel(float* __restrict__ ema, const float* __restrict__ p, long long n, float decay) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    ema[i] = decay * ema[i] + (1.f - decay) * p[i];
}

// L1 pixel loss (losses L1Loss 'mean', loss_weight w): partial[blk] = sum |x - gt| over the block's elements,
// dx = w / n * sign(x - gt)   (the gradient autograd would hand to the network output)
__global__ void __launch_bounds__(256) l1_loss_grad_kernel(const float* __restrict__ x, const float* __restrict__ gt,
                                                           long long n, float gscale, float* __restrict__ dx,
                                                           float* __restrict__ partial) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = x[i] - gt[i];
    s += fabsf(d);
    dx[i] = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void l1_loss_final_kernel(const float* __restrict__ partial, int n, float scale, float* __restrict__ loss) {
  __shared__ double red[8];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    loss[0] = (float)(t * scale);
  }
}

inline int grid_flat(long long n) {
  long long g = (n + 1023) / 1024;
  const long long cap = (long long)tdr_num_sms() * 8;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" int tdr_sumsq_partial_count(void) { return kSumsqBlocks; }

extern "C" int tdr_sumsq_partial(const float* g, long long n, float* partial, cudaStream_t stream) {
  TDR_CHECK_ARG(g && partial && n > 0 && ((uintptr_t)g & 15) == 0, "tdr_sumsq_partial: bad arguments");
  sumsq_partial_kernel<<<kSumsqBlocks, 256, 0, stream>>>(g, n, partial);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_clip_coef(const float* partial, int n, float max_norm, float grad_scale, float* out2,
                             cudaStream_t stream) {
  TDR_CHECK_ARG(partial && out2 && n > 0 && max_norm > 0.f, "tdr_clip_coef: bad arguments");
  clip_coef_kernel<<<1, 256, 0, stream>>>(partial, n, max_norm, grad_scale, out2);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, float grad_scale,
                              const float* clip_coef, cudaStream_t stream) {
  TDR_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "tdr_adamw_step: bad arguments");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  adamw_kernel<<<grid_flat(n), 256, 0, stream>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt,
                                                 grad_scale, clip_coef);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_ema_update(float* ema, const float* p, long long n, float decay, cudaStream_t stream) {
  TDR_CHECK_ARG(ema && p && n > 0, "tdr_ema_update: bad arguments");
  ema_kernel<<<grid_flat(n), 256, 0, stream>>>(ema, p, n, decay);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_l1_loss_grad(const float* x, const float* gt, long long n, float loss_weight, float* dx, float* loss,
                                float* partial /* tdr_sumsq_partial_count() floats */, cudaStream_t stream) {
  TDR_CHECK_ARG(x && gt && dx && loss && partial && n > 0, "tdr_l1_loss_grad: bad arguments");
  const int blocks = grid_flat(n) < kSumsqBlocks ? grid_flat(n) : kSumsqBlocks;
  l1_loss_grad_kernel<<<blocks, 256, 0, stream>>>(x, gt, n, loss_weight / (float)n, dx, partial);
  TDR_CHECK_LAUNCH();
  l1_loss_final_kernel<<<1, 256, 0, stream>>>(partial, blocks, (double)loss_weight / (double)n, loss);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
