// Generic grouped / depthwise K x K stencil (K in {1, 3, 5, 7}, any dilation with dil * (K - 1) / 2 <= 6) over NHWC 16-bit
// activations, for the DRSformer family of the reference (/root/reference/models/archs/network_drsformer_guided_arch.py):
//   * MSFN feed-forward :216-256 -- depthwise 3x3 / 5x5 + ReLU on the 2h hidden channels, then the grouped convs
//     ``Conv2d(2h, h, k, groups=h)`` (two input channels per output channel) over the re-interleaved halves;
//   * MEFC experts :454-520 -- SepConv / DilConv depthwise stages (1x1 .. 7x7, dilation 2), AvgPool2d(3, count_include_pad
//     = False) :440.
// Every output channel names its ``ipg`` (1 or 2) input channels through an index table, so the chunk / cat re-orderings
// of the reference (:243-252) are address arithmetic, never copies.  A CTA computes an 8 x 32 pixel tile of 8 output
// channels from a haloed fp32 tile in shared memory (each staged value is reused K*K times); fp32 accumulation;
// a thread owns 2 channels x 4 adjacent pixels with the taps in registers.
// These layers are HBM / shared-memory-bound byte work on CUDA cores; the 3x3 depthwise convs of the Restormer blocks keep
// their dedicated TMA kernel (tdr_dwconv3x3).
#include "tdr_common.cuh"

namespace {

constexpr int kTH = 8, kTW = 32, kCo = 8;        // tile: 8 rows x 32 columns x 8 output channels

struct StencilArgs {
  const uint16_t* in; long long in_ld;
  int B, H, W, Co, ipg, K, dil, act, pool, fp16;
  const int* idx; const float* w; const float* bias;
  uint16_t* out; long long out_ld;
  int tiles_x, tiles_y;
};

// Thread mapping: 256 threads = 4 channel pairs x 64 pixel quads; a thread computes 2 output channels x 4 horizontally
// adjacent pixels.  Per (channel, input, tap row) it loads the 4 + (K-1) dil staged values once and reuses each for up to K
// outputs; the K*K taps of the channel sit in registers (K is a template parameter, so the loops unroll).
template <int K, int IPG>
__global__ void __launch_bounds__(256) grouped_stencil_kernel(const StencilArgs a) {
  extern __shared__ float tile[];                  // [nci][TH + 2p][TW + 2p]
  __shared__ int s_idx[kCo * 2];
  const int p = a.dil * (K - 1) / 2;
  const int th = kTH + 2 * p, tw = kTW + 2 * p;
  // row pitch == 1 (mod 4): a warp reads 4 rows x 8 quads (lane stride 4 floats within a row); with an even pitch rows
  // collide 4-way on the shared-memory banks, with pitch 4m + 1 the 32 lanes hit 32 different banks
  const int twp = tw + ((1 - tw) & 3);
  constexpr int nci = kCo * IPG;
  int t = blockIdx.x;
  const int tx = t % a.tiles_x; t /= a.tiles_x;
  const int ty = t % a.tiles_y;
  const int b = t / a.tiles_y;
  const int co0 = blockIdx.y * kCo;
  const int y0 = ty * kTH, x0 = tx * kTW;
  if (threadIdx.x < nci) {
    const int co = co0 + threadIdx.x / IPG;
    s_idx[threadIdx.x] = co < a.Co ? a.idx[(size_t)co * IPG + threadIdx.x % IPG] : -1;
  }
  __syncthreads();
  // stage the haloed input tile.  Fast path: the CTA's input channels are groups of 8 contiguous, 16-byte aligned
  // channels (depthwise tables, and grouped tables away from the half boundary): one 128-bit load per (pixel, group).
  const uint16_t* img = a.in + (size_t)b * a.H * a.W * a.in_ld;
  bool vec = (a.in_ld & 7) == 0 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0 && co0 + kCo <= a.Co;
  for (int i = 0; i < nci && vec; ++i) vec = s_idx[i] == s_idx[i & ~7] + (i & 7) && (s_idx[i & ~7] & 7) == 0;
  // (row loops instead of a flat index: the flat form spent most of the kernel's instructions on runtime div / mod)
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (vec) {
    constexpr int ng = nci >> 3;
    for (int r = wid; r < th; r += 8) {
      const int yy = y0 - p + r;
      const bool row_ok = yy >= 0 && yy < a.H;
      for (int j = lane; j < tw * ng; j += 32) {
        const int g = j % ng, xc = j / ng;
        const int xx = x0 - p + xc;
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (row_ok && xx >= 0 && xx < a.W) unpack8r(img + ((size_t)yy * a.W + xx) * a.in_ld + s_idx[g * 8], v, a.fp16);
#pragma unroll
        for (int e = 0; e < 8; ++e) tile[((g * 8 + e) * th + r) * twp + xc] = v[e];
      }
    }
  } else {
    for (int r = wid; r < th; r += 8) {
      const int yy = y0 - p + r;
      const bool row_ok = yy >= 0 && yy < a.H;
      for (int j = lane; j < tw * nci; j += 32) {
        const int ci = j % nci, xc = j / nci;
        const int xx = x0 - p + xc;
        const int ch = s_idx[ci];
        float v = 0.f, d;
        if (ch >= 0 && row_ok && xx >= 0 && xx < a.W)
          unpack2r(img[((size_t)yy * a.W + xx) * a.in_ld + ch], v, d, a.fp16);
        tile[(ci * th + r) * twp + xc] = v;
      }
    }
  }
  __syncthreads();
  const int cp = threadIdx.x >> 6, q = threadIdx.x & 63;          // channel pair, pixel quad
  const int ly = q >> 3, lx = (q & 7) * 4;
  const int y = y0 + ly;
  constexpr int NV = 4 + (K - 1) * 2;                              // staged values per tap row at dil <= 2
  uint16_t res[2][4];
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    const int c = cp * 2 + cc, co = co0 + c;
#pragma unroll
    for (int i = 0; i < 4; ++i) res[cc][i] = 0;
    if (co >= a.Co || y >= a.H) continue;
    float acc[4];
    const float bias = a.bias ? a.bias[co] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = bias;
    if (a.pool) {                                  // AvgPool2d(3, 1, 1, count_include_pad=False)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int x = x0 + lx + i;
        int cnt = 0;
        float sum = 0.f;
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            if (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) {
              ++cnt;
              sum += tile[(c * th + ly + ky) * twp + lx + i + kx];
            }
          }
        acc[i] = cnt ? sum / (float)cnt : 0.f;
      }
    } else {
#pragma unroll
      for (int j = 0; j < IPG; ++j) {
        const float* tp = tile + (size_t)(c * IPG + j) * th * twp;
        const float* wp = a.w + ((size_t)co * IPG + j) * (K * K);
        float w[K * K];
#pragma unroll
        for (int i = 0; i < K * K; ++i) w[i] = __ldg(wp + i);
#pragma unroll
        for (int ky = 0; ky < K; ++ky) {
          const float* row = tp + (ly + ky * a.dil) * twp + lx;
          float v[NV];
          const int nv = 4 + (K - 1) * a.dil;
#pragma unroll
          for (int i = 0; i < NV; ++i) v[i] = i < nv ? row[i] : 0.f;
          if (a.dil == 1) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(w[ky * K + kx], v[i + kx], acc[i]);
          } else {
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[i] = fmaf(w[ky * K + kx], v[i + 2 * kx], acc[i]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) res[cc][i] = pack1r(a.act == 1 ? fmaxf(acc[i], 0.f) : acc[i], a.fp16);
  }
  // results -> shared memory [pixel][8 channels] (over the consumed input tile), then one 128-bit store per pixel
  __syncthreads();
  uint16_t* so = reinterpret_cast<uint16_t*>(tile);
#pragma unroll
  for (int cc = 0; cc < 2; ++cc)
#pragma unroll
    for (int i = 0; i < 4; ++i) so[(ly * kTW + lx + i) * kCo + cp * 2 + cc] = res[cc][i];
  __syncthreads();
  {
    const int px = threadIdx.x, py = px / kTW, pxx = px % kTW;
    const int yy = y0 + py, xx = x0 + pxx;
    if (yy < a.H && xx < a.W) {
      uint16_t* dst = a.out + (((size_t)b * a.H + yy) * a.W + xx) * a.out_ld + co0;
      const int nco = a.Co - co0 < kCo ? a.Co - co0 : kCo;
      if (nco == kCo && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(so + px * kCo);
      } else {
        for (int c = 0; c < nco; ++c) dst[c] = so[px * kCo + c];
      }
    }
  }
}

// MEFC gate (OALayer.forward :423-432 + the softmax of subnet.forward :541-543): one CTA per sample,
//   out[b, s, :] = softmax_over_ops( W2 relu(W1 emb[b] + b1) + b2 ) viewed as [steps, num_ops]
__global__ void __launch_bounds__(128) mefc_gate_kernel(const float* __restrict__ emb, long long emb_ld, int C,
                                                        const float* __restrict__ w1, const float* __restrict__ b1, int H1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2, int O,
                                                        int num_ops, float* __restrict__ out) {
  __shared__ float h[256], o[256];
  const int b = blockIdx.x;
  const float* e = emb + (size_t)b * emb_ld;
  for (int i = threadIdx.x; i < H1; i += blockDim.x) {
    float s = b1 ? b1[i] : 0.f;
    for (int c = 0; c < C; ++c) s = fmaf(w1[(size_t)i * C + c], e[c], s);
    h[i] = fmaxf(s, 0.f);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O; i += blockDim.x) {
    float s = b2 ? b2[i] : 0.f;
    for (int c = 0; c < H1; ++c) s = fmaf(w2[(size_t)i * H1 + c], h[c], s);
    o[i] = s;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < O; i += blockDim.x) {
    const int g0 = i / num_ops * num_ops;
    float mx = -INFINITY, sum = 0.f;
    for (int j = 0; j < num_ops; ++j) mx = fmaxf(mx, o[g0 + j]);
    for (int j = 0; j < num_ops; ++j) sum += expf(o[g0 + j] - mx);
    out[(size_t)b * O + i] = expf(o[i] - mx) / sum;
  }
}

// Per-sample weights of OperationLayer._out (:378-388): states[k] = op_k(x) * w[b, k] are concatenated and reduced by a 1x1
// conv, i.e. the conv's weight columns of block k are scaled by w[b, k]:  out16[b][co][k*C + ci] = W[co][k*C + ci] * gate[b, k]
__global__ void mefc_mix_weights_kernel(const float* __restrict__ w, int Co, int Ci, int C, const float* __restrict__ gate,
                                        long long gate_ld, int B, uint16_t* __restrict__ out, long long ld, int fp16) {
  const long long total = (long long)B * Co * Ci;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Ci);
    const int co = (int)((i / Ci) % Co);
    const int b = (int)(i / ((long long)Ci * Co));
    out[((size_t)b * Co + co) * ld + ci] = pack1r(w[(size_t)co * Ci + ci] * gate[(size_t)b * gate_ld + ci / C], fp16);
  }
}

}  // namespace

extern "C" int tdr_mefc_gate(const float* emb, long long emb_ld, int B, int C, const float* w1, const float* b1, int H1,
                             const float* w2, const float* b2, int O, int num_ops, float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(emb && w1 && w2 && out && B > 0 && C > 0 && H1 > 0 && H1 <= 256 && O > 0 && O <= 256 && num_ops > 0 &&
                O % num_ops == 0 && emb_ld >= C, "tdr_mefc_gate: bad arguments (hidden / output widths <= 256)");
  mefc_gate_kernel<<<B, 128, 0, stream>>>(emb, emb_ld, C, w1, b1, H1, w2, b2, O, num_ops, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_mefc_mix_weights(const float* weight, int Co, int Ci, int C, const float* gate, long long gate_ld, int B,
                                    void* out16, long long ld, int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(weight && gate && out16 && Co > 0 && Ci > 0 && C > 0 && Ci % C == 0 && B > 0 && ld >= Ci && ld % 8 == 0,
                "tdr_mefc_mix_weights: bad arguments");
  const long long total = (long long)B * Co * Ci;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  mefc_mix_weights_kernel<<<(unsigned)blocks, 256, 0, stream>>>(weight, Co, Ci, C, gate, gate_ld, B,
                                                               reinterpret_cast<uint16_t*>(out16), ld, fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

// out16[b, y, x, co] = act( bias[co] + sum_{j < ipg} sum_{taps} w[co][j][ky][kx] * in16[b, y + (ky - K/2) dil, x + (kx - K/2) dil, idx[co * ipg + j]] )
// zero padding; pool != 0: AvgPool2d(3, stride 1, padding 1, count_include_pad=False) of channel idx[co] (w ignored).
// ``out16`` points at the first output slot (a channel slice of a wider buffer is fine); fp16: IEEE fp16 rows, else bf16.
extern "C" int tdr_grouped_stencil(const void* in16, long long in_ld, int B, int H, int W, int Co, int ipg, const int* idx,
                                   const float* weight, const float* bias, int K, int dil, int act, int pool, void* out16,
                                   long long out_ld, int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(in16 && out16 && idx && (weight || pool) && B > 0 && H > 0 && W > 0 && Co > 0,
                "tdr_grouped_stencil: bad arguments");
  TDR_CHECK_ARG((ipg == 1 || ipg == 2) && (K == 1 || K == 3 || K == 5 || K == 7) && (dil == 1 || dil == 2),
                "tdr_grouped_stencil: ipg in {1, 2}, K in {1, 3, 5, 7}, dil in {1, 2} (got ipg %d K %d dil %d)", ipg, K, dil);
  TDR_CHECK_ARG(!pool || (K == 3 && dil == 1 && ipg == 1), "tdr_grouped_stencil: pooling is 3x3, one input channel");
  TDR_CHECK_ARG(act == 0 || act == 1, "tdr_grouped_stencil: act must be 0 or 1 (ReLU)");
  StencilArgs a;
  a.in = reinterpret_cast<const uint16_t*>(in16); a.in_ld = in_ld;
  a.B = B; a.H = H; a.W = W; a.Co = Co; a.ipg = ipg; a.K = K; a.dil = dil; a.act = act; a.pool = pool; a.fp16 = fp16 ? 1 : 0;
  a.idx = idx; a.w = weight; a.bias = bias;
  a.out = reinterpret_cast<uint16_t*>(out16); a.out_ld = out_ld;
  a.tiles_x = tdr_cdiv(W, kTW); a.tiles_y = tdr_cdiv(H, kTH);
  const int p = dil * (K - 1) / 2;
  const int tw = kTW + 2 * p;
  const size_t smem = (size_t)kCo * ipg * (kTH + 2 * p) * (tw + ((1 - tw) & 3)) * sizeof(float);
#define TDR_STENCIL_ATTR(KK, II) \
  TDR_CHECK_CUDA(cudaFuncSetAttribute(grouped_stencil_kernel<KK, II>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024))
  static bool attr_set = false;
  if (!attr_set) {
    TDR_STENCIL_ATTR(1, 1); TDR_STENCIL_ATTR(3, 1); TDR_STENCIL_ATTR(5, 1); TDR_STENCIL_ATTR(7, 1);
    TDR_STENCIL_ATTR(1, 2); TDR_STENCIL_ATTR(3, 2); TDR_STENCIL_ATTR(5, 2); TDR_STENCIL_ATTR(7, 2);
    attr_set = true;
  }
#undef TDR_STENCIL_ATTR
  TDR_CHECK_ARG(smem <= 64 * 1024, "tdr_grouped_stencil: tile does not fit (%zu bytes)", smem);
  TDR_CHECK_ARG((long long)a.tiles_x * a.tiles_y * B < (1LL << 31) && tdr_cdiv(Co, kCo) <= 65535, "tdr_grouped_stencil: grid");
  dim3 grid((unsigned)(a.tiles_x * a.tiles_y * B), (unsigned)tdr_cdiv(Co, kCo));
#define TDR_STENCIL_GO(KK) \
  do { if (ipg == 1) grouped_stencil_kernel<KK, 1><<<grid, 256, smem, stream>>>(a); \
       else grouped_stencil_kernel<KK, 2><<<grid, 256, smem, stream>>>(a); } while (0)
  switch (K) {
    case 1: TDR_STENCIL_GO(1); break;
    case 3: TDR_STENCIL_GO(3); break;
    case 5: TDR_STENCIL_GO(5); break;
    default: TDR_STENCIL_GO(7); break;
  }
#undef TDR_STENCIL_GO
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
