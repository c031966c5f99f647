// MASA match-and-transfer glue kernels (reference: network_restormer_guided_arch.py:642-734, 753-900; closed form in
// SURVEY.md appendix A).  The two correlation searches themselves run on tensor cores through tdr_conv_gemm (the
// normalised lq descriptors are laid out as per-sample 3x3 filters here); these kernels build the filters and the
// 1/|v| row scales, take the arg-max, place the windows and do the final gather-average ("transfer") without ever
// materialising an unfold/fold buffer (the reference's x8 level im2col is 1.2 GB per sample).
#include "tdr_common.cuh"

namespace {

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = blockDim.x >> 5;
  float t = lane < nw ? red[lane] : 0.f;
  t = warp_sum(t);
  __syncthreads();
  return t;
}

__device__ __forceinline__ void load8(const float* p, float* f) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// v = hi + lo with hi = bf16(v), lo = bf16(v - hi): 16 significand bits in two bf16 numbers (lo keeps bf16's full exponent
// range, so there is no underflow issue as with an fp16 pair)
__device__ __forceinline__ void split8(const float* v, float* hi, float* lo) {
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    hi[e] = __bfloat162float(__float2bfloat16(v[e]));
    lo[e] = v[e] - hi[e];
  }
}

__global__ void sqnorm_rows_kernel(const float* __restrict__ x, long long ld, long long rows, int C, float* n2) {
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float s = 0.f;
  for (int v = lane; v < (C >> 3); v += 32) {
    float f[8];
    load8(x + row * ld + v * 8, f);
#pragma unroll
    for (int e = 0; e < 8; ++e) s = fmaf(f[e], f[e], s);
  }
  s = warp_sum(s);
  if (lane == 0) n2[row] = s;
}

// out[r] = [hi | lo | hi] (3C bf16): the A operand of the split-bf16 correlations; the filters are laid out [hi | hi | lo],
// so one bf16 GEMM over 3C channels accumulates hi*hi + lo*hi + hi*lo in fp32 (the lo*lo term, 2^-16 relative, is dropped)
__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ x, long long ld, long long rows, int C,
                                                     bf16* __restrict__ out, long long out_ld) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i % nvec) * 8;
    float f[8], hi[8], lo[8];
    load8(x + r * ld + c, f);
    split8(f, hi, lo);
    const bf16x8 h8 = pack8(hi);
    bf16* o = out + r * out_ld + c;
    *reinterpret_cast<bf16x8*>(o) = h8;
    *reinterpret_cast<bf16x8*>(o + C) = pack8(lo);
    *reinterpret_cast<bf16x8*>(o + 2 * C) = h8;
  }
}

__global__ void ref_invnorm_kernel(const float* __restrict__ n2, int B, int H, int W, int d0, int d1, int d2, int ndil,
                                   float* __restrict__ inv) {
  const long long per = (long long)B * H * W;
  const long long total = per * ndil;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int di = (int)(i / per);
    const long long p = i % per;
    const int x = (int)(p % W), y = (int)((p / W) % H), b = (int)(p / ((long long)W * H));
    const int d = di == 0 ? d0 : (di == 1 ? d1 : d2);
    float s = 0.f;
    for (int ty = -1; ty <= 1; ++ty)
      for (int tx = -1; tx <= 1; ++tx) {
        const int yy = y + ty * d, xx = x + tx * d;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) s += n2[((long long)b * H + yy) * W + xx];
      }
    inv[i] = 1.f / fmaxf(sqrtf(s), 1e-12f);
  }
}

// grid (co_pad, B, ndil), block 128
__global__ void __launch_bounds__(128) coarse_filters_kernel(const float* __restrict__ f, int B, int H, int W, int C,
                                                             int k_y, int k_x, int d0, int d1, int d2, int co_pad,
                                                             bf16* __restrict__ w) {
  __shared__ float red[4];
  const int blk = blockIdx.x, b = blockIdx.y, di = blockIdx.z;
  const int px = W / k_x, py = H / k_y;
  const int nvec = C >> 3;
  const size_t C3 = 3 * (size_t)C;                                               // [hi | hi | lo] per filter row
  bf16* wout = w + ((((size_t)di * B + b) * 9) * co_pad + blk) * C3;             // + tap * co_pad * 3C
  if (blk >= py * px) {
    for (int i = threadIdx.x; i < 9 * 3 * nvec; i += blockDim.x) {
      const int tap = i / (3 * nvec), v = i % (3 * nvec);
      *reinterpret_cast<uint4*>(wout + (size_t)tap * co_pad * C3 + v * 8) = make_uint4(0, 0, 0, 0);
    }
    return;
  }
  const int d = di == 0 ? d0 : (di == 1 ? d1 : d2);
  const int by = blk / px, bx = blk % px;
  const int cy = (k_y + 2) / 2, cx = (k_x + 2) / 2;        // centre of the (k+2)^2 haloed tile (R:680)
  float s = 0.f;
  for (int i = threadIdx.x; i < 9 * nvec; i += blockDim.x) {
    const int tap = i / nvec, v = i % nvec;
    const int yy = clampi(by * k_y - 1 + cy + (tap / 3 - 1) * d, 0, H - 1);
    const int xx = clampi(bx * k_x - 1 + cx + (tap % 3 - 1) * d, 0, W - 1);
    float e[8];
    load8(f + (((size_t)b * H + yy) * W + xx) * C + v * 8, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(e[j], e[j], s);
  }
  s = block_sum(s, red);
  const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
  for (int i = threadIdx.x; i < 9 * nvec; i += blockDim.x) {
    const int tap = i / nvec, v = i % nvec;
    const int yy = clampi(by * k_y - 1 + cy + (tap / 3 - 1) * d, 0, H - 1);
    const int xx = clampi(bx * k_x - 1 + cx + (tap % 3 - 1) * d, 0, W - 1);
    float e[8], hi[8], lo[8];
    load8(f + (((size_t)b * H + yy) * W + xx) * C + v * 8, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] *= inv;
    split8(e, hi, lo);
    bf16* o = wout + (size_t)tap * co_pad * C3 + v * 8;
    const bf16x8 h8 = pack8(hi);
    *reinterpret_cast<bf16x8*>(o) = h8;
    *reinterpret_cast<bf16x8*>(o + C) = h8;
    *reinterpret_cast<bf16x8*>(o + 2 * C) = pack8(lo);
  }
}

// grid (ceil(nblk/32), B), block (32, 8): lanes over blocks (coalesced), 8 position slices reduced in smem.
__global__ void __launch_bounds__(256) coarse_argmax_kernel(const float* __restrict__ score, int Hr, int Wr, int nblk,
                                                            int co_pad, int d_y, int d_x, int* __restrict__ idx_out,
                                                            int* __restrict__ origin) {
  __shared__ float sv[8][32];
  __shared__ int sp[8][32];
  const int bl = threadIdx.x, ps = threadIdx.y;
  const int blk = blockIdx.x * 32 + bl, b = blockIdx.y;
  const int npos = Hr * Wr;
  float best = -INFINITY;
  int bpos = 0;
  if (blk < nblk) {
    for (int p = ps; p < npos; p += 8) {
      const float v = score[((size_t)b * npos + p) * co_pad + blk];
      if (v > best) { best = v; bpos = p; }
    }
  }
  sv[ps][bl] = best;
  sp[ps][bl] = bpos;
  __syncthreads();
  if (ps == 0 && blk < nblk) {
    for (int s = 1; s < 8; ++s) {
      const float v = sv[s][bl];
      const int p = sp[s][bl];
      if (v > best || (v == best && p < bpos)) { best = v; bpos = p; }
    }
    const int ix = bpos % Wr, iy = bpos / Wr;
    // window placement R:793-815
    int x1 = ix - d_x / 2 - 1, x2 = ix + d_x / 2 + 1;
    if (x1 < 0) { x1 = 0; x2 = d_x + 1; }
    if (x2 > Wr - 1) { x2 = Wr - 1; x1 = x2 - (d_x + 1); }
    int y1 = iy - d_y / 2 - 1, y2 = iy + d_y / 2 + 1;
    if (y1 < 0) { y1 = 0; y2 = d_y + 1; }
    if (y2 > Hr - 1) { y2 = Hr - 1; y1 = y2 - (d_y + 1); }
    const int o = b * nblk + blk;
    if (idx_out) idx_out[o] = bpos;
    origin[3 * o] = b;
    origin[3 * o + 1] = y1;
    origin[3 * o + 2] = x1;
  }
}

// grid (nq, nwin), block 128:  w[win][tap][q][C]
__global__ void __launch_bounds__(128) fine_filters_kernel(const float* __restrict__ f, int H, int W, int C, int k_y,
                                                           int k_x, bf16* __restrict__ w) {
  __shared__ float red[4];
  const int q = blockIdx.x, win = blockIdx.y;
  const int px = W / k_x, py = H / k_y;
  const int nblk = py * px, nq = k_y * k_x;
  const int b = win / nblk, blk = win % nblk;
  const int by = blk / px, bx = blk % px;
  const int qy = q / k_x, qx = q % k_x;
  const int nvec = C >> 3;
  float s = 0.f;
  for (int i = threadIdx.x; i < 9 * nvec; i += blockDim.x) {
    const int tap = i / nvec, v = i % nvec;
    const int yy = clampi(by * k_y - 1 + qy + tap / 3, 0, H - 1);
    const int xx = clampi(bx * k_x - 1 + qx + tap % 3, 0, W - 1);
    float e[8];
    load8(f + (((size_t)b * H + yy) * W + xx) * C + v * 8, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(e[j], e[j], s);
  }
  s = block_sum(s, red);
  const float inv = 1.f / fmaxf(sqrtf(s), 1e-12f);
  for (int i = threadIdx.x; i < 9 * nvec; i += blockDim.x) {
    const int tap = i / nvec, v = i % nvec;
    const int yy = clampi(by * k_y - 1 + qy + tap / 3, 0, H - 1);
    const int xx = clampi(bx * k_x - 1 + qx + tap % 3, 0, W - 1);
    float e[8], hi[8], lo[8];
    load8(f + (((size_t)b * H + yy) * W + xx) * C + v * 8, e);
#pragma unroll
    for (int j = 0; j < 8; ++j) e[j] *= inv;
    split8(e, hi, lo);
    bf16* o = w + (((size_t)win * 9 + tap) * nq + q) * 3 * (size_t)C + v * 8;       // [hi | hi | lo]
    const bf16x8 h8 = pack8(hi);
    *reinterpret_cast<bf16x8*>(o) = h8;
    *reinterpret_cast<bf16x8*>(o + C) = h8;
    *reinterpret_cast<bf16x8*>(o + 2 * C) = pack8(lo);
  }
}

__global__ void win_invnorm_kernel(const float* __restrict__ n2, int Hr, int Wr, const int* __restrict__ origin,
                                   int nwin, int d_y, int d_x, float* __restrict__ inv) {
  const long long total = (long long)nwin * d_y * d_x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int jx = (int)(i % d_x), jy = (int)((i / d_x) % d_y), win = (int)(i / ((long long)d_x * d_y));
    const int b = origin[3 * win], y1 = origin[3 * win + 1], x1 = origin[3 * win + 2];
    float s = 0.f;
    for (int ty = 0; ty < 3; ++ty)
      for (int tx = 0; tx < 3; ++tx) {
        const int yy = y1 + jy + ty, xx = x1 + jx + tx;
        if (yy >= 0 && yy < Hr && xx >= 0 && xx < Wr) s += n2[((long long)b * Hr + yy) * Wr + xx];
      }
    inv[i] = 1.f / fmaxf(sqrtf(s), 1e-12f);
  }
}

// grid (nwin), block nq threads (rounded up to 32)
__global__ void fine_argmax_kernel(const float* __restrict__ corr, int npos, int nq, int* __restrict__ index,
                                   float* __restrict__ att) {
  const int win = blockIdx.x;
  for (int q = threadIdx.x; q < nq; q += blockDim.x) {
    float best = -INFINITY;
    int bp = 0;
    for (int p = 0; p < npos; ++p) {
      const float v = corr[((size_t)win * npos + p) * nq + q];
      if (v > best) { best = v; bp = p; }
    }
    index[(size_t)win * nq + q] = bp;
    att[(size_t)win * nq + q] = best;
  }
}

// Every quotient below is a multiply-high by a host-prepared constant (tdr_fast_div_setup): the kernel used to spend
// most of its issue slots in ~30 runtime integer divisions per thread for 9 gathered 16-byte loads.
struct TransferDiv {
  uint32_t nvec_m, nvec_s, ow_m, ow_s, oh_m, oh_s, wy_m, wy_s, wx_m, wx_s, s_m, s_s, dx_m, dx_s;
};
__device__ __forceinline__ int tr_fdiv(int n, uint32_t m, uint32_t s) { return m ? (int)(__umulhi((uint32_t)n, m) >> s) : n; }

__global__ void __launch_bounds__(256) transfer_kernel(const bf16* __restrict__ f, int B, int Hs, int Ws, int C,
                                                       const int* __restrict__ origin, const int* __restrict__ index,
                                                       const float* __restrict__ att, int py, int px, int k_y, int k_x,
                                                       int d_x, int s, float* __restrict__ o32, long long ld32,
                                                       bf16* __restrict__ o16, long long ld16, const TransferDiv fd) {
  const int OH = py * k_y * s, OW = px * k_x * s;
  const int nvec = C >> 3;
  const int nq = k_y * k_x, nblk = py * px;
  const int total = B * OH * OW * nvec;                   // < 2^31 (checked by the caller)
  const float inv_s = 1.f / (float)s;
  for (long long it64 = blockIdx.x * (long long)blockDim.x + threadIdx.x; it64 < total;
       it64 += (long long)gridDim.x * blockDim.x) {
    const int it = (int)it64;
    const int p = tr_fdiv(it, fd.nvec_m, fd.nvec_s);
    const int v = it - p * nvec;
    const int t = tr_fdiv(p, fd.ow_m, fd.ow_s);
    const int X = p - t * OW;
    const int b = tr_fdiv(t, fd.oh_m, fd.oh_s);
    const int Y = t - b * OH;
    const int tby = tr_fdiv(Y, fd.wy_m, fd.wy_s), Yl = Y - tby * (k_y * s);
    const int tbx = tr_fdiv(X, fd.wx_m, fd.wx_s), Xl = X - tbx * (k_x * s);
    const int qy0 = tr_fdiv(Yl, fd.s_m, fd.s_s), qx0 = tr_fdiv(Xl, fd.s_m, fd.s_s);
    const int win = b * nblk + tby * px + tbx;
    const int y1 = origin[3 * win + 1] * s, x1 = origin[3 * win + 2] * s;
    const int* idx = index + (size_t)win * nq;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    int cnt = 0;
#pragma unroll
    for (int oy = -1; oy <= 1; ++oy) {
      const int qy = qy0 + oy;
      if (qy < 0 || qy >= k_y) continue;
#pragma unroll
      for (int ox = -1; ox <= 1; ++ox) {
        const int qx = qx0 + ox;
        if (qx < 0 || qx >= k_x) continue;
        const int j = idx[qy * k_x + qx];
        const int jy = tr_fdiv(j, fd.dx_m, fd.dx_s);
        const int sy = y1 + jy * s + (Yl - qy * s + s);
        const int sx = x1 + (j - jy * d_x) * s + (Xl - qx * s + s);
        ++cnt;
        if (sy >= 0 && sy < Hs && sx >= 0 && sx < Ws) {
          float e8[8];
          unpack8(*reinterpret_cast<const bf16x8*>(f + (((size_t)b * Hs + sy) * Ws + sx) * C + v * 8), e8);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += e8[e];
        }
      }
    }
    // bilinear (align_corners=False) up-sampling of the per-window confidence map R:712
    float fy = ((float)Yl + 0.5f) * inv_s - 0.5f, fx = ((float)Xl + 0.5f) * inv_s - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1i = y0 + 1 < k_y ? y0 + 1 : k_y - 1, x1i = x0 + 1 < k_x ? x0 + 1 : k_x - 1;
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* aw = att + (size_t)win * nq;
    const float a = (1.f - ly) * ((1.f - lx) * aw[y0 * k_x + x0] + lx * aw[y0 * k_x + x1i]) +
                    ly * ((1.f - lx) * aw[y1i * k_x + x0] + lx * aw[y1i * k_x + x1i]);
    const float scale = a / (float)cnt;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= scale;
    if (o32) {
      float4* q4 = reinterpret_cast<float4*>(o32 + (long long)p * ld32 + v * 8);
      q4[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      q4[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    if (o16) *reinterpret_cast<bf16x8*>(o16 + (long long)p * ld16 + v * 8) = pack8(acc);
  }
}


// ------------------------------------------------------------------------------------------------ backward
__device__ __forceinline__ void atomic_add8(float* dst, const float* v) {
  atomicAdd(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));       // red.global.add.v4.f32
  atomicAdd(reinterpret_cast<float4*>(dst + 4), make_float4(v[4], v[5], v[6], v[7]));
}

// Backward of transfer_kernel.  G lanes (power of two) cooperate on one output pixel, each lane owning 8-channel
// vectors v = l, l+G, ...:  out = mean_gathered(f) * a(p)  =>  dref[gathered] += dout * a / cnt  (vector fp32 atomics),
// da = sum_c dout * mean_gathered(f), distributed to the 4 bilinear taps of the per-window confidence map (atomics).
__global__ void __launch_bounds__(256) transfer_bwd_kernel(const float* __restrict__ dout, long long dld,
                                                           const bf16* __restrict__ f, int B, int Hs, int Ws, int C,
                                                           const int* __restrict__ origin, const int* __restrict__ index,
                                                           const float* __restrict__ att, int py, int px, int k_y,
                                                           int k_x, int d_x, int s, float* __restrict__ dref,
                                                           float* __restrict__ datt, int G) {
  const int OH = py * k_y * s, OW = px * k_x * s;
  const int nvec = C >> 3;
  const int nq = k_y * k_x, nblk = py * px;
  const long long npix = (long long)B * OH * OW;
  const float inv_s = 1.f / (float)s;
  const int ppb = 256 / G;                              // pixels per block iteration
  const int slot = threadIdx.x / G, l = threadIdx.x % G;
  for (long long p0 = (long long)blockIdx.x * ppb; p0 < npix; p0 += (long long)gridDim.x * ppb) {
    const long long p = p0 + slot;
    const bool ok = p < npix;
    float da = 0.f;
    int win = 0, y0 = 0, x0 = 0, y1i = 0, x1i = 0;
    float ly = 0.f, lx = 0.f;
    if (ok) {
      const int X = (int)(p % OW), Y = (int)((p / OW) % OH), b = (int)(p / ((long long)OW * OH));
      const int tby = Y / (k_y * s), Yl = Y % (k_y * s);
      const int tbx = X / (k_x * s), Xl = X % (k_x * s);
      win = b * nblk + tby * px + tbx;
      const int y1 = origin[3 * win + 1] * s, x1 = origin[3 * win + 2] * s;
      const int* idx = index + (size_t)win * nq;
      int sy[9], sx[9];
      int cnt = 0, nval = 0;
#pragma unroll
      for (int oy = -1; oy <= 1; ++oy) {
        const int qy = Yl / s + oy;
        if (qy < 0 || qy >= k_y) continue;
#pragma unroll
        for (int ox = -1; ox <= 1; ++ox) {
          const int qx = Xl / s + ox;
          if (qx < 0 || qx >= k_x) continue;
          const int j = idx[qy * k_x + qx];
          const int yy = y1 + (j / d_x) * s + (Yl - qy * s + s);
          const int xx = x1 + (j % d_x) * s + (Xl - qx * s + s);
          ++cnt;
          if (yy >= 0 && yy < Hs && xx >= 0 && xx < Ws) { sy[nval] = yy; sx[nval] = xx; ++nval; }
        }
      }
      float fy = ((float)Yl + 0.5f) * inv_s - 0.5f, fx = ((float)Xl + 0.5f) * inv_s - 0.5f;
      fy = fy < 0.f ? 0.f : fy;
      fx = fx < 0.f ? 0.f : fx;
      y0 = (int)fy; x0 = (int)fx;
      y1i = y0 + 1 < k_y ? y0 + 1 : k_y - 1; x1i = x0 + 1 < k_x ? x0 + 1 : k_x - 1;
      ly = fy - (float)y0; lx = fx - (float)x0;
      const float* aw = att + (size_t)win * nq;
      const float a = (1.f - ly) * ((1.f - lx) * aw[y0 * k_x + x0] + lx * aw[y0 * k_x + x1i]) +
                      ly * ((1.f - lx) * aw[y1i * k_x + x0] + lx * aw[y1i * k_x + x1i]);
      const float inv_cnt = 1.f / (float)cnt;
      for (int v = l; v < nvec; v += G) {
        const float4* g4 = reinterpret_cast<const float4*>(dout + p * dld + v * 8);
        const float4 ga = g4[0], gb = g4[1];
        const float g[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        float acc[8], sc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { acc[e] = 0.f; sc[e] = g[e] * a * inv_cnt; }
        for (int k = 0; k < nval; ++k) {
          const size_t off = (((size_t)b * Hs + sy[k]) * Ws + sx[k]) * C + v * 8;
          float e8[8];
          unpack8(*reinterpret_cast<const bf16x8*>(f + off), e8);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] += e8[e];
          atomic_add8(dref + off, sc);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) da = fmaf(g[e], acc[e] * inv_cnt, da);
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) da += __shfl_xor_sync(0xffffffffu, da, o);      // G-lane groups are aligned
    if (ok && l == 0) {
      float* dw = datt + (size_t)win * nq;
      atomicAdd(dw + y0 * k_x + x0, (1.f - ly) * (1.f - lx) * da);
      atomicAdd(dw + y0 * k_x + x1i, (1.f - ly) * lx * da);
      atomicAdd(dw + y1i * k_x + x0, ly * (1.f - lx) * da);
      atomicAdd(dw + y1i * k_x + x1i, ly * lx * da);
    }
  }
}

// Backward of the fine-search confidence R:661-670: att[win, q] = cos(a, r), a = 3x3xC lq patch at interior position q
// (replicate-padded halo), r = 3x3xC ref-window patch at the arg-max position.  grid (nq, nwin), block 128.
//   da = datt / |a| * (r/|r| - cos * a/|a|),   dr = datt / |r| * (a/|a| - cos * r/|r|)   (scattered with fp32 atomics)
__global__ void __launch_bounds__(128) fine_bwd_kernel(const bf16* __restrict__ flq, int H, int W,
                                                       const bf16* __restrict__ fref, int Hr, int Wr, int C, int k_y,
                                                       int k_x, int d_x, const int* __restrict__ origin,
                                                       const int* __restrict__ index, const float* __restrict__ datt,
                                                       float* __restrict__ dlq, float* __restrict__ dref) {
  __shared__ float red[4];
  const int q = blockIdx.x, win = blockIdx.y;
  const int px = W / k_x, py = H / k_y;
  const int nblk = py * px, nq = k_y * k_x;
  const float g = datt[(size_t)win * nq + q];
  if (g == 0.f) return;
  const int b = win / nblk, blk = win % nblk;
  const int by = blk / px, bx = blk % px;
  const int qy = q / k_x, qx = q % k_x;
  const int j = index[(size_t)win * nq + q];
  const int ry = origin[3 * win + 1] + j / d_x, rx = origin[3 * win + 2] + j % d_x;
  const int nvec = C >> 3;
  float saa = 0.f, srr = 0.f, sar = 0.f;
  for (int i = threadIdx.x; i < 9 * nvec; i += blockDim.x) {
    const int tap = i / nvec, v = i % nvec;
    const int yy = clampi(by * k_y - 1 + qy + tap / 3, 0, H - 1);
    const int xx = clampi(bx * k_x - 1 + qx + tap % 3, 0, W - 1);
    float a8[8], r8[8];
    unpack8(*reinterpret_cast<const bf16x8*>(flq + (((size_t)b * H + yy) * W + xx) * C + v * 8), a8);
    unpack8(*reinterpret_cast<const bf16x8*>(fref + (((size_t)b * Hr + ry + tap / 3) * Wr + rx + tap % 3) * C + v * 8), r8);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      saa = fmaf(a8[e], a8[e], saa);
      srr = fmaf(r8[e], r8[e], srr);
      sar = fmaf(a8[e], r8[e], sar);
    }
  }
  saa = block_sum(saa, red);
  srr = block_sum(srr, red);
  sar = block_sum(sar, red);
  const float na = fmaxf(sqrtf(saa), 1e-12f), nr = fmaxf(sqrtf(srr), 1e-12f);
  const float cosv = sar / (na * nr);
  const float ca_r = g / (na * nr), ca_a = -g * cosv / (na * na);      // da = ca_r * r + ca_a * a
  const float cr_a = g / (na * nr), cr_r = -g * cosv / (nr * nr);      // dr = cr_a * a + cr_r * r
  for (int i = threadIdx.x; i < 9 * nvec; i += blockDim.x) {
    const int tap = i / nvec, v = i % nvec;
    const int yy = clampi(by * k_y - 1 + qy + tap / 3, 0, H - 1);
    const int xx = clampi(bx * k_x - 1 + qx + tap % 3, 0, W - 1);
    const size_t oa = (((size_t)b * H + yy) * W + xx) * C + v * 8;
    const size_t orr = (((size_t)b * Hr + ry + tap / 3) * Wr + rx + tap % 3) * C + v * 8;
    float a8[8], r8[8], da[8], dr[8];
    unpack8(*reinterpret_cast<const bf16x8*>(flq + oa), a8);
    unpack8(*reinterpret_cast<const bf16x8*>(fref + orr), r8);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      da[e] = ca_r * r8[e] + ca_a * a8[e];
      dr[e] = cr_a * a8[e] + cr_r * r8[e];
    }
    atomic_add8(dlq + oa, da);
    atomic_add8(dref + orr, dr);
  }
}

// zero insertion (adjoint of a stride-2 subsampling): out[b, 2y, 2x, c] = in[b, y, x, c]; out is pre-zeroed
__global__ void __launch_bounds__(256) dilate2_kernel(const bf16* __restrict__ in, long long in_ld, int B, int H, int W,
                                                      int C, bf16* __restrict__ out, long long out_ld, int OH, int OW) {
  const int nvec = C >> 3;
  const long long total = (long long)B * H * W * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % nvec);
    long long p = i / nvec;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    if (2 * y < OH && 2 * x < OW)
      *reinterpret_cast<uint4*>(out + ((((long long)b * OH + 2 * y) * OW) + 2 * x) * out_ld + v * 8) =
          *reinterpret_cast<const uint4*>(in + ((((long long)b * H + y) * W) + x) * in_ld + v * 8);
  }
}

inline int grid1d(long long items, int per_block) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = (long long)tdr_num_sms() * 16;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

extern "C" int tdr_sqnorm_rows(const float* x, long long ld, long long rows, int C, float* n2, cudaStream_t stream) {
  TDR_CHECK_ARG(x && n2 && C % 8 == 0 && ld % 4 == 0 && rows > 0 && ((uintptr_t)x & 15) == 0, "tdr_sqnorm_rows: bad arguments");
  const long long blocks = (rows * 32 + 255) / 256;
  sqnorm_rows_kernel<<<(int)blocks, 256, 0, stream>>>(x, ld, rows, C, n2);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_split3(const float* x, long long ld, long long rows, int C, void* out_bf16, long long out_ld,
                               cudaStream_t stream) {
  TDR_CHECK_ARG(x && out_bf16 && C % 8 == 0 && ld % 4 == 0 && out_ld % 8 == 0 && out_ld >= 3LL * C && rows > 0 &&
                ((uintptr_t)x & 15) == 0 && ((uintptr_t)out_bf16 & 15) == 0, "tdr_masa_split3: bad arguments");
  split3_kernel<<<grid1d(rows * (C / 8), 256), 256, 0, stream>>>(x, ld, rows, C, reinterpret_cast<bf16*>(out_bf16), out_ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_ref_invnorm(const float* n2, int B, int H, int W, const int* host_dils, int ndil, float* inv,
                                    cudaStream_t stream) {
  TDR_CHECK_ARG(n2 && inv && host_dils && ndil >= 1 && ndil <= 3, "tdr_masa_ref_invnorm: bad arguments (ndil 1..3)");
  const int d0 = host_dils[0], d1 = ndil > 1 ? host_dils[1] : 1, d2 = ndil > 2 ? host_dils[2] : 1;
  ref_invnorm_kernel<<<grid1d((long long)B * H * W * ndil, 256), 256, 0, stream>>>(n2, B, H, W, d0, d1, d2, ndil, inv);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_coarse_filters(const float* f_lq, int B, int H, int W, int C, int k_y, int k_x,
                                       const int* host_dils, int ndil, int co_pad, void* w_bf16, cudaStream_t stream) {
  TDR_CHECK_ARG(f_lq && w_bf16 && host_dils && ndil >= 1 && ndil <= 3, "tdr_masa_coarse_filters: bad arguments");
  TDR_CHECK_ARG(C % 8 == 0 && H % k_y == 0 && W % k_x == 0, "tdr_masa_coarse_filters: bad geometry");
  TDR_CHECK_ARG(co_pad % 8 == 0 && co_pad >= (H / k_y) * (W / k_x), "tdr_masa_coarse_filters: bad co_pad");
  const int d0 = host_dils[0], d1 = ndil > 1 ? host_dils[1] : 1, d2 = ndil > 2 ? host_dils[2] : 1;
  dim3 grid(co_pad, B, ndil);
  coarse_filters_kernel<<<grid, 128, 0, stream>>>(f_lq, B, H, W, C, k_y, k_x, d0, d1, d2, co_pad,
                                                  reinterpret_cast<bf16*>(w_bf16));
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_coarse_argmax(const float* score, int B, int Hr, int Wr, int nblk, int co_pad, int d_y, int d_x,
                                      int* idx_out, int* origin, cudaStream_t stream) {
  TDR_CHECK_ARG(score && origin && B > 0 && nblk > 0 && co_pad >= nblk, "tdr_masa_coarse_argmax: bad arguments");
  TDR_CHECK_ARG(Wr >= d_x + 2 && Hr >= d_y + 2, "tdr_masa_coarse_argmax: reference feature map (%dx%d) smaller than "
                "the %dx%d search window", Hr, Wr, d_y + 2, d_x + 2);
  dim3 grid((nblk + 31) / 32, B), block(32, 8);
  coarse_argmax_kernel<<<grid, block, 0, stream>>>(score, Hr, Wr, nblk, co_pad, d_y, d_x, idx_out, origin);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_fine_filters(const float* f_lq, int B, int H, int W, int C, int k_y, int k_x, void* w_bf16,
                                     cudaStream_t stream) {
  TDR_CHECK_ARG(f_lq && w_bf16 && C % 8 == 0 && H % k_y == 0 && W % k_x == 0, "tdr_masa_fine_filters: bad arguments");
  dim3 grid(k_y * k_x, B * (H / k_y) * (W / k_x));
  fine_filters_kernel<<<grid, 128, 0, stream>>>(f_lq, H, W, C, k_y, k_x, reinterpret_cast<bf16*>(w_bf16));
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_win_invnorm(const float* n2_ref, int Hr, int Wr, const int* origin, int nwin, int d_y, int d_x,
                                    float* inv, cudaStream_t stream) {
  TDR_CHECK_ARG(n2_ref && origin && inv && nwin > 0, "tdr_masa_win_invnorm: bad arguments");
  win_invnorm_kernel<<<grid1d((long long)nwin * d_y * d_x, 256), 256, 0, stream>>>(n2_ref, Hr, Wr, origin, nwin, d_y,
                                                                                   d_x, inv);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_fine_argmax(const float* corr, int nwin, int npos, int nq, int* index, float* att,
                                    cudaStream_t stream) {
  TDR_CHECK_ARG(corr && index && att && nwin > 0 && npos > 0 && nq > 0, "tdr_masa_fine_argmax: bad arguments");
  fine_argmax_kernel<<<nwin, ((nq + 31) / 32) * 32, 0, stream>>>(corr, npos, nq, index, att);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_transfer(const void* f_ref_bf16, int B, int Hr_s, int Wr_s, int C, const int* origin,
                                 const int* index, const float* att, int py, int px, int k_y, int k_x, int d_x, int s,
                                 float* out, long long out_ld, void* out_bf16, long long out_bf16_ld,
                                 cudaStream_t stream) {
  TDR_CHECK_ARG(f_ref_bf16 && origin && index && att && (out || out_bf16), "tdr_masa_transfer: null pointer");
  TDR_CHECK_ARG(C % 8 == 0 && s >= 1 && out_ld % 4 == 0 && out_bf16_ld % 8 == 0, "tdr_masa_transfer: bad arguments");
  const long long items = (long long)B * py * k_y * s * px * k_x * s * (C / 8);
  TDR_CHECK_ARG(items < (1ll << 31) && k_y >= 1 && k_x >= 1 && d_x >= 1 && py >= 1 && px >= 1,
                "tdr_masa_transfer: %lld output vectors (limit 2^31 per call: split the batch)", items);
  TransferDiv fd;
  tdr_fast_div_setup(C / 8, &fd.nvec_m, &fd.nvec_s);
  tdr_fast_div_setup(px * k_x * s, &fd.ow_m, &fd.ow_s);
  tdr_fast_div_setup(py * k_y * s, &fd.oh_m, &fd.oh_s);
  tdr_fast_div_setup(k_y * s, &fd.wy_m, &fd.wy_s);
  tdr_fast_div_setup(k_x * s, &fd.wx_m, &fd.wx_s);
  tdr_fast_div_setup(s, &fd.s_m, &fd.s_s);
  tdr_fast_div_setup(d_x, &fd.dx_m, &fd.dx_s);
  transfer_kernel<<<grid1d(items, 256), 256, 0, stream>>>(reinterpret_cast<const bf16*>(f_ref_bf16), B, Hr_s, Wr_s, C,
                                                          origin, index, att, py, px, k_y, k_x, d_x, s, out, out_ld,
                                                          reinterpret_cast<bf16*>(out_bf16), out_bf16_ld, fd);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_transfer_bwd(const float* dout, long long dout_ld, const void* f_ref_bf16, int B, int Hr_s,
                                     int Wr_s, int C, const int* origin, const int* index, const float* att, int py,
                                     int px, int k_y, int k_x, int d_x, int s, float* dref, float* datt,
                                     cudaStream_t stream) {
  TDR_CHECK_ARG(dout && f_ref_bf16 && origin && index && att && dref && datt, "tdr_masa_transfer_bwd: null pointer");
  TDR_CHECK_ARG(C % 8 == 0 && s >= 1 && dout_ld % 4 == 0 && ((uintptr_t)dref & 15) == 0 && ((uintptr_t)dout & 15) == 0,
                "tdr_masa_transfer_bwd: bad arguments");
  const int nvec = C / 8;
  int G = 1;
  while (G < 32 && G < nvec) G <<= 1;
  const long long npix = (long long)B * py * k_y * s * px * k_x * s;
  transfer_bwd_kernel<<<grid1d(npix, 256 / G), 256, 0, stream>>>(dout, dout_ld, reinterpret_cast<const bf16*>(f_ref_bf16),
                                                                 B, Hr_s, Wr_s, C, origin, index, att, py, px, k_y, k_x,
                                                                 d_x, s, dref, datt, G);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_fine_bwd(const void* f_lq_bf16, int B, int H, int W, const void* f_ref_bf16, int Hr, int Wr, int C,
                                 int k_y, int k_x, int d_x, const int* origin, const int* index, const float* datt,
                                 float* dlq, float* dref, cudaStream_t stream) {
  TDR_CHECK_ARG(f_lq_bf16 && f_ref_bf16 && origin && index && datt && dlq && dref, "tdr_masa_fine_bwd: null pointer");
  TDR_CHECK_ARG(C % 8 == 0 && H % k_y == 0 && W % k_x == 0 && B > 0, "tdr_masa_fine_bwd: bad geometry");
  dim3 grid(k_y * k_x, B * (H / k_y) * (W / k_x));
  fine_bwd_kernel<<<grid, 128, 0, stream>>>(reinterpret_cast<const bf16*>(f_lq_bf16), H, W,
                                            reinterpret_cast<const bf16*>(f_ref_bf16), Hr, Wr, C, k_y, k_x, d_x, origin,
                                            index, datt, dlq, dref);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_dilate2_nhwc(const void* in_bf16, long long in_ld, int B, int H, int W, int C, void* out_bf16,
                                long long out_ld, int OH, int OW, cudaStream_t stream) {
  TDR_CHECK_ARG(in_bf16 && out_bf16 && B > 0 && H > 0 && W > 0 && C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0,
                "tdr_dilate2_nhwc: bad arguments");
  TDR_CHECK_ARG(OH >= 2 * H - 1 && OW >= 2 * W - 1, "tdr_dilate2_nhwc: output too small");
  dilate2_kernel<<<grid1d((long long)B * H * W * (C / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(in_bf16), in_ld, B, H, W, C, reinterpret_cast<bf16*>(out_bf16), out_ld, OH, OW);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
