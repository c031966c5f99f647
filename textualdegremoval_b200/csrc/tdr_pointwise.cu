// HBM-bound row/stencil kernels on NHWC activations: channel LayerNorm / cast, depthwise 3x3 (+ GDFN / SimpleGate
// gating), image-boundary layout changes, slice copies, and the two small-channel direct convolutions at the image
// boundary.  All access is 128-bit vectorised along the channel (innermost) dimension; reductions use warp shuffles.
#include <stdlib.h>

#include "tdr_common.cuh"
#include "tdr_stencil.cuh"

namespace {

// ------------------------------------------------------------------------------------------------ rownorm
// G lanes cooperate on one row (G = power of two <= 32), each lane owns NV float4 (strided by G).
template <int NV>
__global__ void __launch_bounds__(256) rownorm_kernel(const float* __restrict__ in, long long in_ld, long long rows,
                                                      int C, int mode, const float* __restrict__ w,
                                                      const float* __restrict__ bvec, float eps, int act,
                                                      bf16* __restrict__ out, long long out_ld,
                                                      float* __restrict__ out32, long long out32_ld, int G) {
  const int lane = threadIdx.x & 31;
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int rpw = 32 / G;
  const int sub = lane / G, l = lane % G;
  const long long row = warp_id * rpw + sub;
  const int nvec = C >> 2;
  const bool row_ok = row < rows;
  pdl_wait();
  pdl_launch();
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = l + i * G;
    if (row_ok && idx < nvec) {
      v[i] = *reinterpret_cast<const float4*>(in + row * in_ld + idx * 4);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    s += v[i].x + v[i].y + v[i].z + v[i].w;
  }
  float mean = 0.f, rstd = 1.f;
  if (mode != 0) {
    for (int o = G >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    mean = s / (float)C;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = l + i * G;
      if (idx < nvec) {
        const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    for (int o = G >> 1; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    rstd = rsqrtf(q / (float)C + eps);
  }
  if (!row_ok) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = l + i * G;
    if (idx < nvec) {
      float4 x = v[i];
      if (mode == 1) {
        const float4 ww = *reinterpret_cast<const float4*>(w + idx * 4);
        const float4 bb = bvec ? *reinterpret_cast<const float4*>(bvec + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        x.x = (x.x - mean) * rstd * ww.x + bb.x;
        x.y = (x.y - mean) * rstd * ww.y + bb.y;
        x.z = (x.z - mean) * rstd * ww.z + bb.z;
        x.w = (x.w - mean) * rstd * ww.w + bb.w;
      } else if (mode == 2) {
        const float4 ww = *reinterpret_cast<const float4*>(w + idx * 4);
        x.x = x.x * rstd * ww.x;
        x.y = x.y * rstd * ww.y;
        x.z = x.z * rstd * ww.z;
        x.w = x.w * rstd * ww.w;
      }
      if ((act & 15) == 3) {            // LeakyReLU(0.01)
        x.x = x.x > 0.f ? x.x : 0.01f * x.x; x.y = x.y > 0.f ? x.y : 0.01f * x.y;
        x.z = x.z > 0.f ? x.z : 0.01f * x.z; x.w = x.w > 0.f ? x.w : 0.01f * x.w;
      }
      if (out) {
        uint2 pk;
        pk.x = pack2r(x.x, x.y, act & 16);               // act bit 4: the 16-bit output is IEEE fp16
        pk.y = pack2r(x.z, x.w, act & 16);
        *reinterpret_cast<uint2*>(out + row * out_ld + idx * 4) = pk;
      }
      if (out32) *reinterpret_cast<float4*>(out32 + row * out32_ld + idx * 4) = x;
    }
  }
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3
// Column-strip sliding window: a thread owns one x position and VEC channels (8 when ungated, 4 + 4 of the two GDFN
// halves when gated) and walks down R rows.  Each input row is loaded once per thread (3 vectors: x-1, x, x+1) and
// immediately scattered into the three pending output rows it contributes to, so nothing but 3 accumulator rows and the
// 9 x VEC fp32 weights live in registers.  x-neighbours are served by L1 (adjacent threads = adjacent channel groups of
// the same pixel, so every warp load is a contiguous 512 B run); rows are re-used from registers, never re-read.
// Packed fp32x2 arithmetic (FFMA2 on sm_100): halves the FMA issue slots of this issue-bound stencil.
template <int VEC>
struct RawT;
template <>
struct RawT<8> {
  typedef uint4 type;
};
template <>
struct RawT<4> {
  typedef uint2 type;
};

template <int VEC>
__device__ __forceinline__ typename RawT<VEC>::type load_raw(const bf16* p, bool ok) {
  typename RawT<VEC>::type raw;
  if (ok) {
    raw = __ldg(reinterpret_cast<const typename RawT<VEC>::type*>(p));
  } else {
    memset(&raw, 0, sizeof(raw));
  }
  return raw;
}

// NH = number of channel halves handled per thread (1 ungated, 2 gated); VEC channels per half.
template <int NH, int VEC, int R>
__global__ void __launch_bounds__(256, 2) dwconv3x3_strip_kernel(const bf16* __restrict__ in, long long in_ld, int H, int W,
                                                              int C, const float* __restrict__ wt,
                                                              const float* __restrict__ bias, int gate,
                                                              bf16* __restrict__ out, long long out_ld) {
  typedef typename RawT<VEC>::type raw_t;
  constexpr int NP = VEC / 2;                  // fp32 pairs per vector
  const int Cout = NH == 2 ? (C >> 1) : C;
  const int ncg = Cout / VEC;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= W * ncg) return;
  const int cg = item % ncg, x = item / ncg;
  const int y0 = blockIdx.y * R;
  const int b = blockIdx.z;
  const int c0 = cg * VEC;

  f2 w[NH][9][NP];
  f2 bv[NH][NP];
#pragma unroll
  for (int h = 0; h < NH; ++h) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int e = 0; e < NP; e += 2) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(wt + (size_t)t * C + h * Cout + c0 + 2 * e));
        w[h][t][e] = pk2(v.x, v.y);
        w[h][t][e + 1] = pk2(v.z, v.w);
      }
#pragma unroll
    for (int e = 0; e < NP; ++e)
      bv[h][e] = bias ? pk2(bias[h * Cout + c0 + 2 * e], bias[h * Cout + c0 + 2 * e + 1]) : pk2(0.f, 0.f);
  }
  // Three pending output rows live in acc[]; the row loop is fully unrolled (R + 2 static iterations) so the slot
  // rotation and the double-buffered raw rows are pure register renaming -- no MOVs are issued for them.
  f2 acc[3][NH][NP];
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int e = 0; e < NP; ++e) acc[k][h][e] = bv[h][e];

  const bool xl_ok = x > 0, xr_ok = x + 1 < W;
  const long long row_stride = (long long)W * in_ld;
  const bf16* rowp = in + ((long long)b * H * W + x) * in_ld + c0 + (long long)(y0 - 1) * row_stride;   // input row y0-1
  bf16* outp = out + (((long long)b * H + y0) * W + x) * out_ld + c0;                                   // output row y0
  const int r_end = min(y0 + R, H);
  raw_t buf[2][NH][3];
  {
    const bool ok = y0 - 1 >= 0;
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      buf[0][h][0] = load_raw<VEC>(rowp + h * Cout - in_ld, ok && xl_ok);
      buf[0][h][1] = load_raw<VEC>(rowp + h * Cout, ok);
      buf[0][h][2] = load_raw<VEC>(rowp + h * Cout + in_ld, ok && xr_ok);
    }
  }
#pragma unroll
  for (int i = 0; i < R + 2; ++i) {
    const int r = y0 - 1 + i;                   // input row consumed in this iteration
    if (i + 1 < R + 2) {                        // prefetch input row r + 1
      rowp += row_stride;
      const bool ok = r + 1 <= r_end && r + 1 < H;
#pragma unroll
      for (int h = 0; h < NH; ++h) {
        buf[(i + 1) & 1][h][0] = load_raw<VEC>(rowp + h * Cout - in_ld, ok && xl_ok);
        buf[(i + 1) & 1][h][1] = load_raw<VEC>(rowp + h * Cout, ok);
        buf[(i + 1) & 1][h][2] = load_raw<VEC>(rowp + h * Cout + in_ld, ok && xr_ok);
      }
    }
    // input row r is tap ky = 2 - k of output row r - 1 + k, which lives in slot (i + k) % 3
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&buf[i & 1][h][kx]);
#pragma unroll
        for (int e = 0; e < NP; ++e) {
          const f2 v = bf2_to_f2(u[e]);
#pragma unroll
          for (int k = 0; k < 3; ++k)
            acc[(i + k) % 3][h][e] = fma2(w[h][(2 - k) * 3 + kx][e], v, acc[(i + k) % 3][h][e]);
        }
      }
    const int ro = r - 1;                       // output row r - 1 (slot i % 3) is complete
    if (i >= 2 && ro < r_end) {
      raw_t o;
      uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
      for (int e = 0; e < NP; ++e) {
        f2 val = acc[i % 3][0][e];
        if (NH == 2) val = mul2(gate == 1 ? gelu2(val) : val, acc[i % 3][NH - 1][e]);
        float a0, a1;
        upk2(val, a0, a1);
        ou[e] = pack2(a0, a1);
      }
      *reinterpret_cast<raw_t*>(outp) = o;
    }
    if (i >= 2) outp += (long long)W * out_ld;
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int e = 0; e < NP; ++e) acc[i % 3][h][e] = bv[h][e];
  }
}

// ------------------------------------------------------------------------------------------------ depthwise 3x3, TMA
// Persistent CTAs (2 per SM) stream haloed tiles [10 rows x 34 px x 64 ch] (ungated) or 2 x [10 x 34 x 32 ch] (the two
// GDFN halves) through a double-buffered TMA pipeline: the bytes in flight no longer depend on registers or occupancy,
// the conv padding is the TMA out-of-bounds zero fill, and each thread reads its 3-pixel neighbourhood from shared
// memory with conflict-free 128-bit (64-bit) loads while walking down 8 output rows with the same register sliding
// window as the strip kernel.
constexpr int kDwRows = 8, kDwCols = 32;
constexpr int kDwBoxBytes = (kDwRows + 2) * (kDwCols + 2) * 64 * 2;      // 43520
constexpr int kDwStageBytes = 43776;                                      // rounded to 256 B

struct DwArgs {
  int B, H, W, C, Cout;
  int tiles_x, tiles_y, chunks, total_tiles;
  const float* wt;
  const float* bias;
  int gate;
  bf16* out;
  long long out_ld;
  const bf16* dg;          // GATE == 2 (gate backward): gradient of the gated product [B,H,W,Cout]
  long long dg_ld;
  const float* dg_add;     // optional per-sample term [B][Cout] added to dg (SCA pool gradient)
  bf16* y_out;             // GATE == 1, optional (training): also store the pre-gate halves [a | b] (2 * Cout channels)
  long long y_ld;
  int relu;                // GATE == 0: ReLU on the output
  uint32_t fd_img_m, fd_img_s, fd_tx_m, fd_tx_s;   // n / (tiles_y * tiles_x), n / tiles_x as umulhi + shift
};

__device__ __forceinline__ int dw_fdiv(int n, uint32_t m, uint32_t s) { return m ? (int)(__umulhi((uint32_t)n, m) >> s) : n; }
// spatial tile index -> (image, first row, first column)
__device__ __forceinline__ void dw_tile_coords(const DwArgs& a, int sp, int& b, int& y0, int& x0) {
  b = dw_fdiv(sp, a.fd_img_m, a.fd_img_s);
  const int r = sp - b * (a.tiles_y * a.tiles_x);
  const int ty = dw_fdiv(r, a.fd_tx_m, a.fd_tx_s);
  y0 = ty * kDwRows;
  x0 = (r - ty * a.tiles_x) * kDwCols;
}

// GATE: 0 plain, 1 gated forward (a.gate 1 = GELU gate, 2 = SimpleGate), 3 = 1 + the pre-gate tensor is stored too
// (training forward; a separate instantiation so that the inference kernel is untouched), 2 gate BACKWARD: recomputes the two depthwise
// halves (a | b) exactly as the forward does and writes d[a | b] = [dg * b * act'(a) | dg * act(a)] (2 * Cout channels),
// i.e. tdr_dwconv3x3(gate 0) + tdr_gate_bwd without the round trip of the pre-gate tensor through HBM.
template <int GATE, bool HALF>
__global__ void __launch_bounds__(256, 2) dwconv3x3_tma_kernel(const __grid_constant__ TdrTensorMap map, const DwArgs a) {
  constexpr int NH = GATE ? 2 : 1;
  constexpr int VEC = GATE ? 4 : 8;
  constexpr int NP = VEC / 2;
  constexpr int CB = GATE ? 32 : 64;                       // channels per box (= output channels per tile)
  constexpr int PIX_PITCH = CB * 2;                        // bytes per pixel in the staged tile
  constexpr int BOX_BYTES = (kDwRows + 2) * (kDwCols + 2) * PIX_PITCH;
  typedef typename RawT<VEC>::type raw_t;
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) uint64_t full[2];
  const int tid = threadIdx.x;
  const int cgi = tid & 7, xl = tid >> 3;                  // channel vector within the chunk, x within the tile

  if (tid == 0) {
    tma_prefetch_desc(&map);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  pdl_wait();
  pdl_launch();

  // CTA -> (channel chunk, spatial group): the chunk is fixed for the CTA's lifetime (taps stay in registers) and the
  // CTAs running concurrently cover ALL chunks of neighbouring spatial tiles, so whole pixel rows are touched together.
  const int cc = blockIdx.x % a.chunks;
  const int grp = blockIdx.x / a.chunks, ngrp = gridDim.x / a.chunks;
  const int n_spatial = a.B * a.tiles_y * a.tiles_x;
  auto issue = [&](int sp, int stage) {                    // thread 0 only
    int b, y0, x0;
    dw_tile_coords(a, sp, b, y0, x0);
    uint8_t* dst = dsm + stage * kDwStageBytes;
    mbar_expect_tx(&full[stage], NH * BOX_BYTES);
#pragma unroll
    for (int h = 0; h < NH; ++h)
      tma_load_4d(dst + h * BOX_BYTES, &map, &full[stage], h * a.Cout + cc * CB, x0 - 1, y0 - 1, b);
  };

  f2 w[NH][9][NP];
  f2 bv[NH][NP];
  const int c0 = cc * CB + cgi * VEC;                      // channel within a half
  const bool c_ok = c0 < a.Cout;
  {
#pragma unroll
      for (int h = 0; h < NH; ++h) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
#pragma unroll
          for (int e = 0; e < NP; e += 2) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c_ok) v = __ldg(reinterpret_cast<const float4*>(a.wt + (size_t)t * a.C + h * a.Cout + c0 + 2 * e));
            w[h][t][e] = pk2(v.x, v.y);
            w[h][t][e + 1] = pk2(v.z, v.w);
          }
#pragma unroll
        for (int e = 0; e < NP; ++e)
          bv[h][e] = (a.bias && c_ok) ? pk2(a.bias[h * a.Cout + c0 + 2 * e], a.bias[h * a.Cout + c0 + 2 * e + 1])
                                      : pk2(0.f, 0.f);
      }
  }
  int it = 0;
  if (tid == 0 && grp < n_spatial) issue(grp, 0);
  for (int sp = grp; sp < n_spatial; sp += ngrp, ++it) {
    const int stage = it & 1;
    if (tid == 0 && sp + ngrp < n_spatial) issue(sp + ngrp, stage ^ 1);
    int b, y0, x0;
    dw_tile_coords(a, sp, b, y0, x0);
    mbar_wait(&full[stage], (it >> 1) & 1);
    const uint8_t* tile_s = dsm + stage * kDwStageBytes + xl * PIX_PITCH + cgi * (VEC * 2);
    f2 acc[3][NH][NP];                                     // a slot is (re)started by its first tap: fma(w00, v, bias)
    const int x = x0 + xl;
    bf16* outp = a.out + (((long long)b * a.H + y0) * a.W + x) * a.out_ld + c0;
    const bf16* dgp = GATE == 2 ? a.dg + (((long long)b * a.H + y0) * a.W + x) * a.dg_ld + c0 : nullptr;
    bf16* yp = GATE == 3 ? a.y_out + (((long long)b * a.H + y0) * a.W + x) * a.y_ld + c0 : nullptr;
    const bool st_ok = c_ok && x < a.W;
#pragma unroll
    for (int i = 0; i < kDwRows + 2; ++i) {                // staged row i = input row y0 - 1 + i
#pragma unroll
      for (int h = 0; h < NH; ++h)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const raw_t rv = *reinterpret_cast<const raw_t*>(tile_s + h * BOX_BYTES +
                                                           (i * (kDwCols + 2) + kx) * PIX_PITCH);
          const uint32_t* u = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
          for (int e = 0; e < NP; ++e) {
            const f2 v = x2_to_f2<HALF>(u[e]);
#pragma unroll
            for (int k = 0; k < 3; ++k)                    // staged row i feeds output rows i - 2 + k: skip the ones outside
              if (i + k >= 2 && i + k < kDwRows + 2)       // the tile (20 % of the halo tile's FMAs; i, k are constants)
                acc[(i + k) % 3][h][e] = fma2(w[h][(2 - k) * 3 + kx][e], v,
                                              (k == 2 && kx == 0) ? bv[h][e] : acc[(i + k) % 3][h][e]);
          }
        }
      if (i >= 2) {                                        // output row y0 + i - 2 (slot i % 3) is complete
        // The gate arithmetic runs unconditionally and only the stores are predicated: inside a divergent region the
        // GELU of row r could not be interleaved with the taps of row r + 1 (BSSY / BSYNC around every row).
        const bool row_ok = st_ok && y0 + i - 2 < a.H;
        {
          if constexpr (GATE == 2) {
            raw_t gr;
            {
              uint32_t* gz = reinterpret_cast<uint32_t*>(&gr);
#pragma unroll
              for (int e = 0; e < NP; ++e) gz[e] = 0u;
            }
            if (row_ok) gr = *reinterpret_cast<const raw_t*>(dgp);
            const uint32_t* gu = reinterpret_cast<const uint32_t*>(&gr);
            raw_t oa, ob;
            uint32_t* oau = reinterpret_cast<uint32_t*>(&oa);
            uint32_t* obu = reinterpret_cast<uint32_t*>(&ob);
#pragma unroll
            for (int e = 0; e < NP; ++e) {
              const f2 av = acc[i % 3][0][e], bvv = acc[i % 3][1][e];
              f2 gv = bf2_to_f2(gu[e]);
              if (a.dg_add && row_ok)
                gv = fma2(pk2(1.f, 1.f), gv, pk2(a.dg_add[(long long)b * a.Cout + c0 + 2 * e],
                                                  a.dg_add[(long long)b * a.Cout + c0 + 2 * e + 1]));
              f2 da, db;
              if (a.gate == 1) {
                f2 ga, dga;
                gelu2_grad(av, ga, dga);
                da = mul2(mul2(gv, bvv), dga);
                db = mul2(gv, ga);
              } else {
                da = mul2(gv, bvv);
                db = mul2(gv, av);
              }
              float t0, t1;
              upk2(da, t0, t1);
              oau[e] = pack2(t0, t1);
              upk2(db, t0, t1);
              obu[e] = pack2(t0, t1);
            }
            if (row_ok) {
              *reinterpret_cast<raw_t*>(outp) = oa;
              *reinterpret_cast<raw_t*>(outp + a.Cout) = ob;
            }
          } else {
            raw_t o;
            uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
            if (GATE == 3 && row_ok) {                      // training: keep the pre-gate tensor for the gate backward
              raw_t ya, yb;
              uint32_t* yau = reinterpret_cast<uint32_t*>(&ya);
              uint32_t* ybu = reinterpret_cast<uint32_t*>(&yb);
#pragma unroll
              for (int e = 0; e < NP; ++e) {
                float t0, t1;
                upk2(acc[i % 3][0][e], t0, t1);
                yau[e] = pack2(t0, t1);
                upk2(acc[i % 3][NH - 1][e], t0, t1);
                ybu[e] = pack2(t0, t1);
              }
              *reinterpret_cast<raw_t*>(yp) = ya;
              *reinterpret_cast<raw_t*>(yp + a.Cout) = yb;
            }
#pragma unroll
            for (int e = 0; e < NP; ++e) {
              f2 val = acc[i % 3][0][e];
              if (GATE) val = mul2(a.gate == 1 ? gelu2(val) : val, acc[i % 3][NH - 1][e]);
              float a0, a1;
              upk2(val, a0, a1);
              if (!GATE && a.relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); }
              ou[e] = pack2t<HALF>(a0, a1);
            }
            if (row_ok) *reinterpret_cast<raw_t*>(outp) = o;
          }
        }
        outp += (long long)a.W * a.out_ld;
        if (GATE == 2) dgp += (long long)a.W * a.dg_ld;
        if (GATE == 3) yp += (long long)a.W * a.y_ld;
      }
    }
    __syncthreads();                                       // everyone is done with this stage before it is refilled
  }
}


// ------------------------------------------------------------------------------------------------ weight packing
// nn.Conv2d weight fp32 [Co][Ci][KH][KW] -> bf16 GEMM operand [T][rows][ld] (K contiguous, zero padded), optionally
// through padded->logical channel maps (GDFN halves) and a per-output-channel scale (NAF gamma fold):
//   fwd  : out[t][co_p][ci_p]   = scale[co] * w[co][ci][t]
//   dgrad: out_t[t][ci_p][co_p] = scale[co] * w[co][ci][T-1-t]     (transposed, taps flipped)
// One launch per parameter replaces the chain of small torch ops that re-packed the weights after every optimizer step.
__global__ void __launch_bounds__(256) pack_conv_weight_kernel(const float* __restrict__ w, int Co, int Ci, int T,
                                                               const int* __restrict__ co_map, int Co_p,
                                                               const int* __restrict__ ci_map, int Ci_p,
                                                               const float* __restrict__ scale, bf16* __restrict__ out,
                                                               long long ld, bf16* __restrict__ out_t, long long ld_t,
                                                               int fwd_fp16) {
  const long long n_f = out ? (long long)T * Co_p * ld : 0;
  const long long n_t = out_t ? (long long)T * Ci_p * ld_t : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_f + n_t; i += (long long)gridDim.x * blockDim.x) {
    const bool tr = i >= n_f;
    long long r = tr ? i - n_f : i;
    const long long l = tr ? ld_t : ld;
    const int col = (int)(r % l); r /= l;
    const int rows = tr ? Ci_p : Co_p;
    const int row = (int)(r % rows);
    const int t = (int)(r / rows);
    const int co_p = tr ? col : row, ci_p = tr ? row : col;
    float v = 0.f;
    if (co_p < Co_p && ci_p < Ci_p) {
      const int co = co_map ? co_map[co_p] : co_p;
      const int ci = ci_map ? ci_map[ci_p] : ci_p;
      if (co >= 0 && co < Co && ci >= 0 && ci < Ci) {
        v = w[((long long)co * Ci + ci) * T + (tr ? T - 1 - t : t)];
        if (scale) v *= scale[co];
      }
    }
    if (tr) out_t[i - n_f] = __float2bfloat16(v);
    else reinterpret_cast<uint16_t*>(out)[i] = pack1r(v, fwd_fp16);
  }
}

// depthwise weight [C][1][3][3] (+ bias [C]) -> fp32 tap-major [9][C_p] (+ flipped copy for the data gradient, + padded bias)
__global__ void __launch_bounds__(256) pack_dw_weight_kernel(const float* __restrict__ w, const float* __restrict__ bias,
                                                             int C, const int* __restrict__ c_map, int C_p,
                                                             float* __restrict__ out, float* __restrict__ out_flip,
                                                             float* __restrict__ out_bias) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 10 * C_p) return;
  const int t = i / C_p, cp = i % C_p;
  const int c = c_map ? c_map[cp] : cp;
  const bool ok = c >= 0 && c < C;
  if (t == 9) {
    if (out_bias) out_bias[cp] = (ok && bias) ? bias[c] : 0.f;
    return;
  }
  const float v = ok ? w[(long long)c * 9 + t] : 0.f;
  if (out) out[(long long)t * C_p + cp] = v;
  if (out_flip) out_flip[(long long)(8 - t) * C_p + cp] = v;
}

// out[i] = map[i] >= 0 ? v[map[i]] : 0   (padded bias vectors)
__global__ void gather_vec_kernel(const float* __restrict__ v, const int* __restrict__ map, int n, int n_src,
                                  float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = map ? map[i] : i;
  out[i] = (j >= 0 && j < n_src) ? v[j] : 0.f;
}

// ------------------------------------------------------------------------------------------------ layout
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, int B, int C, int H, int W, int PH, int PW,
                                    float* __restrict__ d32, long long ld32, bf16* __restrict__ d16, long long ld16) {
  const long long total = (long long)B * PH * PW * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int x = (int)(p % PW);
    const int y = (int)((p / PW) % PH);
    const int b = (int)(p / ((long long)PW * PH));
    const float v = (y < H && x < W) ? src[(((long long)b * C + c) * H + y) * W + x] : 0.f;
    if (d32) d32[p * ld32 + c] = v;
    if (d16) d16[p * ld16 + c] = __float2bfloat16(v);
  }
}

// NCHW fp32 image -> zero-padded 16-bit NHWC rows with C16 >= C channels (all C16 are written): the tensor-core operand of
// the image-boundary 3x3 convs (patch_embed, MASA conv_L1), whose 3 input channels are padded to one 16-byte TMA row
__global__ void image_to_rows16_kernel(const float* __restrict__ src, int B, int C, int H, int W, int PH, int PW, int C16,
                                       uint16_t* __restrict__ dst, int fp16) {
  const long long total = (long long)B * PH * PW * C16;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C16);
    long long p = i / C16;
    const int x = (int)(p % PW);
    const int y = (int)((p / PW) % PH);
    const int b = (int)(p / ((long long)PW * PH));
    const float v = (c < C && y < H && x < W) ? src[(((long long)b * C + c) * H + y) * W + x] : 0.f;
    dst[i] = pack1r(v, fp16);
  }
}

__global__ void nhwc_to_nchw_kernel(const float* __restrict__ src, long long ld, int B, int C, int H, int W, int OH,
                                    int OW, const float* __restrict__ res, long long res_ld,
                                    float* __restrict__ dst) {
  const long long total = (long long)B * C * OH * OW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % OW);
    const int y = (int)((i / OW) % OH);
    const int c = (int)((i / ((long long)OW * OH)) % C);
    const int b = (int)(i / ((long long)OW * OH * C));
    const long long p = ((long long)b * H + y) * W + x;
    dst[i] = src[p * ld + c] + (res ? res[p * res_ld + c] : 0.f);
  }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, long long src_ld, long long rows, int C,
                                 float* __restrict__ dst, long long dst_ld, bf16* __restrict__ d16, long long ld16,
                                 int fp16) {
  const int nvec = C >> 2;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i % nvec) * 4;
    const float4 v = *reinterpret_cast<const float4*>(src + r * src_ld + c);
    if (dst) *reinterpret_cast<float4*>(dst + r * dst_ld + c) = v;
    if (d16) {
      uint2 pk;
      pk.x = pack2r(v.x, v.y, fp16);
      pk.y = pack2r(v.z, v.w, fp16);
      *reinterpret_cast<uint2*>(d16 + r * ld16 + c) = pk;
    }
  }
}

// ------------------------------------------------------------------------------------------------ small convs
// Ci <= 8 input channels (fp32 NHWC image), Co % 8 == 0 outputs.  A thread owns TWO horizontally adjacent pixels: the
// 3x4 input window is loaded once into registers and the thread walks the output channels in groups of 16, so every
// weight read from shared memory ([ci*9+tap][Co]) is one warp-wide broadcast (all lanes are on the same channel group:
// lanes on different groups made every LDS.128 a two-way bank conflict, 128 B apart) that feeds two pixels, and each
// pixel's 16 outputs are stored as 64 B runs.
template <int CI, int CPT>
__global__ void __launch_bounds__(256) conv3x3_small_ci_kernel(const float* __restrict__ in, int B, int H, int W,
                                                               const float* __restrict__ weight,
                                                               const float* __restrict__ bias, int Co, int relu,
                                                               float* __restrict__ o32, long long ld32,
                                                               bf16* __restrict__ o16, long long ld16) {
  extern __shared__ float ws[];
  for (int i = threadIdx.x; i < Co * CI * 9; i += blockDim.x) {
    const int co = i / (CI * 9), r = i % (CI * 9);       // weight[co][ci][ky][kx] -> ws[(ci*9+tap)*Co + co]
    ws[r * Co + co] = weight[i];
  }
  __syncthreads();
  const int ncg = Co / CPT;                               // CPT-channel groups (16, or 8 when Co % 16 != 0)
  const int wp = (W + 1) >> 1;                            // pixel pairs per row
  const long long total = (long long)B * H * wp;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < total;
       p += (long long)gridDim.x * blockDim.x) {
    const int xp = (int)(p % wp);
    const int y = (int)((p / wp) % H);
    const int b = (int)(p / ((long long)wp * H));
    const int x0 = xp * 2;
    float win[3][4][CI];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int iy = y + dy - 1, ix = x0 + dx - 1;
        const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
        const float* ip = in + (((long long)b * H + (ok ? iy : 0)) * W + (ok ? ix : 0)) * CI;
#pragma unroll
        for (int ci = 0; ci < CI; ++ci) win[dy][dx][ci] = ok ? __ldg(ip + ci) : 0.f;
      }
    for (int cg = 0; cg < ncg; ++cg) {
    float acc[2][CPT];
#pragma unroll
    for (int e = 0; e < CPT; ++e) acc[0][e] = acc[1][e] = bias ? bias[cg * CPT + e] : 0.f;
#pragma unroll
    for (int ci = 0; ci < CI; ++ci)
#pragma unroll
      for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4* wp4 = reinterpret_cast<const float4*>(ws + (ci * 9 + dy * 3 + dx) * Co + cg * CPT);
          const float v0 = win[dy][dx][ci], v1 = win[dy][dx + 1][ci];
#pragma unroll
          for (int q4 = 0; q4 < CPT / 4; ++q4) {
            const float4 w4 = wp4[q4];
            acc[0][q4 * 4 + 0] = fmaf(v0, w4.x, acc[0][q4 * 4 + 0]); acc[1][q4 * 4 + 0] = fmaf(v1, w4.x, acc[1][q4 * 4 + 0]);
            acc[0][q4 * 4 + 1] = fmaf(v0, w4.y, acc[0][q4 * 4 + 1]); acc[1][q4 * 4 + 1] = fmaf(v1, w4.y, acc[1][q4 * 4 + 1]);
            acc[0][q4 * 4 + 2] = fmaf(v0, w4.z, acc[0][q4 * 4 + 2]); acc[1][q4 * 4 + 2] = fmaf(v1, w4.z, acc[1][q4 * 4 + 2]);
            acc[0][q4 * 4 + 3] = fmaf(v0, w4.w, acc[0][q4 * 4 + 3]); acc[1][q4 * 4 + 3] = fmaf(v1, w4.w, acc[1][q4 * 4 + 3]);
          }
        }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
      if (x0 + px >= W) continue;
      const long long pix = ((long long)b * H + y) * W + x0 + px;
      if (relu & 1) {
#pragma unroll
        for (int e = 0; e < CPT; ++e) acc[px][e] = fmaxf(acc[px][e], 0.f);
      }
      if (o32) {
        float4* q = reinterpret_cast<float4*>(o32 + pix * ld32 + cg * CPT);
#pragma unroll
        for (int q4 = 0; q4 < CPT / 4; ++q4)
          q[q4] = make_float4(acc[px][q4 * 4], acc[px][q4 * 4 + 1], acc[px][q4 * 4 + 2], acc[px][q4 * 4 + 3]);
      }
      if (o16) {
#pragma unroll
        for (int q8 = 0; q8 < CPT / 8; ++q8) {
          const float* v = acc[px] + q8 * 8;
          if (relu & 2)                               // IEEE fp16 output (MASA feature encoder operands)
            *reinterpret_cast<uint4*>(o16 + pix * ld16 + cg * CPT + q8 * 8) =
                make_uint4(pack2h(v[0], v[1]), pack2h(v[2], v[3]), pack2h(v[4], v[5]), pack2h(v[6], v[7]));
          else
            *reinterpret_cast<bf16x8*>(o16 + pix * ld16 + cg * CPT + q8 * 8) = pack8(v);
        }
      }
    }
    }   // cg
  }
}

// fp32 rows -> bf16 and / or fp16 copies (8 channels per thread), optional device-side power-of-two scales
__global__ void __launch_bounds__(256) cast_rows_kernel(float* __restrict__ src, long long src_ld, long long rows,
                                                        int C, bf16* __restrict__ d16, long long ld16,
                                                        __half* __restrict__ dh, long long ldh,
                                                        const float* __restrict__ scale16,
                                                        const float* __restrict__ scale_bf16, int rescale_in) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  pdl_wait();
  pdl_launch();
  const float sh = scale16 ? *scale16 : 1.f, sb = scale_bf16 ? *scale_bf16 : 1.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i % nvec) * 8;
    float4 a = *reinterpret_cast<const float4*>(src + r * src_ld + c);
    float4 b = *reinterpret_cast<const float4*>(src + r * src_ld + c + 4);
    if (d16)
      *reinterpret_cast<uint4*>(d16 + r * ld16 + c) = make_uint4(pack2(a.x * sb, a.y * sb), pack2(a.z * sb, a.w * sb),
                                                                 pack2(b.x * sb, b.y * sb), pack2(b.z * sb, b.w * sb));
    a.x *= sh; a.y *= sh; a.z *= sh; a.w *= sh; b.x *= sh; b.y *= sh; b.z *= sh; b.w *= sh;
    if (dh)
      *reinterpret_cast<uint4*>(dh + r * ldh + c) = make_uint4(pack2h(a.x, a.y), pack2h(a.z, a.w), pack2h(b.x, b.y), pack2h(b.z, b.w));
    if (rescale_in) {
      *reinterpret_cast<float4*>(src + r * src_ld + c) = a;
      *reinterpret_cast<float4*>(src + r * src_ld + c + 4) = b;
    }
  }
}

// max |x| over fp32 rows (x >= 0 bit patterns order like unsigned integers) -> state[3]
__global__ void __launch_bounds__(256) absmax_rows_kernel(const float* __restrict__ x, long long ld, long long rows, int C,
                                                          float* __restrict__ state) {
  const int nvec = C >> 2;
  const long long total = rows * nvec;
  float m = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float4 v = *reinterpret_cast<const float4*>(x + (i / nvec) * ld + (i % nvec) * 4);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && m == m) atomicMax(reinterpret_cast<unsigned int*>(state + 3), __float_as_uint(m));
}

__global__ void level_scale_finalize_kernel(const float* __restrict__ prev, float* __restrict__ cur) {
  const float m = cur[3];
  float r = 1.f;
  if (m > 64.f && m < 3.0e38f) r = exp2f(ceilf(log2f(m)) - 6.f);          // max |x / r| in (32, 64]
  const float s = (prev ? prev[0] : 1.f) * r;
  cur[0] = s;
  cur[1] = 1.f / s;
  cur[2] = 1.f / r;
}

__global__ void __launch_bounds__(256) scale_vec_kernel(const float* __restrict__ v, long long n,
                                                        const float* __restrict__ scale, float* __restrict__ out) {
  const float sc = *scale;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = v[i] * sc;
}

__global__ void __launch_bounds__(256) cvt_f16_bf16_kernel(const __half* __restrict__ src, long long src_ld, long long rows,
                                                           int C, bf16* __restrict__ dst, long long dst_ld) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i % nvec) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(src + r * src_ld + c);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[k]));
      o[k] = pack2(f.x, f.y);
    }
    *reinterpret_cast<uint4*>(dst + r * dst_ld + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Co <= 4 outputs from bf16 NHWC input; 4 lanes share one pixel (channel vectors strided by 4), shuffle-reduced.
__global__ void __launch_bounds__(256) conv3x3_small_co_kernel(const bf16* __restrict__ in, long long in_ld, int B,
                                                               int H, int W, int Ci, const float* __restrict__ weight,
                                                               const float* __restrict__ bias, int Co,
                                                               const float* __restrict__ res, float* __restrict__ out) {
  extern __shared__ float ws[];       // ws[(tap*Ci + ci)*4 + co]
  for (int i = threadIdx.x; i < 9 * Ci * 4; i += blockDim.x) {
    const int co = i & 3, r = i >> 2;
    const int tap = r / Ci, ci = r % Ci;
    ws[i] = co < Co ? weight[((long long)co * Ci + ci) * 9 + tap] : 0.f;
  }
  __syncthreads();
  const int nvec = Ci >> 3;
  const long long total = (long long)B * H * W;
  const int l4 = threadIdx.x & 3;
  for (long long p0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 2; p0 < ((total + 7) / 8) * 8;
       p0 += ((long long)gridDim.x * blockDim.x) >> 2) {
    const bool ok = p0 < total;
    const long long p = ok ? p0 : total - 1;
    const int x = (int)(p % W);
    const int y = (int)((p / W) % H);
    const int b = (int)(p / ((long long)W * H));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int tap = 0; tap < 9; ++tap) {
      const int iy = y + tap / 3 - 1, ix = x + tap % 3 - 1;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      const bf16* ip = in + (((long long)b * H + iy) * W + ix) * in_ld;
      for (int v = l4; v < nvec; v += 4) {
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(ip + v * 8), f);
        const float4* wp = reinterpret_cast<const float4*>(ws + (tap * Ci + v * 8) * 4);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float4 wv = wp[e];
          acc[0] = fmaf(f[e], wv.x, acc[0]);
          acc[1] = fmaf(f[e], wv.y, acc[1]);
          acc[2] = fmaf(f[e], wv.z, acc[2]);
          acc[3] = fmaf(f[e], wv.w, acc[3]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 1);
      acc[c] += __shfl_xor_sync(0xffffffffu, acc[c], 2);
    }
    if (ok && l4 < Co) {
      float v = acc[l4] + (bias ? bias[l4] : 0.f);
      if (res) v += res[p * Co + l4];
      out[p * Co + l4] = v;
    }
  }
}

inline int grid_for(long long items, int per_block, int max_waves = 8) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = (long long)tdr_num_sms() * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

extern "C" int tdr_rownorm(const float* in, long long in_ld, long long rows, int C, int mode, const float* weight,
                           const float* bias, float eps, int act, void* out_bf16, long long out_ld, float* out_f32,
                           long long out_f32_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(in && (out_bf16 || out_f32) && rows >= 0 && C > 0, "tdr_rownorm: bad arguments");
  TDR_CHECK_ARG(C % 4 == 0 && in_ld % 4 == 0 && out_ld % 4 == 0 && out_f32_ld % 4 == 0,
                "tdr_rownorm: C and strides must be multiples of 4");
  TDR_CHECK_ARG((act & ~16) == 0 || (act & ~16) == 3, "tdr_rownorm: act must be 0 or 3 (LeakyReLU), optionally | 16 (fp16 output)");
  TDR_CHECK_ARG(mode >= 0 && mode <= 2, "tdr_rownorm: bad mode");
  TDR_CHECK_ARG(mode == 0 || weight, "tdr_rownorm: weight required");
  TDR_CHECK_ARG(C <= 4096, "tdr_rownorm: C too large (%d)", C);
  if (rows == 0) return TDR_OK;
  const int nvec = C / 4;
  // G lanes per row: the smallest power of two that keeps <= 4 float4 per lane (all lanes busy, short shuffle trees,
  // 32/G rows per warp); very wide rows fall back to a full warp per row.
  int G = 1;
  while (G < 32 && (nvec + G - 1) / G > 4) G <<= 1;
  const int nv = (nvec + G - 1) / G;
  const long long warps = (rows + (32 / G) - 1) / (32 / G);
  const int blocks = (int)((warps + 7) / 8);
  bf16* o = reinterpret_cast<bf16*>(out_bf16);
#define TDR_RN(NV) \
  TDR_CHECK_CUDA(tdr_launch_pdl(rownorm_kernel<NV>, dim3(blocks), dim3(256), 0, stream, in, in_ld, rows, C, mode, weight, bias, eps, \
                                act, o, out_ld, out_f32, out_f32_ld, G))
  if (nv <= 1) TDR_RN(1);
  else if (nv <= 2) TDR_RN(2);
  else if (nv <= 3) TDR_RN(3);
  else if (nv <= 4) TDR_RN(4);
  else if (nv <= 8) TDR_RN(8);
  else if (nv <= 16) TDR_RN(16);
  else TDR_RN(32);
#undef TDR_RN
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

static int dwconv_launch(const void* in_bf16, long long in_ld, int B, int H, int W, int C, const float* weight,
                         const float* bias, int gate, void* out_bf16, long long out_ld, const void* dg_bf16,
                         long long dg_ld, const float* dg_add, cudaStream_t stream, void* y_out = nullptr,
                         long long y_ld = 0) {
  TDR_CHECK_ARG(!y_out || (gate >= 1 && !dg_bf16 && y_ld >= C && y_ld % 4 == 0 && ((uintptr_t)y_out & 7) == 0),
                "tdr_dwconv3x3_gated_train: y_out needs a gated forward launch and room for C channels");
  const bool bwd = dg_bf16 != nullptr;
  TDR_CHECK_ARG(in_bf16 && out_bf16 && weight, "tdr_dwconv3x3: null pointer");
  TDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0, "tdr_dwconv3x3: bad dims");
  const bool half = (gate & 16) != 0;               // bit 4: activations are IEEE fp16 (inference forward), else bf16
  const bool relu = (gate & 32) != 0;               // bit 5: ReLU on the (ungated) output -- DRSformer MSFN dwconv3x3 :242
  gate &= 15;
  TDR_CHECK_ARG(!relu || (gate == 0 && !dg_bf16), "tdr_dwconv3x3: ReLU goes with the plain (ungated) forward only");
  TDR_CHECK_ARG(gate >= 0 && gate <= 2, "tdr_dwconv3x3: bad gate");
  TDR_CHECK_ARG(!half || (!dg_bf16 && !y_out), "tdr_dwconv3x3: the training / backward variants take bf16 activations");
  TDR_CHECK_ARG(C % 8 == 0, "tdr_dwconv3x3: C must be a multiple of 8");
  TDR_CHECK_ARG(in_ld % 8 == 0 && out_ld % 4 == 0, "tdr_dwconv3x3: bad strides");
  TDR_CHECK_ARG(((uintptr_t)in_bf16 & 15) == 0 && ((uintptr_t)out_bf16 & 7) == 0, "tdr_dwconv3x3: alignment");
  if (!gate) TDR_CHECK_ARG(out_ld % 8 == 0 && ((uintptr_t)out_bf16 & 15) == 0, "tdr_dwconv3x3: output alignment");
  if (bwd) TDR_CHECK_ARG(gate >= 1 && (C / 2) % 4 == 0 && dg_ld % 4 == 0 && ((uintptr_t)dg_bf16 & 7) == 0,
                         "tdr_dwconv3x3_gate_bwd: bad gate / alignment");
  const bf16* in = reinterpret_cast<const bf16*>(in_bf16);
  bf16* out = reinterpret_cast<bf16*>(out_bf16);
  static const bool use_strip = getenv("TDR_DWCONV_STRIP") != nullptr;     // previous (non-TMA) kernel, experiments only
  if (use_strip && !bwd && !half) {
    constexpr int R = 16;
    if (gate) {
      const int items = W * (C / 2 / 4);
      dim3 grid((items + 255) / 256, (H + R - 1) / R, B);
      dwconv3x3_strip_kernel<2, 4, R><<<grid, 256, 0, stream>>>(in, in_ld, H, W, C, weight, bias, gate, out, out_ld);
    } else {
      const int items = W * (C / 8);
      dim3 grid((items + 255) / 256, (H + R - 1) / R, B);
      dwconv3x3_strip_kernel<1, 8, R><<<grid, 256, 0, stream>>>(in, in_ld, H, W, C, weight, bias, gate, out, out_ld);
    }
    TDR_CHECK_LAUNCH();
    return TDR_OK;
  }
  DwArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C; a.Cout = gate ? C / 2 : C;
  a.tiles_x = tdr_cdiv(W, kDwCols); a.tiles_y = tdr_cdiv(H, kDwRows);
  tdr_fast_div_setup(a.tiles_y * a.tiles_x, &a.fd_img_m, &a.fd_img_s);
  tdr_fast_div_setup(a.tiles_x, &a.fd_tx_m, &a.fd_tx_s);
  const int cb = gate ? 32 : 64;
  a.chunks = tdr_cdiv(a.Cout, cb);
  a.total_tiles = B * a.tiles_x * a.tiles_y * a.chunks;
  a.wt = weight; a.bias = bias; a.gate = gate; a.out = out; a.out_ld = out_ld;
  a.dg = reinterpret_cast<const bf16*>(dg_bf16); a.dg_ld = dg_ld; a.dg_add = dg_add;
  a.y_out = reinterpret_cast<bf16*>(y_out); a.y_ld = y_ld;
  a.relu = relu ? 1 : 0;
  TdrTensorMap map;
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
  const uint64_t strides[3] = {(uint64_t)in_ld * 2, (uint64_t)in_ld * 2 * W, (uint64_t)in_ld * 2 * W * H};
  const uint32_t box[4] = {(uint32_t)cb, (uint32_t)(kDwCols + 2), (uint32_t)(kDwRows + 2), 1};
  const uint32_t es[4] = {1, 1, 1, 1};
  int rc = tdr_make_tensor_map_bf16_noswizzle(&map, in, 4, dims, strides, box, es);   // L2 promotion 128 B
  if (rc) return rc;
  const int n_spatial = B * a.tiles_x * a.tiles_y;
  int groups = (2 * tdr_num_sms()) / a.chunks;
  if (groups < 1) groups = 1;
  if (groups > n_spatial) groups = n_spatial;
  const int grid = groups * a.chunks;
  const size_t smem = 2 * kDwStageBytes;
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_tma_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_tma_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_tma_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_tma_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_tma_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dwconv3x3_tma_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (bwd) TDR_CHECK_CUDA(tdr_launch_pdl(dwconv3x3_tma_kernel<2, false>, dim3(grid), dim3(256), smem, stream, map, a));
  else if (gate && y_out) TDR_CHECK_CUDA(tdr_launch_pdl(dwconv3x3_tma_kernel<3, false>, dim3(grid), dim3(256), smem, stream, map, a));
  else if (gate && half) TDR_CHECK_CUDA(tdr_launch_pdl(dwconv3x3_tma_kernel<1, true>, dim3(grid), dim3(256), smem, stream, map, a));
  else if (gate) TDR_CHECK_CUDA(tdr_launch_pdl(dwconv3x3_tma_kernel<1, false>, dim3(grid), dim3(256), smem, stream, map, a));
  else if (half) TDR_CHECK_CUDA(tdr_launch_pdl(dwconv3x3_tma_kernel<0, true>, dim3(grid), dim3(256), smem, stream, map, a));
  else TDR_CHECK_CUDA(tdr_launch_pdl(dwconv3x3_tma_kernel<0, false>, dim3(grid), dim3(256), smem, stream, map, a));
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_dwconv3x3(const void* in_bf16, long long in_ld, int B, int H, int W, int C, const float* weight,
                             const float* bias, int gate, void* out_bf16, long long out_ld, cudaStream_t stream) {
  return dwconv_launch(in_bf16, in_ld, B, H, W, C, weight, bias, gate, out_bf16, out_ld, nullptr, 0, nullptr, stream);
}

extern "C" int tdr_dwconv3x3_gated_train(const void* in_bf16, long long in_ld, int B, int H, int W, int C,
                                         const float* weight, const float* bias, int gate, void* out_bf16,
                                         long long out_ld, void* y_bf16, long long y_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(gate == 1 || gate == 2, "tdr_dwconv3x3_gated_train: gate must be 1 or 2");
  return dwconv_launch(in_bf16, in_ld, B, H, W, C, weight, bias, gate, out_bf16, out_ld, nullptr, 0, nullptr, stream,
                       y_bf16, y_ld);
}

extern "C" int tdr_dwconv3x3_gate_bwd(const void* in_bf16, long long in_ld, int B, int H, int W, int C,
                                      const float* weight, const float* bias, int gate, const void* dg_bf16,
                                      long long dg_ld, const float* dg_add, void* dy_bf16, long long dy_ld,
                                      cudaStream_t stream) {
  TDR_CHECK_ARG(dg_bf16 != nullptr && (gate == 1 || gate == 2), "tdr_dwconv3x3_gate_bwd: dg and gate 1|2 required");
  TDR_CHECK_ARG(dy_ld >= C && dy_ld % 4 == 0, "tdr_dwconv3x3_gate_bwd: dy must hold 2 * (C/2) channels");
  return dwconv_launch(in_bf16, in_ld, B, H, W, C, weight, bias, gate, dy_bf16, dy_ld, dg_bf16, dg_ld, dg_add, stream);
}

extern "C" int tdr_nchw_to_nhwc(const float* src, int B, int C, int H, int W, int pad_h, int pad_w, float* dst_f32,
                                long long dst_f32_ld, void* dst_bf16, long long dst_bf16_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(src && (dst_f32 || dst_bf16), "tdr_nchw_to_nhwc: null pointer");
  TDR_CHECK_ARG(pad_h >= H && pad_w >= W && B > 0 && C > 0, "tdr_nchw_to_nhwc: bad dims");
  const long long total = (long long)B * pad_h * pad_w * C;
  nchw_to_nhwc_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(src, B, C, H, W, pad_h, pad_w, dst_f32, dst_f32_ld,
                                                                    reinterpret_cast<bf16*>(dst_bf16), dst_bf16_ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_image_to_rows16(const float* src, int B, int C, int H, int W, int pad_h, int pad_w, int C16, void* dst16,
                                   int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(src && dst16 && pad_h >= H && pad_w >= W && B > 0 && C > 0 && C16 >= C && C16 % 8 == 0,
                "tdr_image_to_rows16: bad arguments (C16 must be a multiple of 8, >= C)");
  const long long total = (long long)B * pad_h * pad_w * C16;
  image_to_rows16_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(src, B, C, H, W, pad_h, pad_w, C16,
                                                                       reinterpret_cast<uint16_t*>(dst16), fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_nhwc_to_nchw(const float* src, long long src_ld, int B, int C, int H, int W, int out_h, int out_w,
                                const float* res, long long res_ld, float* dst, cudaStream_t stream) {
  TDR_CHECK_ARG(src && dst && out_h <= H && out_w <= W && out_h > 0 && out_w > 0, "tdr_nhwc_to_nchw: bad arguments");
  const long long total = (long long)B * C * out_h * out_w;
  nhwc_to_nchw_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(src, src_ld, B, C, H, W, out_h, out_w, res, res_ld,
                                                                    dst);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_copy_rows_f32(const float* src, long long src_ld, long long rows, int C, float* dst,
                                 long long dst_ld, void* dst_bf16, long long dst_bf16_ld, int dst16_fp16,
                                 cudaStream_t stream) {
  TDR_CHECK_ARG(src && (dst || dst_bf16), "tdr_copy_rows_f32: null pointer");
  TDR_CHECK_ARG(C % 4 == 0 && src_ld % 4 == 0 && dst_ld % 4 == 0 && dst_bf16_ld % 4 == 0,
                "tdr_copy_rows_f32: C / strides must be multiples of 4");
  if (rows == 0) return TDR_OK;
  copy_rows_kernel<<<grid_for(rows * (C / 4), 256, 16), 256, 0, stream>>>(src, src_ld, rows, C, dst, dst_ld,
                                                                         reinterpret_cast<bf16*>(dst_bf16), dst_bf16_ld,
                                                                         dst16_fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_cast_rows(float* in, long long in_ld, long long rows, int C, void* out_bf16, long long out_bf16_ld,
                             void* out_fp16, long long out_fp16_ld, const float* scale16, const float* scale_bf16,
                             int rescale_in, cudaStream_t stream) {
  TDR_CHECK_ARG(in && (out_bf16 || out_fp16), "tdr_cast_rows: null pointer");
  TDR_CHECK_ARG(C % 8 == 0 && in_ld % 4 == 0 && out_bf16_ld % 8 == 0 && out_fp16_ld % 8 == 0 && ((uintptr_t)in & 15) == 0 &&
                ((uintptr_t)out_bf16 & 15) == 0 && ((uintptr_t)out_fp16 & 15) == 0, "tdr_cast_rows: C %% 8, 16 B-aligned rows");
  if (rows == 0) return TDR_OK;
  TDR_CHECK_CUDA(tdr_launch_pdl(cast_rows_kernel, dim3(grid_for(rows * (C / 8), 256, 16)), dim3(256), 0, stream, in, in_ld, rows,
                                C, reinterpret_cast<bf16*>(out_bf16), out_bf16_ld, reinterpret_cast<__half*>(out_fp16),
                                out_fp16_ld, scale16, scale_bf16, rescale_in));
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_masa_level_scale(const float* x, long long ld, long long rows, int C, const float* state_prev,
                                    float* state_cur, cudaStream_t stream) {
  TDR_CHECK_ARG(x && state_cur && rows > 0 && C % 4 == 0 && ld % 4 == 0 && ((uintptr_t)x & 15) == 0,
                "tdr_masa_level_scale: bad arguments");
  absmax_rows_kernel<<<grid_for(rows * (C / 4), 256, 8), 256, 0, stream>>>(x, ld, rows, C, state_cur);
  TDR_CHECK_LAUNCH();
  level_scale_finalize_kernel<<<1, 1, 0, stream>>>(state_prev, state_cur);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_scale_vec(const float* v, long long n, const float* scale, float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(v && scale && out && n > 0, "tdr_scale_vec: bad arguments");
  scale_vec_kernel<<<grid_for(n, 256, 4), 256, 0, stream>>>(v, n, scale, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_cvt_f16_bf16(const void* in_fp16, long long in_ld, long long rows, int C, void* out_bf16, long long out_ld,
                                cudaStream_t stream) {
  TDR_CHECK_ARG(in_fp16 && out_bf16 && C % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && ((uintptr_t)in_fp16 & 15) == 0 &&
                ((uintptr_t)out_bf16 & 15) == 0, "tdr_cvt_f16_bf16: C %% 8, 16 B-aligned rows");
  if (rows == 0) return TDR_OK;
  cvt_f16_bf16_kernel<<<grid_for(rows * (C / 8), 256, 16), 256, 0, stream>>>(reinterpret_cast<const __half*>(in_fp16), in_ld,
                                                                            rows, C, reinterpret_cast<bf16*>(out_bf16), out_ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_conv3x3_small_ci(const float* in, int B, int H, int W, int Ci, const float* weight,
                                    const float* bias, int Co, int relu, float* out_f32, long long out_f32_ld,
                                    void* out_bf16, long long out_bf16_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(in && weight && (out_f32 || out_bf16), "tdr_conv3x3_small_ci: null pointer");
  TDR_CHECK_ARG(Ci >= 1 && Ci <= 8 && Co % 8 == 0 && Co * Ci * 9 * 4 <= 48 * 1024,
                "tdr_conv3x3_small_ci: need 1 <= Ci <= 8 and Co %% 8 == 0 (got Ci=%d Co=%d)", Ci, Co);
  TDR_CHECK_ARG(out_f32_ld % 4 == 0 && out_bf16_ld % 8 == 0, "tdr_conv3x3_small_ci: bad strides");
  const int cpt = Co % 16 == 0 ? 16 : 8;
  const long long items = (long long)B * H * ((W + 1) / 2);
  const size_t smem = (size_t)Co * Ci * 9 * sizeof(float);
  bf16* o16 = reinterpret_cast<bf16*>(out_bf16);
#define TDR_SCI(N)                                                                                                   \
  case N:                                                                                                            \
    if (cpt == 16)                                                                                                   \
      conv3x3_small_ci_kernel<N, 16><<<grid_for(items, 256, 8), 256, smem, stream>>>(                                \
          in, B, H, W, weight, bias, Co, relu, out_f32, out_f32_ld, o16, out_bf16_ld);                               \
    else                                                                                                             \
      conv3x3_small_ci_kernel<N, 8><<<grid_for(items, 256, 8), 256, smem, stream>>>(                                 \
          in, B, H, W, weight, bias, Co, relu, out_f32, out_f32_ld, o16, out_bf16_ld);                               \
    break
  switch (Ci) {
    TDR_SCI(1); TDR_SCI(2); TDR_SCI(3); TDR_SCI(4); TDR_SCI(5); TDR_SCI(6); TDR_SCI(7); TDR_SCI(8);
  }
#undef TDR_SCI
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_conv3x3_small_co(const void* in_bf16, long long in_ld, int B, int H, int W, int Ci,
                                    const float* weight, const float* bias, int Co, const float* res, float* out,
                                    cudaStream_t stream) {
  TDR_CHECK_ARG(in_bf16 && weight && out, "tdr_conv3x3_small_co: null pointer");
  TDR_CHECK_ARG(Co >= 1 && Co <= 4 && Ci % 8 == 0 && in_ld % 8 == 0 && 9 * Ci * 16 <= 48 * 1024,
                "tdr_conv3x3_small_co: bad channels");
  const long long items = (long long)B * H * W * 4;
  conv3x3_small_co_kernel<<<grid_for(items, 256, 8), 256, 9 * Ci * 4 * sizeof(float), stream>>>(
      reinterpret_cast<const bf16*>(in_bf16), in_ld, B, H, W, Ci, weight, bias, Co, res, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

// ------------------------------------------------------------------------------------------------ NAFNet helpers
// SimpleGate (network_nafnet_guided_arch.py:170-175) on bf16 rows: out[r, c] = x[r, c] * x[r, C + c]
namespace {
__global__ void __launch_bounds__(256) gate_mul_kernel(const bf16* __restrict__ x, long long ld, long long rows, int C,
                                                       bf16* __restrict__ out, long long out_ld, int fp16) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i % nvec) * 8;
    float a[8], b[8];
    unpack8r(x + r * ld + c, a, fp16);
    unpack8r(x + r * ld + C + c, b, fp16);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] *= b[e];
    pack8r(out + r * out_ld + c, a, fp16);
  }
}

// Global average pool, stage 1: grid (chunks, B); partial[b][chunk][c] = sum over the chunk's pixels of x[b, p, c]
__global__ void __launch_bounds__(256) pool_partial_kernel(const bf16* __restrict__ x, long long ld, long long P, int C,
                                                           int chunks, float* __restrict__ partial, int fp16) {
  const int nvec = C >> 3;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const long long per = (P + chunks - 1) / chunks;
  const long long p0 = chunk * per, p1 = p0 + per < P ? p0 + per : P;
  // thread -> (channel vector, pixel lane): all threads sharing a vector are reduced through shared memory
  const int lanes = blockDim.x / nvec > 0 ? blockDim.x / nvec : 1;
  __shared__ float red[256 * 8];
  for (int v0 = 0; v0 < nvec; v0 += blockDim.x) {          // nvec > 256 only for C > 2048
    const int v = v0 + (int)(threadIdx.x % (nvec < (int)blockDim.x ? nvec : (int)blockDim.x));
    const int pl = threadIdx.x / (nvec < (int)blockDim.x ? nvec : (int)blockDim.x);
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (v < nvec && pl < lanes) {
      for (long long p = p0 + pl; p < p1; p += lanes) {
        float f[8];
        unpack8r(x + ((long long)b * P + p) * ld + v * 8, f, fp16);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += f[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x * 8 + e] = acc[e];
    __syncthreads();
    const int nv_here = nvec - v0 < (int)blockDim.x ? nvec - v0 : (int)blockDim.x;
    if ((int)threadIdx.x < nv_here) {
      float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      for (int l = 0; l < lanes && l * nv_here + (int)threadIdx.x < (int)blockDim.x; ++l)
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += red[(l * nv_here + threadIdx.x) * 8 + e];
      float* dst = partial + ((size_t)b * chunks + chunk) * C + (v0 + threadIdx.x) * 8;
#pragma unroll
      for (int e = 0; e < 8; ++e) dst[e] = s[e];
    }
    __syncthreads();
  }
}

// Stage 2 (N:192-196): mean[b][c] = sum_chunks partial / P, then s[b][c] = W_sca[c,:] . mean[b] + b_sca[c].
// One warp per (b, c): the W_sca row is read coalesced, the mean vector comes from global (L1/L2 resident).
__global__ void __launch_bounds__(256) sca_mean_kernel(const float* __restrict__ partial, int chunks, long long P, int C,
                                                       int B, float* __restrict__ mean) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  float t = 0.f;
  for (int ch = 0; ch < chunks; ++ch) t += partial[((size_t)b * chunks + ch) * C + c];
  mean[idx] = t / (float)P;
}

__global__ void __launch_bounds__(256) sca_vec_kernel(const float* __restrict__ mean, const float* __restrict__ w_sca,
                                                      const float* __restrict__ b_sca, int B, int C,
                                                      float* __restrict__ s) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * C) return;
  const int b = warp / C, c = warp % C;
  const float* wr = w_sca + (size_t)c * C;
  const float* m = mean + (size_t)b * C;
  float t = 0.f;
  for (int k = lane; k < C; k += 32) t = fmaf(wr[k], m[k], t);
  t = warp_sum(t);
  if (lane == 0) s[warp] = t + (b_sca ? b_sca[c] : 0.f);
}

// Stage 3 (N:225-229 folded): Weff[b][co][ci] = rowscale[co] * W3[co][ci] * s[b][ci]  (+ transposed twin for the dgrad)
__global__ void __launch_bounds__(256) sca_fold_kernel(const float* __restrict__ s, int B, int C,
                                                       const float* __restrict__ w3, int Co,
                                                       const float* __restrict__ rowscale, bf16* __restrict__ weff,
                                                       long long weff_ld, bf16* __restrict__ weff_t,
                                                       long long weff_t_ld, int fp16) {
  const long long total = (long long)B * Co * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(i % C);
    const long long r = i / C;
    const int co = (int)(r % Co), b = (int)(r / Co);
    const float vf = w3[(size_t)co * C + ci] * s[(size_t)b * C + ci] * (rowscale ? rowscale[co] : 1.f);
    reinterpret_cast<uint16_t*>(weff)[((size_t)b * Co + co) * weff_ld + ci] = pack1r(vf, fp16);
    if (weff_t) weff_t[((size_t)b * C + ci) * weff_t_ld + co] = __float2bfloat16(vf);      // transposed copy: dgrad operand
  }
}
}  // namespace

extern "C" int tdr_gate_mul(const void* x_bf16, long long ld, long long rows, int C, void* out_bf16, long long out_ld,
                            int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(x_bf16 && out_bf16 && rows > 0 && C > 0 && C % 8 == 0 && ld % 8 == 0 && out_ld % 8 == 0,
                "tdr_gate_mul: bad arguments");
  gate_mul_kernel<<<grid_for(rows * (C / 8), 256, 16), 256, 0, stream>>>(reinterpret_cast<const bf16*>(x_bf16), ld, rows,
                                                                        C, reinterpret_cast<bf16*>(out_bf16), out_ld, fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

static int naf_pool_chunks(long long P) {
  // 128 pixels per chunk (up to 128 chunks per sample): at the coarse NAFNet levels (64 x 64 pixels, 28 blocks at C = 512)
  // 1024-pixel chunks left the pooling pass with 16 CTAs for 17 MB -- 73 us per fold, 8 % of the guided-NAFNet forward
  long long c = (P + 127) / 128;
  if (c > 128) c = 128;
  if (c < 1) c = 1;
  return (int)c;
}

extern "C" size_t tdr_naf_sca_workspace_bytes(int B, long long P, int C) {
  if (B <= 0 || P <= 0 || C <= 0) return 0;
  return ((size_t)B * naf_pool_chunks(P) * C + 2 * (size_t)B * C) * sizeof(float);   // pool partials + mean + s
}

extern "C" int tdr_naf_sca_fold(const void* g_bf16, long long ld, int B, long long P, int C, const float* w_sca,
                                const float* b_sca, const float* w3, int Co, const float* rowscale, void* weff_bf16,
                                long long weff_ld, float* workspace, float* mean_out, float* s_out, void* weff_t_bf16,
                                long long weff_t_ld, int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG((mean_out == nullptr) == (s_out == nullptr), "tdr_naf_sca_fold: mean_out and s_out go together");
  TDR_CHECK_ARG(!weff_t_bf16 || (weff_t_ld >= Co && weff_t_ld % 8 == 0), "tdr_naf_sca_fold: bad weff_t_ld");
  TDR_CHECK_ARG(g_bf16 && w_sca && w3 && weff_bf16 && workspace, "tdr_naf_sca_fold: null pointer");
  TDR_CHECK_ARG(B > 0 && P > 0 && C % 8 == 0 && ld % 8 == 0 && weff_ld >= C && weff_ld % 8 == 0 && Co > 0,
                "tdr_naf_sca_fold: bad dims");
  TDR_CHECK_ARG(C <= 6144, "tdr_naf_sca_fold: C too large");
  const int chunks = naf_pool_chunks(P);
  float* mean = mean_out ? mean_out : workspace + (size_t)B * chunks * C;
  float* svec = s_out ? s_out : workspace + (size_t)B * chunks * C + (size_t)B * C;
  dim3 g1(chunks, B);
  pool_partial_kernel<<<g1, 256, 0, stream>>>(reinterpret_cast<const bf16*>(g_bf16), ld, P, C, chunks, workspace, fp16);
  TDR_CHECK_LAUNCH();
  sca_mean_kernel<<<tdr_cdiv((long long)B * C, 256), 256, 0, stream>>>(workspace, chunks, P, C, B, mean);
  TDR_CHECK_LAUNCH();
  sca_vec_kernel<<<tdr_cdiv((long long)B * C * 32, 256), 256, 0, stream>>>(mean, w_sca, b_sca, B, C, svec);
  TDR_CHECK_LAUNCH();
  sca_fold_kernel<<<grid_for((long long)B * Co * C, 256, 8), 256, 0, stream>>>(
      svec, B, C, w3, Co, rowscale, reinterpret_cast<bf16*>(weff_bf16), weff_ld, reinterpret_cast<bf16*>(weff_t_bf16),
      weff_t_ld, fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_pack_conv_weight(const float* w, int Co, int Ci, int KH, int KW, const int* co_map, int Co_p,
                                    const int* ci_map, int Ci_p, const float* scale, void* out_bf16, long long ld,
                                    void* out_t_bf16, long long ld_t, int fwd_fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(w && (out_bf16 || out_t_bf16) && Co > 0 && Ci > 0 && KH > 0 && KW > 0 && Co_p > 0 && Ci_p > 0,
                "tdr_pack_conv_weight: bad arguments");
  TDR_CHECK_ARG((!out_bf16 || ld >= Ci_p) && (!out_t_bf16 || ld_t >= Co_p), "tdr_pack_conv_weight: row stride too small");
  const int T = KH * KW;
  const long long n = (out_bf16 ? (long long)T * Co_p * ld : 0) + (out_t_bf16 ? (long long)T * Ci_p * ld_t : 0);
  pack_conv_weight_kernel<<<grid_for(n, 256, 8), 256, 0, stream>>>(w, Co, Ci, T, co_map, Co_p, ci_map, Ci_p, scale,
                                                                    reinterpret_cast<bf16*>(out_bf16), ld,
                                                                    reinterpret_cast<bf16*>(out_t_bf16), ld_t, fwd_fp16);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_pack_dw_weight(const float* w, const float* bias, int C, const int* c_map, int C_p, float* out,
                                  float* out_flip, float* out_bias, cudaStream_t stream) {
  TDR_CHECK_ARG(w && C > 0 && C_p > 0 && (out || out_flip || out_bias), "tdr_pack_dw_weight: bad arguments");
  pack_dw_weight_kernel<<<tdr_cdiv(10 * C_p, 256), 256, 0, stream>>>(w, bias, C, c_map, C_p, out, out_flip, out_bias);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_gather_vec(const float* v, const int* map, int n, int n_src, float* out, cudaStream_t stream) {
  TDR_CHECK_ARG(v && out && n > 0 && n_src > 0, "tdr_gather_vec: bad arguments");
  gather_vec_kernel<<<tdr_cdiv(n, 256), 256, 0, stream>>>(v, map, n, n_src, out);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
