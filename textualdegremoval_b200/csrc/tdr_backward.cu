// Backward kernels of the restoration-network hot path (training step, SURVEY.md 8(a) a21 and the "bwd" half of a1-a16).
// The reference gets these from autograd over stock PyTorch ops (loss.backward() in
// /root/reference/models/image_restoration_ref_model.py:268-275); here every gradient is an explicit kernel:
//
//   tdr_wgrad            weight gradient of any dense conv / 1x1 / per-sample product: a tcgen05 GEMM whose contraction
//                        runs over PIXELS (both operands MN-major straight from NHWC TMA boxes, like tdr_mdta_gram)
//   tdr_dwconv3x3_wgrad  depthwise 3x3 weight + bias gradient (per-channel 9-tap correlation, HBM-bound)
//   tdr_colsum           bias gradients (column sums of a bf16 activation-gradient)
//   tdr_rownorm_bwd      LayerNorm (WithBias / BiasFree / LayerNorm2d) backward fused with the residual-stream add
//   tdr_gate_bwd         GELU-gate (GDFN) / SimpleGate (NAF) backward
//   tdr_mdta_bwd         softmax / normalise / temperature backward of the MDTA score path, emitted as the per-sample
//                        [2C x 2C] matrix that maps [q;k] to [dq;dk] (so the pixel-sized part is one tdr_conv_gemm)
//   tdr_scale_add_f32, tdr_dot_f32, tdr_pixel_shuffle_nhwc, tdr_relu_mask: small glue of the fusion blocks / U-Net wiring
//
// Data gradients (dgrad) of convolutions are tdr_conv_gemm / tdr_dwconv3x3 calls with transposed + flipped weights.
// Reductions are two-stage and deterministic (no atomics).
#include <stdlib.h>
#include <string.h>

#include "tdr_common.cuh"

namespace {

constexpr int kPixTile = 128;                 // pixels (K) per pipeline stage = TH x TW
constexpr int kBoxBytes = 64 * kPixTile * 2;  // one [64 ch x 128 px] bf16 box

// ------------------------------------------------------------------------------------------------ wgrad (tcgen05)
struct WgradPlan {
  int OH, OW, T;
  int TW, TH, tiles_x, tiles_y;
  int BM, BN, m_tiles, n_tiles, boxes_m, boxes_n, stages;
  int nb, nchunks, tiles_per_chunk, tiles_total;
  uint32_t tmem_cols;
  int fused;          // 1: 3x3 stride-1 conv with Ci <= 48: ONE CTA accumulates all 9 taps (9 x BN TMEM columns) from
                      //    one dy tile and three kx-shifted haloed x tiles [18 rows x 8 px]; a ky shift is 1024 B = one
                      //    swizzle atom, so the taps are descriptor offsets (dy / x are fetched 1 + 3.4 times, not 9 + 9)
  int stage_bytes;
};

constexpr int kHaloXBytes = 18 * 8 * 128;      // fused mode: one kx-shifted x tile, [18 rows][8 px][64 ch]

struct WgradArgs {
  int Co, Ci, KW, stride, pad, dil, per_sample;
  WgradPlan plan;
  float* partials;
};

static int wgrad_plan(const tdr_wgrad_desc* d, WgradPlan* p) {
  p->OH = (d->H + 2 * d->pad - d->dil * (d->KH - 1) - 1) / d->stride + 1;
  p->OW = (d->W + 2 * d->pad - d->dil * (d->KW - 1) - 1) / d->stride + 1;
  if (p->OH <= 0 || p->OW <= 0) return -1;
  p->T = d->KH * d->KW;
  if (p->OH == 1) { p->TW = 128; p->TH = 1; }          // flat [rows, C] view (ViT / mapper linears)
  else if (p->OW >= 16) { p->TW = 16; p->TH = 8; }
  else { p->TW = 8; p->TH = 16; }
  p->tiles_x = tdr_cdiv(p->OW, p->TW);
  p->tiles_y = tdr_cdiv(p->OH, p->TH);
  p->BM = d->Co <= 64 ? 64 : 128;
  p->m_tiles = tdr_cdiv(d->Co, p->BM);
  p->boxes_m = p->BM / 64;
  const int ci16 = tdr_cdiv(d->Ci, 16) * 16;
  p->n_tiles = tdr_cdiv(ci16, 256);
  p->BN = tdr_cdiv(tdr_cdiv(ci16, p->n_tiles), 16) * 16;
  p->boxes_n = tdr_cdiv(p->BN, 64);
  p->fused = (d->KH == 3 && d->KW == 3 && d->stride == 1 && d->dil == 1 && d->pad == 1 && !d->per_sample && p->OH > 1 &&
              9 * p->BN <= 512 && p->n_tiles == 1 && getenv("TDR_WGRAD_NO_FUSE") == nullptr)
                 ? 1 : 0;
  if (p->fused) {
    p->TW = 8; p->TH = 16;
    p->tiles_x = tdr_cdiv(p->OW, p->TW);
    p->tiles_y = tdr_cdiv(p->OH, p->TH);
  }
  p->stage_bytes = p->fused ? p->boxes_m * kBoxBytes + 3 * kHaloXBytes : (p->boxes_m + p->boxes_n) * kBoxBytes;
  p->stages = (220 * 1024) / p->stage_bytes;
  if (p->stages > 4) p->stages = 4;
  if (p->stages < 2) return -1;
  p->nb = d->per_sample ? d->B : 1;
  p->tiles_total = (d->per_sample ? 1 : d->B) * p->tiles_y * p->tiles_x;
  const int outer = (p->fused ? 1 : p->T) * p->m_tiles * p->n_tiles * p->nb;
  int want = tdr_num_sms() / outer;          // one wave of CTAs (1 CTA / SM: the smem ring fills the SM)
  if (want < 1) want = 1;
  if (want > p->tiles_total) want = p->tiles_total;
  p->tiles_per_chunk = tdr_cdiv(p->tiles_total, want);
  p->nchunks = tdr_cdiv(p->tiles_total, p->tiles_per_chunk);
  uint32_t cols = 32;
  while (cols < (uint32_t)(p->fused ? 9 * p->BN : p->BN)) cols <<= 1;
  p->tmem_cols = cols;
  return 0;
}

// grid ((tap * m_tiles + mt) * n_tiles + nt, chunk, sample-or-1); 6 warps: TMA producer, MMA issuer, 4 epilogue.
__global__ void __launch_bounds__(192, 1) wgrad_kernel(const __grid_constant__ TdrTensorMap map_dy,
                                                       const __grid_constant__ TdrTensorMap map_x, const WgradArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const WgradPlan& pl = a.plan;
  const int stage_bytes = pl.stage_bytes;                          // dy boxes then x boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + pl.stages;
  uint64_t* tfull = bars + 2 * pl.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * pl.stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = (tap, m tile, n tile) is the fastest-varying launch coordinate, so the CTAs that share a pixel chunk
  // (and therefore re-read the same dy / x tiles) run concurrently and hit in L2
  const int chunk = blockIdx.y, bz = blockIdx.z;
  int yy = blockIdx.x;
  const int nt = yy % pl.n_tiles; yy /= pl.n_tiles;
  const int mt = yy % pl.m_tiles;
  const int tap = yy / pl.m_tiles;                                 // 0 in fused mode (all taps in this CTA)
  const int ky = tap / a.KW, kx = tap % a.KW;
  const int tile0 = chunk * pl.tiles_per_chunk;
  int ntiles = pl.tiles_total - tile0;
  if (ntiles > pl.tiles_per_chunk) ntiles = pl.tiles_per_chunk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_x);
    for (int s = 0; s < pl.stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tfull, 1);
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, pl.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      // (tx, ty, b) of the CTA's first pixel tile by division ONCE, then advanced as counters: the single producer
      // thread used to run five runtime divisions per stage next to its TMA issues
      int tx = tile0 % pl.tiles_x, ty = (tile0 / pl.tiles_x) % pl.tiles_y;
      int b = a.per_sample ? bz : tile0 / pl.tiles_x / pl.tiles_y;
      for (int t = 0; t < ntiles; ++t) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], stage_bytes);
        uint8_t* base = smem + stage * stage_bytes;
        for (int j = 0; j < pl.boxes_m; ++j)
          tma_load_4d(base + j * kBoxBytes, &map_dy, &full[stage], mt * pl.BM + 64 * j, tx * pl.TW, ty * pl.TH, b);
        if (pl.fused) {                                            // three kx-shifted haloed x tiles (map_x box = 8 x 18)
          for (int k3 = 0; k3 < 3; ++k3)
            tma_load_4d(base + pl.boxes_m * kBoxBytes + k3 * kHaloXBytes, &map_x, &full[stage], 0, tx * pl.TW + k3 - 1,
                        ty * pl.TH - 1, b);
          if (++stage == pl.stages) { stage = 0; phase ^= 1; }
          if (++tx == pl.tiles_x) { tx = 0; if (++ty == pl.tiles_y) { ty = 0; if (!a.per_sample) ++b; } }
          continue;
        }
        const int x0 = tx * pl.TW * a.stride + kx * a.dil - a.pad;
        const int y0 = ty * pl.TH * a.stride + ky * a.dil - a.pad;
        for (int j = 0; j < pl.boxes_n; ++j)
          tma_load_4d(base + (pl.boxes_m + j) * kBoxBytes, &map_x, &full[stage], nt * pl.BN + 64 * j, x0, y0, b);
        if (++stage == pl.stages) { stage = 0; phase ^= 1; }
        if (++tx == pl.tiles_x) { tx = 0; if (++ty == pl.tiles_y) { ty = 0; if (!a.per_sample) ++b; } }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(pl.BM, pl.BN, 1, 1);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_u32(smem + stage * stage_bytes);
        const uint32_t sb = sa + pl.boxes_m * kBoxBytes;
        // descriptors are built once per stage and advanced through their start-address field (16 B units): the
        // single issuing thread is instruction-bound for these small-N MMAs
        const uint64_t da0 = umma_desc_sw128(sa, kBoxBytes, 1024);
        const uint64_t db0 = umma_desc_sw128(sb, kBoxBytes, 1024);
        if (pl.fused) {
          uint32_t d_tmem = tmem_base;
          for (int ky = 0; ky < 3; ++ky) {
            for (int kx = 0; kx < 3; ++kx) {
              // tap (ky, kx): x tile kx, shifted down by ky image rows = ky * 8 px * 128 B (one swizzle atom)
              const uint64_t db = db0 + (uint64_t)((kx * kHaloXBytes + ky * 1024) >> 4);
#pragma unroll
              for (int ks = 0; ks < kPixTile / 16; ++ks)
                umma_bf16(d_tmem, da0 + ks * 128, db + ks * 128, idesc, (t | ks) != 0);
              d_tmem += pl.BN;
            }
          }
        } else {
#pragma unroll
          for (int ks = 0; ks < kPixTile / 16; ++ks)     // 16 pixels (K) = two 8-row swizzle atoms = 2048 B = 128 units
            umma_bf16(tmem_base, da0 + ks * 128, db0 + ks * 128, idesc, (t | ks) != 0);
        }
        umma_commit(&empty[stage]);
        if (t == ntiles - 1) umma_commit(tfull);
      }
      __syncwarp();
      if (++stage == pl.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // epilogue warps 2..5: TMEM lane quadrant = warp % 4
    const int quad = warp & 3;
    // M = 128: row r lives in lane r.  M = 64: row r lives in lane (r/16)*32 + r%16 (16 rows per quadrant).
    const int row = pl.BM == 128 ? quad * 32 + lane : quad * 16 + lane;
    const int co = mt * pl.BM + row;
    const bool row_ok = (pl.BM == 128 || lane < 16) && co < a.Co;
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int ntap = pl.fused ? 9 : 1;
    for (int tp = 0; tp < ntap; ++tp) {
      float* out = a.partials + ((((size_t)bz * pl.nchunks + chunk) * pl.T + tap + tp) * a.Co + co) * (size_t)a.Ci;
      for (int c16 = 0; c16 < pl.BN / 16; ++c16) {
        uint32_t g[16];
        tmem_ld16(t_lane + tp * pl.BN + c16 * 16, g);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int ci = nt * pl.BN + c16 * 16 + i;
            if (ci < a.Ci) out[ci] = __uint_as_float(g[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, pl.tmem_cols);
  }
}

// out[b*s_b + map(co)*s_co + map(ci)*s_ci + t*s_t] (+)= scale * sum_chunk partials[b][chunk][t][co][ci]
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partials, int nb, int nchunks, int T,
                                                           int Co, int Ci, float* __restrict__ out, long long s_b,
                                                           long long s_co, long long s_ci, long long s_t,
                                                           const int* __restrict__ co_map, const int* __restrict__ ci_map,
                                                           int accumulate, float scale, const float* __restrict__ scale_ptr) {
  if (scale_ptr) scale *= *scale_ptr;                  // device-side factor (MASA encoder level scale)
  const long long per = (long long)T * Co * Ci;
  const long long total = (long long)nb * per;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / per);
    long long r = i % per;
    const int ci = (int)(r % Ci); r /= Ci;
    const int co = (int)(r % Co);
    const int t = (int)(r / Co);
    const int lco = co_map ? co_map[co] : co;
    const int lci = ci_map ? ci_map[ci] : ci;
    if (lco < 0 || lci < 0) continue;
    const float* p = partials + (size_t)b * nchunks * per + (i % per);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;      // independent chains: the chunk loop is L2-latency bound
    int c = 0;
    for (; c + 3 < nchunks; c += 4) {
      s0 += p[(size_t)c * per];
      s1 += p[(size_t)(c + 1) * per];
      s2 += p[(size_t)(c + 2) * per];
      s3 += p[(size_t)(c + 3) * per];
    }
    for (; c < nchunks; ++c) s0 += p[(size_t)c * per];
    const float s = (s0 + s1) + (s2 + s3);
    float* o = out + b * s_b + lco * s_co + lci * s_ci + t * s_t;
    *o = (accumulate ? *o : 0.f) + scale * s;
  }
}

// ------------------------------------------------------------------------------------------------ generic helpers
inline int grid_for(long long items, int per_block, int max_waves = 8) {
  long long g = (items + per_block - 1) / per_block;
  const long long cap = (long long)tdr_num_sms() * max_waves;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// out[map(c)*stride] (+)= scale * sum_{p < nparts} partials[p][c]
__global__ void __launch_bounds__(256) reduce_parts_kernel(const float* __restrict__ partials, int nparts, int n,
                                                           float* __restrict__ out, long long stride,
                                                           const int* __restrict__ map, int accumulate, float scale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int lc = map ? map[c] : c;
  if (lc < 0) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;        // fixed association order: still deterministic
  int p = 0;
  for (; p + 3 < nparts; p += 4) {
    s0 += partials[(size_t)p * n + c];
    s1 += partials[(size_t)(p + 1) * n + c];
    s2 += partials[(size_t)(p + 2) * n + c];
    s3 += partials[(size_t)(p + 3) * n + c];
  }
  for (; p < nparts; ++p) s0 += partials[(size_t)p * n + c];
  const float s = (s0 + s1) + (s2 + s3);
  float* o = out + lc * stride;
  *o = (accumulate ? *o : 0.f) + scale * s;
}

// out[map(c)*stride] (+)= sum_p partials[p][c] with the partial rows split over the 8 warps of a block (many partial rows,
// few columns: colsum of small activations).  grid (ceil(n/32)), block (32, 8); fixed summation order.
__global__ void __launch_bounds__(256) reduce_parts_wide_kernel(const float* __restrict__ partials, int nparts, int n,
                                                                float* __restrict__ out, long long stride,
                                                                const int* __restrict__ map, int accumulate) {
  __shared__ float red[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  float s = 0.f;
  if (c < n)
    for (int p = ty; p < nparts; p += 8) s += partials[(size_t)p * n + c];
  red[ty][threadIdx.x] = s;
  __syncthreads();
  if (ty == 0 && c < n) {
    const int lc = map ? map[c] : c;
    if (lc < 0) return;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    float* o = out + lc * stride;
    *o = (accumulate ? *o : 0.f) + t;
  }
}

constexpr int kRedBlocks = 296;   // partial rows written by the first stage of the column reductions

// ------------------------------------------------------------------------------------------------ colsum / dw wgrad
// Thread mapping shared by the two column reductions: CG = C/8 channel groups; a block of 256 threads covers
// cgb = min(CG, 256) groups x (256 / cgb) pixel slots (thread = slot * cgb + group), so narrow tensors still use every
// lane and a warp's loads are contiguous runs.  grid (blocks over pixels, ceil(CG / 256)).
struct ColMap {
  int cgb, slots, slot, c0;
  bool active;
};
__device__ __forceinline__ ColMap col_map(int C) {
  ColMap m;
  const int CG = C >> 3;
  m.cgb = CG < 256 ? CG : 256;
  m.slots = 256 / m.cgb;
  m.slot = threadIdx.x / m.cgb;
  const int cg = blockIdx.y * 256 + (int)(threadIdx.x % m.cgb);
  m.c0 = cg * 8;
  m.active = m.slot < m.slots && cg < CG;
  return m;
}
// sums sm[slot][g][i] over slots and writes dst[c] for the block's channel range
__device__ __forceinline__ void col_reduce_store(const float (*sm)[8], const ColMap& m, int C, float* dst) {
  for (int e = threadIdx.x; e < m.cgb * 8; e += 256) {
    const int g = e >> 3, i = e & 7;
    float s = 0.f;
    for (int k = 0; k < m.slots; ++k) s += sm[k * m.cgb + g][i];
    const int c = (blockIdx.y * 256 + g) * 8 + i;
    if (c < C) dst[c] = s;
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const bf16* __restrict__ x, long long ld, long long rows, int C,
                                                     float* __restrict__ partials) {
  __shared__ float sm[256][8];
  const ColMap m = col_map(C);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (m.active) {
    for (long long r = (long long)blockIdx.x * m.slots + m.slot; r < rows; r += (long long)gridDim.x * m.slots) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(x + r * ld + m.c0), f);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += f[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm[threadIdx.x][i] = acc[i];
  __syncthreads();
  col_reduce_store(sm, m, C, partials + (size_t)blockIdx.x * C);
}

// ---- depthwise 3x3 weight gradient ----------------------------------------------------------------------------------
// dW[tap][c] = sum_p dy[p, c] * x[p + off(tap), c];  db[c] = sum_p dy[p, c].  partials [blk][10][C] (tap 9 = bias).
// blockIdx.y = block of 128 channels.  A lane owns 4 channels (8 B loads; a warp row = up to 256 B contiguous) and walks
// a 32-pixel segment of one image row with a 3x3 register window: per pixel 3 new x vectors + 1 dy vector are loaded and
// 9 taps x 4 channels accumulate in packed fp32x2 FMAs.  Narrow channel blocks (< 32 quads) put several x sub-segments
// in one warp.  The 8 warps of a block take 8 consecutive rows, so the rows above / below are L1 hits.
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk2(float a, float b) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f2 r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ f2 bf2_to_f2(uint32_t u) { return pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

struct Q4 {   // 4 channels as two packed pairs
  f2 a, b;
};
__device__ __forceinline__ Q4 ldq(const bf16* p, bool ok) {
  Q4 q;
  if (ok) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    q.a = bf2_to_f2(u.x);
    q.b = bf2_to_f2(u.y);
  } else {
    q.a = q.b = 0ull;
  }
  return q;
}

// TMA version: a persistent CTA owns one 64-channel chunk and streams [8 x 32]-pixel tiles: the haloed x tile
// [10 x 34 x 64ch] and the dy tile [8 x 32 x 64ch] land in shared memory through a 2-stage mbarrier ring (76 KB per
// stage in flight per SM, independent of occupancy); out-of-image pixels are zero-filled by TMA, which is exactly the
// conv's zero padding (x) and "no contribution" (dy), so the inner loop has no boundary logic.  A thread owns 4 channels
// and 2 pixel columns and walks down the 10 staged rows keeping 3 dy rows in registers: 4 LDS.64 per 18 packed FMAs.
constexpr int kDgRows = 8, kDgCols = 32;
constexpr int kDgXBytes = (kDgRows + 2) * (kDgCols + 2) * 128;     // 43520
constexpr int kDgYBytes = kDgRows * kDgCols * 128;                 // 32768
constexpr int kDgStageBytes = 76544;                               // x box + dy box, rounded to 256 B

struct DwgArgs {
  int B, H, W, C;
  int tiles_x, tiles_y, chunks;
  float* partials;
};

__device__ __forceinline__ Q4 ldsq(const uint8_t* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  Q4 q;
  q.a = bf2_to_f2(u.x);
  q.b = bf2_to_f2(u.y);
  return q;
}

__global__ void __launch_bounds__(256, 1) dw_wgrad_tma_kernel(const __grid_constant__ TdrTensorMap map_x,
                                                              const __grid_constant__ TdrTensorMap map_dy,
                                                              const DwgArgs a) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) uint64_t full[2];
  __shared__ float red[256][4];
  const int tid = threadIdx.x;
  const int cq = tid & 15, xl = tid >> 4;                  // channel quad within the chunk, pixel column (and +16)
  if (tid == 0) {
    tma_prefetch_desc(&map_x);
    tma_prefetch_desc(&map_dy);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  const int cc = blockIdx.x % a.chunks;
  const int grp = blockIdx.x / a.chunks, ngrp = gridDim.x / a.chunks;
  const int n_spatial = a.B * a.tiles_y * a.tiles_x;
  auto issue = [&](int sp, int stage) {                    // thread 0 only
    int r = sp;
    const int b = r / (a.tiles_y * a.tiles_x);
    r %= a.tiles_y * a.tiles_x;
    const int y0 = (r / a.tiles_x) * kDgRows, x0 = (r % a.tiles_x) * kDgCols;
    uint8_t* dst = dsm + stage * kDgStageBytes;
    mbar_expect_tx(&full[stage], kDgXBytes + kDgYBytes);
    tma_load_4d(dst, &map_x, &full[stage], cc * 64, x0 - 1, y0 - 1, b);
    tma_load_4d(dst + kDgXBytes, &map_dy, &full[stage], cc * 64, x0, y0, b);
  };
  Q4 acc[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) acc[t].a = acc[t].b = 0ull;
  int it = 0;
  if (tid == 0 && grp < n_spatial) issue(grp, 0);
  for (int sp = grp; sp < n_spatial; sp += ngrp, ++it) {
    const int stage = it & 1;
    if (tid == 0 && sp + ngrp < n_spatial) issue(sp + ngrp, stage ^ 1);
    mbar_wait(&full[stage], (it >> 1) & 1);
    const uint8_t* xs = dsm + stage * kDgStageBytes + cq * 8;
    const uint8_t* ys = xs + kDgXBytes;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int col = xl + 16 * half;                      // output column within the tile; x columns col .. col+2
      Q4 g0, g1, g2;                                       // dy rows i-2, i-1, i  (taps ky = 2, 1, 0 of staged row i)
      g0.a = g0.b = g1.a = g1.b = g2.a = g2.b = 0ull;
#pragma unroll
      for (int i = 0; i < kDgRows + 2; ++i) {              // staged x row i = image row y0 - 1 + i
        g0 = g1;
        g1 = g2;
        if (i < kDgRows) {
          g2 = ldsq(ys + (i * kDgCols + col) * 128);
          acc[9].a = add2(acc[9].a, g2.a);
          acc[9].b = add2(acc[9].b, g2.b);
        } else {
          g2.a = g2.b = 0ull;
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const Q4 v = ldsq(xs + (i * (kDgCols + 2) + col + kx) * 128);
          // x row i is tap ky of output row (i - ky): dy rows i (ky 0) = g2, i-1 (ky 1) = g1, i-2 (ky 2) = g0
          acc[0 * 3 + kx].a = fma2(g2.a, v.a, acc[0 * 3 + kx].a);
          acc[0 * 3 + kx].b = fma2(g2.b, v.b, acc[0 * 3 + kx].b);
          acc[1 * 3 + kx].a = fma2(g1.a, v.a, acc[1 * 3 + kx].a);
          acc[1 * 3 + kx].b = fma2(g1.b, v.b, acc[1 * 3 + kx].b);
          acc[2 * 3 + kx].a = fma2(g0.a, v.a, acc[2 * 3 + kx].a);
          acc[2 * 3 + kx].b = fma2(g0.b, v.b, acc[2 * 3 + kx].b);
        }
      }
    }
    __syncthreads();                                       // everyone is done with this stage before it is refilled
  }
  // reduce the 16 column-threads of every channel quad
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    upk2(acc[t].a, red[tid][0], red[tid][1]);
    upk2(acc[t].b, red[tid][2], red[tid][3]);
    __syncthreads();
    if (tid < 64) {
      const int q = tid >> 2, i = tid & 3;
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) sum += red[k * 16 + q][i];
      const int c = cc * 64 + q * 4 + i;
      if (c < a.C) a.partials[((size_t)grp * 10 + t) * a.C + c] = sum;
    }
    __syncthreads();
  }
}

// dw[map(c)*9 + tap] (+)= sum_blk partials[blk][tap][c];  db[map(c)] (+)= sum_blk partials[blk][9][c]
__global__ void __launch_bounds__(256) dw_wgrad_reduce_kernel(const float* __restrict__ partials, int nparts, int C,
                                                              float* __restrict__ dw, float* __restrict__ db,
                                                              const int* __restrict__ map, int accumulate) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 10 * C) return;
  const int t = idx / C, c = idx % C;
  const int lc = map ? map[c] : c;
  if (lc < 0) return;
  if (t == 9 && !db) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int p = 0;
  for (; p + 3 < nparts; p += 4) {
    s0 += partials[((size_t)p * 10 + t) * C + c];
    s1 += partials[((size_t)(p + 1) * 10 + t) * C + c];
    s2 += partials[((size_t)(p + 2) * 10 + t) * C + c];
    s3 += partials[((size_t)(p + 3) * 10 + t) * C + c];
  }
  for (; p < nparts; ++p) s0 += partials[((size_t)p * 10 + t) * C + c];
  const float s = (s0 + s1) + (s2 + s3);
  float* o = t == 9 ? db + lc : dw + lc * 9 + t;
  *o = (accumulate ? *o : 0.f) + s;
}

// ------------------------------------------------------------------------------------------------ rownorm backward
// G lanes per row, NV float4 per lane (same mapping as the forward kernel).  dx = add + LN_bwd(dy) ; per-block partial
// sums of dweight / dbias are reduced through shared memory in a fixed order.
template <int NV>
__global__ void __launch_bounds__(256, 3) rownorm_bwd_kernel(const float* __restrict__ x, long long x_ld,
                                                          const bf16* __restrict__ dy, long long dy_ld, long long rows,
                                                          int C, int mode, const float* __restrict__ w, float eps,
                                                          const float* __restrict__ add, long long add_ld,
                                                          float* __restrict__ dx, long long dx_ld,
                                                          bf16* __restrict__ dx16, long long dx16_ld,
                                                          float* __restrict__ partials, int G) {
  extern __shared__ float sm[];                // [2][256/G][C]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rpw = 32 / G;
  const int sub = lane / G, l = lane % G;
  const int nvec = C >> 2;
  const int slot = warp * rpw + sub;           // row slot inside the block
  const int slots = 8 * rpw;
  float4 aw[NV], ab[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) aw[i] = ab[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long row0 = (long long)blockIdx.x * slots; row0 < rows; row0 += (long long)gridDim.x * slots) {
    const long long row = row0 + slot;
    const bool row_ok = row < rows;
    float4 v[NV], g[NV], ad[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int idx = l + i * G;
      ad[i] = (add && row_ok && idx < nvec) ? *reinterpret_cast<const float4*>(add + row * add_ld + idx * 4)
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
      if (row_ok && idx < nvec) {
        v[i] = mode ? *reinterpret_cast<const float4*>(x + row * x_ld + idx * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        const uint2 pk = *reinterpret_cast<const uint2*>(dy + row * dy_ld + idx * 4);
        g[i] = make_float4(__uint_as_float(pk.x << 16), __uint_as_float(pk.x & 0xffff0000u),
                           __uint_as_float(pk.y << 16), __uint_as_float(pk.y & 0xffff0000u));
      } else {
        v[i] = g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      s += v[i].x + v[i].y + v[i].z + v[i].w;
    }
    float mean = 0.f, rstd = 1.f, m1 = 0.f, m2 = 0.f;
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
      // xh = normalised value the forward multiplied by w: (x - mean) * rstd [mode 1] or x * rstd [mode 2]
      // gw = dy * w;  m1 = mean(gw);  m2 = mean(gw * (x - mean) * rstd) [mode 1] / mean(gw * x) [mode 2]
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int idx = l + i * G;
        if (idx < nvec) {
          const float4 ww = *reinterpret_cast<const float4*>(w + idx * 4);
          const float4 gw = make_float4(g[i].x * ww.x, g[i].y * ww.y, g[i].z * ww.z, g[i].w * ww.w);
          float4 xh;
          if (mode == 1) xh = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd, (v[i].w - mean) * rstd);
          else xh = make_float4(v[i].x * rstd, v[i].y * rstd, v[i].z * rstd, v[i].w * rstd);
          if (row_ok) {
            aw[i].x += g[i].x * xh.x; aw[i].y += g[i].y * xh.y; aw[i].z += g[i].z * xh.z; aw[i].w += g[i].w * xh.w;
            ab[i].x += g[i].x; ab[i].y += g[i].y; ab[i].z += g[i].z; ab[i].w += g[i].w;
          }
          m1 += gw.x + gw.y + gw.z + gw.w;
          if (mode == 1) m2 += gw.x * xh.x + gw.y * xh.y + gw.z * xh.z + gw.w * xh.w;
          else m2 += gw.x * v[i].x + gw.y * v[i].y + gw.z * v[i].z + gw.w * v[i].w;
          g[i] = gw;
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) {
        m1 += __shfl_xor_sync(0xffffffffu, m1, o);
        m2 += __shfl_xor_sync(0xffffffffu, m2, o);
      }
      m1 /= (float)C;
      m2 /= (float)C;
    }
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int idx = l + i * G;
        if (idx < nvec) {
          float4 r;
          if (mode == 1) {          // dx = rstd * (gw - mean(gw) - xh * mean(gw * xh))
            r.x = rstd * (g[i].x - m1 - (v[i].x - mean) * rstd * m2);
            r.y = rstd * (g[i].y - m1 - (v[i].y - mean) * rstd * m2);
            r.z = rstd * (g[i].z - m1 - (v[i].z - mean) * rstd * m2);
            r.w = rstd * (g[i].w - m1 - (v[i].w - mean) * rstd * m2);
          } else if (mode == 2) {   // y = x * rstd * w : dx = rstd * gw - rstd^3 * (x - mean) * mean(gw * x)
            const float r3 = rstd * rstd * rstd * m2;
            r.x = rstd * g[i].x - r3 * (v[i].x - mean);
            r.y = rstd * g[i].y - r3 * (v[i].y - mean);
            r.z = rstd * g[i].z - r3 * (v[i].z - mean);
            r.w = rstd * g[i].w - r3 * (v[i].w - mean);
          } else {
            r = g[i];
          }
          r.x += ad[i].x; r.y += ad[i].y; r.z += ad[i].z; r.w += ad[i].w;
          *reinterpret_cast<float4*>(dx + row * dx_ld + idx * 4) = r;
          if (dx16) {                 // bf16 copy: the operand of the next dgrad / wgrad GEMMs
            uint2 pk;
            pk.x = pack2(r.x, r.y);
            pk.y = pack2(r.z, r.w);
            *reinterpret_cast<uint2*>(dx16 + row * dx16_ld + idx * 4) = pk;
          }
        }
      }
    }
  }
  if (!partials) return;
  // block reduction of dweight / dbias in slot order
  float* sw = sm;
  float* sb = sm + (size_t)slots * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int idx = l + i * G;
    if (idx < nvec) {
      *reinterpret_cast<float4*>(sw + (size_t)slot * C + idx * 4) = aw[i];
      *reinterpret_cast<float4*>(sb + (size_t)slot * C + idx * 4) = ab[i];
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += blockDim.x) {
    const float* src = c < C ? sw + c : sb + (c - C);
    float s = 0.f;
    for (int k = 0; k < slots; ++k) s += src[(size_t)k * C];
    partials[(size_t)blockIdx.x * 2 * C + c] = s;
  }
}

// dweight[c] (+)= sum_blk partials[blk][c];  dbias[c] (+)= sum_blk partials[blk][C + c].  grid (ceil(2C/32)), block (32, 8):
// the partial rows (up to 444) are split over the 8 warps; fixed order.
__global__ void __launch_bounds__(256) rownorm_bwd_reduce_kernel(const float* __restrict__ partials, int nblk, int C,
                                                                 float* __restrict__ dweight, float* __restrict__ dbias,
                                                                 int accumulate) {
  __shared__ float red[8][32];
  const int idx = blockIdx.x * 32 + threadIdx.x, ty = threadIdx.y;
  float s = 0.f;
  if (idx < 2 * C)
    for (int p = ty; p < nblk; p += 8) s += partials[(size_t)p * 2 * C + idx];
  red[ty][threadIdx.x] = s;
  __syncthreads();
  if (ty != 0 || idx >= 2 * C) return;
  float* o = idx < C ? dweight + idx : (dbias ? dbias + (idx - C) : nullptr);
  if (!o) return;
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
  *o = (accumulate ? *o : 0.f) + t;
}

// ------------------------------------------------------------------------------------------------ gate backward
__device__ __forceinline__ void gelu_and_grad(float a, float& g, float& dg) {
  const float cdf = 0.5f * (1.f + erff(a * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * a * a);
  g = a * cdf;
  dg = cdf + a * pdf;
}

// y: [rows, 2*Ch] pre-gate (a | b halves), dg: [rows, Ch].  dyo[:, :Ch] = dg * b * act'(a), dyo[:, Ch:] = dg * act(a)
// with act = GELU (gate 1) or identity (gate 2: SimpleGate, dyo = [dg*b | dg*a]).  dyo may alias y.
__global__ void __launch_bounds__(256) gate_bwd_kernel(const bf16* __restrict__ y, long long y_ld,
                                                       const bf16* __restrict__ dg, long long dg_ld, long long rows,
                                                       int Ch, int gate, bf16* __restrict__ dyo, long long dyo_ld,
                                                       const float* __restrict__ dg_add, long long rows_per_sample) {
  const int nv = Ch >> 3;
  const long long total = rows * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i % nv) * 8;
    float a[8], b[8], g[8], oa[8], ob[8];
    unpack8(*reinterpret_cast<const bf16x8*>(y + r * y_ld + c), a);
    unpack8(*reinterpret_cast<const bf16x8*>(y + r * y_ld + Ch + c), b);
    unpack8(*reinterpret_cast<const bf16x8*>(dg + r * dg_ld + c), g);
    if (dg_add) {                     // per-sample channel term (gradient through the SCA average pool N:192-196)
      const float* ad = dg_add + (r / rows_per_sample) * Ch + c;
#pragma unroll
      for (int k = 0; k < 8; ++k) g[k] += ad[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (gate == 1) {
        float ga, dga;
        gelu_and_grad(a[k], ga, dga);
        oa[k] = g[k] * b[k] * dga;
        ob[k] = g[k] * ga;
      } else {
        oa[k] = g[k] * b[k];
        ob[k] = g[k] * a[k];
      }
    }
    *reinterpret_cast<bf16x8*>(dyo + r * dyo_ld + c) = pack8(oa);
    *reinterpret_cast<bf16x8*>(dyo + r * dyo_ld + Ch + c) = pack8(ob);
  }
}

// ------------------------------------------------------------------------------------------------ MDTA backward (small)
// Strided, two-level batched fp32 SGEMM (SIMT, 64x64 tiles, 4x4 per thread): C(m,n) = sum_k A(m,k) * B(k,n).
// Used for the per-(sample, head) products of the MDTA backward, which are far too small for a tensor-core launch.
struct SgemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K, nb2;
  long long sam, sak, sbk, sbn, scm, scn;
  long long a_b1, a_b2, b_b1, b_b2, c_b1, c_b2;
};

__global__ void __launch_bounds__(256) sgemm_batched_kernel(const SgemmArgs a) {
  // K tile of 32 with the next tile's global loads in flight (registers) while the current one is multiplied: these
  // products are a handful of CTAs walking K = C serially, i.e. bound by the load latency per K step.
  constexpr int KT = 32, NL = KT * 64 / 256;
  __shared__ float As[KT][65];
  __shared__ float Bs[KT][65];
  const int b1 = blockIdx.z / a.nb2, b2 = blockIdx.z % a.nb2;
  const float* A = a.A + b1 * a.a_b1 + b2 * a.a_b2;
  const float* Bm = a.B + b1 * a.b_b1 + b2 * a.b_b2;
  float* Cm = a.C + b1 * a.c_b1 + b2 * a.c_b2;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[NL], rb[NL];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int e = threadIdx.x + l * 256;
      int kk, mm;
      if (a.sam == 1) { mm = e & 63; kk = e >> 6; } else { kk = e & (KT - 1); mm = e / KT; }   // coalesce along the unit stride
      const int m = m0 + mm, k = k0 + kk;
      ra[l] = (m < a.M && k < a.K) ? A[m * a.sam + k * a.sak] : 0.f;
      int nn;
      if (a.sbn == 1) { nn = e & 63; kk = e >> 6; } else { kk = e & (KT - 1); nn = e / KT; }
      const int n = n0 + nn, k2 = k0 + kk;
      rb[l] = (n < a.N && k2 < a.K) ? Bm[k2 * a.sbk + n * a.sbn] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < a.K; k0 += KT) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int e = threadIdx.x + l * 256;
      int kk, mm, nn;
      if (a.sam == 1) { mm = e & 63; kk = e >> 6; } else { kk = e & (KT - 1); mm = e / KT; }
      As[kk][mm] = ra[l];
      if (a.sbn == 1) { nn = e & 63; kk = e >> 6; } else { kk = e & (KT - 1); nn = e / KT; }
      Bs[kk][nn] = rb[l];
    }
    __syncthreads();
    if (k0 + KT < a.K) fetch(k0 + KT);
#pragma unroll
    for (int kk = 0; kk < KT; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { av[i] = As[kk][ty + 16 * i]; bv[i] = Bs[kk][tx + 16 * i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + ty + 16 * i, n = n0 + tx + 16 * j;
      if (m < a.M && n < a.N) Cm[m * a.scm + n * a.scn] = acc[i][j];
    }
}

// One CTA per (head, sample): softmax / temperature / normalise backward and the rows of the [2C x 2C] matrix.
// Shared: attn, dattn (-> dS_hat), shat as [c][c+1] (padded against bank conflicts), nq, nk, r, s.
struct MdtaBwdArgs {
  int C, heads;
  const float* shat;       // forward: normalised Gram | |q| | |k| per (sample, head) (tdr_mdta_weff shat_out)
  const float* attn;       // [B, heads, c, c]
  const float* dattn;      // [B, heads, c, c] = W_out[:, head]^T . dWeff[b][:, head]
  const float* temperature;
  bf16* mqk;               // [B][2C][mqk_ld]
  long long mqk_ld;
  float* dtemp_part;       // [B][heads]
};

__global__ void __launch_bounds__(256) mdta_bwd_kernel(const MdtaBwdArgs a) {
  extern __shared__ float sm[];
  const int C = a.C, c = C / a.heads, ld = c + 1;
  const int h = blockIdx.x, b = blockIdx.y;
  float* attn = sm;
  float* dat = attn + c * ld;
  float* shat = dat + c * ld;
  float* nq = shat + c * ld;
  float* nk = nq + c;
  float* rr = nk + c;
  float* ss = rr + c;
  __shared__ float red[256];
  const int tid = threadIdx.x;
  const size_t psz = (size_t)c * c + 2 * c;
  const float* ssrc = a.shat + (size_t)(b * a.heads + h) * psz;
  const float* asrc = a.attn + (size_t)(b * a.heads + h) * c * c;
  const float* dsrc = a.dattn + (size_t)(b * a.heads + h) * c * c;
  const float temp = a.temperature[h];
  for (int i = tid; i < 2 * c; i += 256) nq[i] = ssrc[(size_t)c * c + i];      // nq then nk (contiguous)
  for (int t = tid; t < c * c; t += 256) {
    const int i = t / c, j = t % c;
    shat[i * ld + j] = ssrc[t];
    attn[i * ld + j] = asrc[t];
    dat[i * ld + j] = dsrc[t];
  }
  __syncthreads();
  // softmax backward per row i: dS = attn * (dattn - sum_j dattn*attn); logits = shat * temp
  float dtemp = 0.f;
  for (int i = tid >> 5; i < c; i += 8) {
    const int lane = tid & 31;
    float s = 0.f;
    for (int j = lane; j < c; j += 32) s += dat[i * ld + j] * attn[i * ld + j];
    s = warp_sum(s);
    float r = 0.f;
    for (int j = lane; j < c; j += 32) {
      const float ds = attn[i * ld + j] * (dat[i * ld + j] - s);
      dtemp += ds * shat[i * ld + j];
      const float dsh = ds * temp;
      dat[i * ld + j] = dsh;                         // dat now holds dS_hat
      r += dsh * shat[i * ld + j];
    }
    r = warp_sum(r);
    if (lane == 0) rr[i] = r;
  }
  red[tid] = dtemp;
  __syncthreads();
  if (tid < 32) {
    float s = 0.f;
    for (int k = tid; k < 256; k += 32) s += red[k];
    s = warp_sum(s);
    if (tid == 0 && blockIdx.z == 0) a.dtemp_part[b * a.heads + h] = s;
  }
  for (int j = tid; j < c; j += 256) {
    float s = 0.f;
    for (int i = 0; i < c; ++i) s += dat[i * ld + j] * shat[i * ld + j];
    ss[j] = s;
  }
  __syncthreads();
  // rows of M (zero outside this head's blocks):
  //   dq_i = sum_j dS_hat[i][j] / (nq_i nk_j) * k_j - r_i / nq_i^2 * q_i
  //   dk_j = sum_i dS_hat[i][j] / (nq_i nk_j) * q_i - s_j / nk_j^2 * k_j
  // The (sample, head) is shared by gridDim.z CTAs: every one repeats the small c x c part above (bit-identical) and writes
  // the rows i = blockIdx.z, blockIdx.z + gridDim.z, ... -- one CTA per (sample, head) left the [2C x 2C] writer on
  // heads * B SMs.
  bf16* mq = a.mqk + (size_t)b * 2 * C * a.mqk_ld;
  for (int i = blockIdx.z; i < c; i += gridDim.z) {
    for (int col = tid; col < 2 * C; col += 256) {
      float vq = 0.f, vk = 0.f;
      if (col >= C + h * c && col < C + h * c + c) {          // k columns
        const int j = col - C - h * c;
        vq = dat[i * ld + j] / (nq[i] * nk[j]);
        if (j == i) vk = -ss[i] / (nk[i] * nk[i]);
      } else if (col >= h * c && col < h * c + c) {           // q columns
        const int j = col - h * c;
        vk = dat[j * ld + i] / (nq[j] * nk[i]);
        if (j == i) vq = -rr[i] / (nq[i] * nq[i]);
      }
      mq[(size_t)(h * c + i) * a.mqk_ld + col] = __float2bfloat16(vq);
      mq[(size_t)(C + h * c + i) * a.mqk_ld + col] = __float2bfloat16(vk);
    }
  }
}

// ------------------------------------------------------------------------------------------------ small glue
// out[r, c] = s * x[r, c] + y[r, c]   (s = *scale_ptr * scale; y optional) -- TransformerResFusionBlock R:353
__global__ void __launch_bounds__(256) scale_add_kernel(const float* __restrict__ x, long long x_ld,
                                                        const float* __restrict__ y, long long y_ld, long long rows,
                                                        int C, const float* __restrict__ scale_ptr, float scale,
                                                        float* __restrict__ out, long long out_ld) {
  const int nv = C >> 2;
  const float s = (scale_ptr ? *scale_ptr : 1.f) * scale;
  const long long total = rows * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i % nv) * 4;
    float4 v = *reinterpret_cast<const float4*>(x + r * x_ld + c);
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    if (y) {
      const float4 w = *reinterpret_cast<const float4*>(y + r * y_ld + c);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    *reinterpret_cast<float4*>(out + r * out_ld + c) = v;
  }
}

// partial[blk] = sum over this block's elements of x * y
__global__ void __launch_bounds__(256) dot_kernel(const float* __restrict__ x, long long x_ld, const float* __restrict__ y,
                                                  long long y_ld, long long rows, int C, float* __restrict__ partials) {
  __shared__ float red[8];
  const int nv = C >> 2;
  const long long total = rows * nv;
  float s = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i % nv) * 4;
    const float4 a = *reinterpret_cast<const float4*>(x + r * x_ld + c);
    const float4 b = *reinterpret_cast<const float4*>(y + r * y_ld + c);
    s += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    partials[blockIdx.x] = t;
  }
}

// mode 1 (PixelUnshuffle(2)): out[b, y/2, x/2, c*4 + (y&1)*2 + (x&1)] = in[b, y, x, c]
// mode 2 (PixelShuffle(2))  : out[b, 2y + s/2, 2x + s%2, c/4] = in[b, y, x, c], s = c & 3      (bf16 -> bf16)
__global__ void __launch_bounds__(256) pixel_shuffle_kernel(const bf16* __restrict__ in, long long in_ld, int B, int H,
                                                            int W, int C, int mode, bf16* __restrict__ out,
                                                            long long out_ld) {
  const long long total = (long long)B * H * W * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long p = i / C;
    const int x = (int)(p % W); p /= W;
    const int y = (int)(p % H);
    const int b = (int)(p / H);
    const bf16 v = in[((((long long)b * H + y) * W) + x) * in_ld + c];
    if (mode == 1) {
      out[((((long long)b * (H >> 1) + (y >> 1)) * (W >> 1)) + (x >> 1)) * out_ld + c * 4 + ((y & 1) << 1) + (x & 1)] = v;
    } else {
      const int s = c & 3;
      out[((((long long)b * (H * 2) + y * 2 + (s >> 1)) * (W * 2)) + x * 2 + (s & 1)) * out_ld + (c >> 2)] = v;
    }
  }
}

// Vector forms of the two modes (the element-wise kernel above ran at 0.10 of HBM: four 64-bit divisions and one 2-byte
// scattered store per element).  Unshuffle: a thread owns one OUTPUT pixel and 8 input channels -- four 16-byte loads
// (the 2x2 input pixels), interleaved in registers, four 16-byte stores (32 consecutive output channels).
__global__ void __launch_bounds__(256) pixel_unshuffle_vec_kernel(const uint16_t* __restrict__ in, long long in_ld, int B,
                                                                  int H, int W, int C, uint16_t* __restrict__ out,
                                                                  long long out_ld) {
  const int nv = C >> 3, OW = W >> 1, OH = H >> 1;
  const unsigned total = (unsigned)B * OH * OW * nv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned v = i % nv, p = i / nv;
    const unsigned x = p % OW, t = p / OW, y = t % OH, b = t / OH;
    uint4 r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q)
      r[q] = *reinterpret_cast<const uint4*>(in + (((long long)b * H + 2 * y + (q >> 1)) * W + 2 * x + (q & 1)) * in_ld + v * 8);
    const uint16_t* e = reinterpret_cast<const uint16_t*>(r);            // e[q * 8 + c]
    uint16_t o[32];
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q) o[c * 4 + q] = e[q * 8 + c];
    uint4* dst = reinterpret_cast<uint4*>(out + (((long long)b * OH + y) * OW + x) * out_ld + v * 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) dst[k] = reinterpret_cast<const uint4*>(o)[k];
  }
}

// Shuffle: a thread owns one INPUT pixel and 32 input channels -- four 16-byte loads, de-interleaved into the four
// sub-pixels' 8 consecutive output channels, four 16-byte stores.
__global__ void __launch_bounds__(256) pixel_shuffle_vec_kernel(const uint16_t* __restrict__ in, long long in_ld, int B,
                                                                int H, int W, int C, uint16_t* __restrict__ out,
                                                                long long out_ld) {
  const int nv = C >> 5;
  const unsigned total = (unsigned)B * H * W * nv;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const unsigned v = i % nv, p = i / nv;
    const unsigned x = p % W, t = p / W, y = t % H, b = t / H;
    uint4 r[4];
    const uint4* src = reinterpret_cast<const uint4*>(in + (long long)p * in_ld + v * 32);
#pragma unroll
    for (int k = 0; k < 4; ++k) r[k] = src[k];
    const uint16_t* e = reinterpret_cast<const uint16_t*>(r);            // e[co * 4 + s]
#pragma unroll
    for (int sp = 0; sp < 4; ++sp) {
      uint16_t o[8];
#pragma unroll
      for (int co = 0; co < 8; ++co) o[co] = e[co * 4 + sp];
      *reinterpret_cast<uint4*>(out + (((long long)b * (2 * H) + 2 * y + (sp >> 1)) * (2 * W) + 2 * x + (sp & 1)) * out_ld + v * 8) =
          *reinterpret_cast<const uint4*>(o);
    }
  }
}

// dy_out = y > 0 ? dy : 0   (ReLU backward from the stored post-activation y), bf16 rows
__global__ void __launch_bounds__(256) relu_mask_kernel(const bf16* __restrict__ y, long long y_ld,
                                                        const bf16* __restrict__ dy, long long dy_ld, long long rows,
                                                        int C, bf16* __restrict__ out, long long out_ld) {
  const int nv = C >> 3;
  const long long total = rows * nv;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nv;
    const int c = (int)(i % nv) * 8;
    // y > 0 from the raw 16-bit patterns: sign bit clear and not zero -- identical for bf16 and IEEE fp16 (NaN never occurs)
    const uint4 yb = *reinterpret_cast<const uint4*>(y + r * y_ld + c);
    uint4 g = *reinterpret_cast<const uint4*>(dy + r * dy_ld + c);
    auto mask = [](uint32_t yy) {
      const uint32_t lo = yy & 0xffffu, hi = yy >> 16;
      return ((lo - 1u) < 0x7fffu ? 0xffffu : 0u) | ((hi - 1u) < 0x7fffu ? 0xffff0000u : 0u);
    };
    g.x &= mask(yb.x); g.y &= mask(yb.y); g.z &= mask(yb.z); g.w &= mask(yb.w);
    *reinterpret_cast<uint4*>(out + r * out_ld + c) = g;
  }
}


// ------------------------------------------------------------------------------------------------ NAFBlock scaled convs
// y = x + scale[co] * (W (g * s_b) + bias)  (N:225-237: conv3 with beta and the SCA vector s_b, conv5 with gamma, s = 1).
// Given raw[b][co][ci] = sum_p dy[b,p,co] g[b,p,ci] (tdr_wgrad) and cs[co] = sum_p dy[p,co]:
//   dW[co][ci] += scale[co] * sum_b s_b[ci] raw_b[co][ci];   dscale[co] += sum_{b,ci} W[co][ci] s_b[ci] raw_b[co][ci] + bias[co] cs[co]
//   dbias[co] += scale[co] * cs[co].     grid (Co), block 256.
__global__ void __launch_bounds__(256) naf_scaled_conv_bwd_kernel(const float* __restrict__ raw, int nb, int Co, int C,
                                                                  const float* __restrict__ W,
                                                                  const float* __restrict__ bias,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ cs,
                                                                  const float* __restrict__ s, float* __restrict__ dW,
                                                                  float* __restrict__ dbias, float* __restrict__ dscale) {
  __shared__ float red[8];
  const int co = blockIdx.x;
  const float sc = scale[co];
  float acc = 0.f;
  for (int ci = threadIdx.x; ci < C; ci += blockDim.x) {
    float t = 0.f;
    for (int b = 0; b < nb; ++b) t = fmaf(s ? s[(size_t)b * C + ci] : 1.f, raw[((size_t)b * Co + co) * C + ci], t);
    dW[(size_t)co * C + ci] += sc * t;
    acc = fmaf(W[(size_t)co * C + ci], t, acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    const float b0 = bias ? bias[co] : 0.f;
    dscale[co] += t + b0 * cs[co];
    if (dbias) dbias[co] += sc * cs[co];
  }
}

// ds[b][ci] = sum_co scale[co] W3[co][ci] raw_b[co][ci].  grid (ceil(C/32), B), block (32, 8): lanes over ci (coalesced),
// the 8 warps split the co loop and are reduced through shared memory.
__global__ void __launch_bounds__(256) naf_sca_ds_kernel(const float* __restrict__ raw, int B, int Co, int C,
                                                         const float* __restrict__ W3, const float* __restrict__ scale,
                                                         float* __restrict__ ds) {
  __shared__ float red[8][32];
  const int ci = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y, ty = threadIdx.y;
  float t = 0.f;
  if (ci < C)
    for (int co = ty; co < Co; co += 8)
      t = fmaf(scale[co] * W3[(size_t)co * C + ci], raw[((size_t)b * Co + co) * C + ci], t);
  red[ty][threadIdx.x] = t;
  __syncthreads();
  if (ty == 0 && ci < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    ds[(size_t)b * C + ci] = s;
  }
}

// dgadd[b][j] = sum_i Wsca[i][j] ds_b[i] / P   (same mapping: lanes over j, warps over i)
__global__ void __launch_bounds__(256) naf_sca_dgadd_kernel(const float* __restrict__ ds, const float* __restrict__ w_sca,
                                                            int C, float inv_p, float* __restrict__ dgadd) {
  __shared__ float red[8][32];
  const int j = blockIdx.x * 32 + threadIdx.x, b = blockIdx.y, ty = threadIdx.y;
  float t = 0.f;
  if (j < C)
    for (int i = ty; i < C; i += 8) t = fmaf(w_sca[(size_t)i * C + j], ds[(size_t)b * C + i], t);
  red[ty][threadIdx.x] = t;
  __syncthreads();
  if (ty == 0 && j < C) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    dgadd[(size_t)b * C + j] = s * inv_p;
  }
}

// dWsca[i][j] += sum_b ds_b[i] mean_b[j];  dbsca[i] += sum_b ds_b[i].   grid (C), block 256 over j.
__global__ void __launch_bounds__(256) naf_sca_bwd_kernel(const float* __restrict__ ds, const float* __restrict__ mean,
                                                          int B, int C, float* __restrict__ dwsca,
                                                          float* __restrict__ dbsca) {
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t = fmaf(ds[(size_t)b * C + i], mean[(size_t)b * C + j], t);
    dwsca[(size_t)i * C + j] += t;
  }
  if (threadIdx.x == 0 && dbsca) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += ds[(size_t)b * C + i];
    dbsca[i] += t;
  }
}

}  // namespace

// ================================================================================================ C ABI
extern "C" size_t tdr_wgrad_workspace_bytes(const tdr_wgrad_desc* d) {
  WgradPlan p;
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || d->Ci <= 0 || d->Co <= 0 || d->KH < 1 || d->KW < 1 ||
      d->stride < 1 || d->dil < 1 || wgrad_plan(d, &p))
    return 0;
  return (size_t)p.nb * p.nchunks * p.T * d->Co * d->Ci * sizeof(float);
}

extern "C" int tdr_wgrad(const tdr_wgrad_desc* d, cudaStream_t stream) {
  TDR_CHECK_ARG(d != nullptr, "tdr_wgrad: null descriptor");
  TDR_CHECK_ARG(d->dy && d->x && d->out && d->workspace, "tdr_wgrad: null pointer");
  TDR_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Ci > 0 && d->Co > 0, "tdr_wgrad: bad dims");
  TDR_CHECK_ARG(d->KH >= 1 && d->KW >= 1 && d->stride >= 1 && d->stride <= 2 && d->dil >= 1, "tdr_wgrad: bad filter geometry");
  TDR_CHECK_ARG(d->dy_ld % 8 == 0 && d->x_ld % 8 == 0, "tdr_wgrad: row strides must be multiples of 8 (16 B rows)");
  TDR_CHECK_ARG(((uintptr_t)d->dy & 15) == 0 && ((uintptr_t)d->x & 15) == 0, "tdr_wgrad: 16 B alignment");
  WgradArgs a;
  TDR_CHECK_ARG(wgrad_plan(d, &a.plan) == 0, "tdr_wgrad: unsupported shape (Co=%d Ci=%d)", d->Co, d->Ci);
  const WgradPlan& p = a.plan;
  const size_t need = (size_t)p.nb * p.nchunks * p.T * d->Co * d->Ci * sizeof(float);
  TDR_CHECK_ARG(d->workspace_bytes >= need, "tdr_wgrad: workspace too small (%zu < %zu)", d->workspace_bytes, need);
  a.Co = d->Co; a.Ci = d->Ci; a.KW = d->KW; a.stride = d->stride; a.pad = d->pad; a.dil = d->dil;
  a.per_sample = d->per_sample; a.partials = d->workspace;
  TdrTensorMap map_dy, map_x;
  {
    const uint64_t dims[4] = {(uint64_t)d->Co, (uint64_t)p.OW, (uint64_t)p.OH, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->dy_ld * 2, (uint64_t)d->dy_ld * 2 * p.OW, (uint64_t)d->dy_ld * 2 * p.OW * p.OH};
    const uint32_t box[4] = {64, (uint32_t)p.TW, (uint32_t)p.TH, 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = tdr_make_tensor_map_bf16(&map_dy, d->dy, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d->Ci, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->x_ld * 2, (uint64_t)d->x_ld * 2 * d->W, (uint64_t)d->x_ld * 2 * d->W * d->H};
    uint32_t box[4] = {64, (uint32_t)(p.TW * d->stride), (uint32_t)(p.TH * d->stride), 1};
    if (p.fused) { box[1] = 8; box[2] = 18; }                  // haloed rows, one kx shift per load
    const uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    TDR_CHECK_ARG(box[1] <= 256 && box[2] <= 256, "tdr_wgrad: TMA box too large");
    int rc = tdr_make_tensor_map_bf16(&map_x, d->x, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  const size_t smem = 1024 + (size_t)p.stages * p.stage_bytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((p.fused ? 1 : p.T) * p.m_tiles * p.n_tiles, p.nchunks, p.nb);
  wgrad_kernel<<<grid, 192, smem, stream>>>(map_dy, map_x, a);
  TDR_CHECK_LAUNCH();
  const long long total = (long long)p.nb * p.T * d->Co * d->Ci;
  wgrad_reduce_kernel<<<grid_for(total, 256, 16), 256, 0, stream>>>(
      d->workspace, p.nb, p.nchunks, p.T, d->Co, d->Ci, d->out, d->out_stride_b, d->out_stride_co, d->out_stride_ci,
      d->out_stride_tap, d->co_map, d->ci_map, d->accumulate, d->scale, d->scale_ptr);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" size_t tdr_reduce_workspace_bytes(int C) { return (size_t)kRedBlocks * 10 * (size_t)(C > 0 ? C : 0) * sizeof(float); }   // >= 3 * kRedBlocks * 2 * C (rownorm_bwd)

extern "C" int tdr_colsum(const void* x_bf16, long long ld, long long rows, int C, float* out, long long out_stride,
                          const int* c_map, int accumulate, float* workspace, cudaStream_t stream) {
  TDR_CHECK_ARG(x_bf16 && out && workspace && rows > 0 && C > 0, "tdr_colsum: bad arguments");
  TDR_CHECK_ARG(C % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x_bf16 & 15) == 0, "tdr_colsum: C, ld multiples of 8, 16 B aligned");
  int nblk = (int)((rows + 31) / 32 < kRedBlocks ? (rows + 31) / 32 : kRedBlocks);
  dim3 grid(nblk, tdr_cdiv(C / 8, 256));
  colsum_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const bf16*>(x_bf16), ld, rows, C, workspace);
  TDR_CHECK_LAUNCH();
  reduce_parts_wide_kernel<<<tdr_cdiv(C, 32), dim3(32, 8), 0, stream>>>(workspace, nblk, C, out, out_stride, c_map, accumulate);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_dwconv3x3_wgrad(const void* dy_bf16, long long dy_ld, const void* x_bf16, long long x_ld, int B, int H,
                                   int W, int C, float* dw, float* db, const int* c_map, int accumulate,
                                   float* workspace, cudaStream_t stream) {
  TDR_CHECK_ARG(dy_bf16 && x_bf16 && dw && workspace, "tdr_dwconv3x3_wgrad: null pointer");
  TDR_CHECK_ARG(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "tdr_dwconv3x3_wgrad: bad dims");
  TDR_CHECK_ARG(dy_ld % 8 == 0 && x_ld % 8 == 0 && ((uintptr_t)dy_bf16 & 15) == 0 && ((uintptr_t)x_bf16 & 15) == 0,
                "tdr_dwconv3x3_wgrad: alignment");
  DwgArgs a;
  a.B = B; a.H = H; a.W = W; a.C = C;
  a.tiles_x = tdr_cdiv(W, kDgCols); a.tiles_y = tdr_cdiv(H, kDgRows);
  a.chunks = tdr_cdiv(C, 64);
  a.partials = workspace;
  const int n_spatial = B * a.tiles_x * a.tiles_y;
  int groups = tdr_num_sms() / a.chunks;
  if (groups < 1) groups = 1;
  if (groups > n_spatial) groups = n_spatial;
  if (groups > kRedBlocks) groups = kRedBlocks;
  const int nblk = groups;
  TdrTensorMap map_x, map_dy;
  {
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    const uint32_t es[4] = {1, 1, 1, 1};
    const uint64_t sx[3] = {(uint64_t)x_ld * 2, (uint64_t)x_ld * 2 * W, (uint64_t)x_ld * 2 * W * H};
    const uint32_t bx[4] = {64, (uint32_t)(kDgCols + 2), (uint32_t)(kDgRows + 2), 1};
    int rc = tdr_make_tensor_map_bf16_noswizzle(&map_x, x_bf16, 4, dims, sx, bx, es);
    if (rc) return rc;
    const uint64_t sy[3] = {(uint64_t)dy_ld * 2, (uint64_t)dy_ld * 2 * W, (uint64_t)dy_ld * 2 * W * H};
    const uint32_t by[4] = {64, (uint32_t)kDgCols, (uint32_t)kDgRows, 1};
    rc = tdr_make_tensor_map_bf16_noswizzle(&map_dy, dy_bf16, 4, dims, sy, by, es);
    if (rc) return rc;
  }
  const size_t smem = 2 * kDgStageBytes;
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(dw_wgrad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dw_wgrad_tma_kernel<<<groups * a.chunks, 256, smem, stream>>>(map_x, map_dy, a);
  TDR_CHECK_LAUNCH();
  dw_wgrad_reduce_kernel<<<tdr_cdiv(10 * C, 256), 256, 0, stream>>>(workspace, nblk, C, dw, db, c_map, accumulate);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_rownorm_bwd(const float* x, long long x_ld, const void* dy_bf16, long long dy_ld, long long rows, int C,
                               int mode, const float* weight, float eps, const float* add, long long add_ld, float* dx,
                               long long dx_ld, void* dx_bf16, long long dx_bf16_ld, float* dweight, float* dbias,
                               int accumulate, float* workspace, cudaStream_t stream) {
  TDR_CHECK_ARG(dy_bf16 && dx && rows > 0 && C > 0, "tdr_rownorm_bwd: bad arguments");
  TDR_CHECK_ARG(dx_bf16_ld % 4 == 0, "tdr_rownorm_bwd: dx_bf16_ld must be a multiple of 4");
  TDR_CHECK_ARG(mode >= 0 && mode <= 2, "tdr_rownorm_bwd: bad mode");
  TDR_CHECK_ARG(mode == 0 || (x && weight), "tdr_rownorm_bwd: x and weight required");
  TDR_CHECK_ARG(C % 4 == 0 && x_ld % 4 == 0 && dy_ld % 4 == 0 && add_ld % 4 == 0 && dx_ld % 4 == 0,
                "tdr_rownorm_bwd: C and strides must be multiples of 4");
  TDR_CHECK_ARG(C <= 4096, "tdr_rownorm_bwd: C too large (%d)", C);
  const bool want_w = mode != 0 && dweight != nullptr;
  TDR_CHECK_ARG(!want_w || workspace, "tdr_rownorm_bwd: workspace required for the weight gradient");
  const int nvec = C / 4;
  int G = 1;
  while (G < 32 && (nvec + G - 1) / G > 4) G <<= 1;
  if (const char* e = getenv("TDR_LNB_G")) {                                     // tuning knob (experiments only)
    const int v = atoi(e);
    if ((v == 2 || v == 4 || v == 8 || v == 16 || v == 32) && (nvec + v - 1) / v <= 32) G = v;
  }
  const int nv = (nvec + G - 1) / G;
  const int slots = 8 * (32 / G);
  long long nb = (rows + slots - 1) / slots;
  long long cap = 3LL * tdr_num_sms() < 3 * kRedBlocks ? 3LL * tdr_num_sms() : 3 * kRedBlocks;   // one full wave at 3 CTAs / SM
  if (const char* e = getenv("TDR_LNB_WAVES")) {
    const long long v = atoll(e) * tdr_num_sms();
    if (v >= tdr_num_sms() && v <= 3 * kRedBlocks) cap = v;
  }
  const int blocks = (int)(nb < cap ? nb : cap);
  const size_t smem = want_w ? (size_t)2 * slots * C * sizeof(float) : 0;
  TDR_CHECK_ARG(smem <= 200 * 1024, "tdr_rownorm_bwd: C too large for the weight-gradient reduction");
  const bf16* g = reinterpret_cast<const bf16*>(dy_bf16);
  float* parts = want_w ? workspace : nullptr;
#define TDR_RNB(NV)                                                                                                    \
  do {                                                                                                                 \
    static bool attr_set = false;                                                                                      \
    if (!attr_set) {                                                                                                   \
      TDR_CHECK_CUDA(cudaFuncSetAttribute(rownorm_bwd_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
      attr_set = true;                                                                                                 \
    }                                                                                                                  \
    rownorm_bwd_kernel<NV><<<blocks, 256, smem, stream>>>(x, x_ld, g, dy_ld, rows, C, mode, weight, eps, add, add_ld, dx, \
                                                          dx_ld, reinterpret_cast<bf16*>(dx_bf16), dx_bf16_ld, parts, G); \
  } while (0)
  if (nv <= 1) TDR_RNB(1);
  else if (nv <= 2) TDR_RNB(2);
  else if (nv <= 3) TDR_RNB(3);
  else if (nv <= 4) TDR_RNB(4);
  else if (nv <= 8) TDR_RNB(8);
  else if (nv <= 16) TDR_RNB(16);
  else TDR_RNB(32);
#undef TDR_RNB
  TDR_CHECK_LAUNCH();
  if (want_w) {
    // partials rows are [dweight(C) | dbias(C)] per block
    rownorm_bwd_reduce_kernel<<<tdr_cdiv(2 * C, 32), dim3(32, 8), 0, stream>>>(workspace, blocks, C, dweight, dbias, accumulate);
    TDR_CHECK_LAUNCH();
  }
  return TDR_OK;
}

extern "C" int tdr_gate_bwd(const void* y_bf16, long long y_ld, const void* dg_bf16, long long dg_ld, long long rows,
                            int Ch, int gate, void* dy_bf16, long long dy_ld, const float* dg_add,
                            long long rows_per_sample, cudaStream_t stream) {
  TDR_CHECK_ARG(!dg_add || rows_per_sample > 0, "tdr_gate_bwd: rows_per_sample required with dg_add");
  TDR_CHECK_ARG(y_bf16 && dg_bf16 && dy_bf16 && rows > 0 && Ch > 0, "tdr_gate_bwd: bad arguments");
  TDR_CHECK_ARG(gate == 1 || gate == 2, "tdr_gate_bwd: gate must be 1 (GELU) or 2 (SimpleGate)");
  TDR_CHECK_ARG(Ch % 8 == 0 && y_ld % 8 == 0 && dg_ld % 8 == 0 && dy_ld % 8 == 0, "tdr_gate_bwd: multiples of 8");
  gate_bwd_kernel<<<grid_for(rows * (Ch / 8), 256, 16), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(y_bf16), y_ld, reinterpret_cast<const bf16*>(dg_bf16), dg_ld, rows, Ch, gate,
      reinterpret_cast<bf16*>(dy_bf16), dy_ld, dg_add, rows_per_sample);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_mdta_bwd(const float* shat, const float* attn, int B, long long P, int C, int heads,
                            const float* temperature, const float* w_out, const float* dweff, void* mqk_bf16,
                            long long mqk_ld, float* dw_out, float* dtemperature, int accumulate, float* workspace,
                            cudaStream_t stream) {
  TDR_CHECK_ARG(shat && attn && temperature && w_out && dweff && mqk_bf16 && dw_out && dtemperature && workspace,
                "tdr_mdta_bwd: null pointer");
  (void)P;
  TDR_CHECK_ARG(heads > 0 && C % heads == 0 && C / heads <= 128 && B > 0, "tdr_mdta_bwd: unsupported head width");
  TDR_CHECK_ARG(mqk_ld >= 2 * C && mqk_ld % 8 == 0, "tdr_mdta_bwd: bad mqk_ld");
  const int c = C / heads;
  float* dwout_part = workspace;                              // [B][C][C]
  float* dattn = workspace + (size_t)B * C * C;               // [B][heads][c][c]
  float* dtemp_part = dattn + (size_t)B * heads * c * c;      // [B][heads]
  {
    // dattn[b,h][i][j] = sum_co W_out[co][hc+i] * dWeff[b][co][hc+j]
    SgemmArgs g;
    g.A = w_out; g.B = dweff; g.C = dattn; g.M = c; g.N = c; g.K = C; g.nb2 = heads;
    g.sam = 1; g.sak = C; g.sbk = C; g.sbn = 1; g.scm = c; g.scn = 1;
    g.a_b1 = 0; g.a_b2 = c; g.b_b1 = (long long)C * C; g.b_b2 = c; g.c_b1 = (long long)heads * c * c; g.c_b2 = (long long)c * c;
    sgemm_batched_kernel<<<dim3(tdr_cdiv(c, 64), tdr_cdiv(c, 64), B * heads), 256, 0, stream>>>(g);
    TDR_CHECK_LAUNCH();
    // dW_out[b][co][hc+i] = sum_j dWeff[b][co][hc+j] * attn[b,h][i][j]
    g.A = dweff; g.B = attn; g.C = dwout_part; g.M = C; g.N = c; g.K = c;
    g.sam = C; g.sak = 1; g.sbk = 1; g.sbn = c; g.scm = C; g.scn = 1;
    g.a_b1 = (long long)C * C; g.a_b2 = c; g.b_b1 = (long long)heads * c * c; g.b_b2 = (long long)c * c;
    g.c_b1 = (long long)C * C; g.c_b2 = c;
    sgemm_batched_kernel<<<dim3(tdr_cdiv(c, 64), tdr_cdiv(C, 64), B * heads), 256, 0, stream>>>(g);
    TDR_CHECK_LAUNCH();
  }
  MdtaBwdArgs a;
  a.C = C; a.heads = heads;
  a.shat = shat; a.attn = attn; a.dattn = dattn; a.temperature = temperature;
  a.mqk = reinterpret_cast<bf16*>(mqk_bf16); a.mqk_ld = mqk_ld;
  a.dtemp_part = dtemp_part;
  const size_t smem = ((size_t)3 * c * (c + 1) + 4 * c) * sizeof(float);
  TDR_CHECK_ARG(smem <= 220 * 1024, "tdr_mdta_bwd: head too wide for shared memory");
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(mdta_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    attr_set = true;
  }
  int nz = tdr_cdiv(tdr_num_sms(), heads * B);
  if (nz > c) nz = c;
  if (nz < 1) nz = 1;
  mdta_bwd_kernel<<<dim3(heads, B, nz), 256, smem, stream>>>(a);
  TDR_CHECK_LAUNCH();
  reduce_parts_kernel<<<tdr_cdiv(C * C, 256), 256, 0, stream>>>(dwout_part, B, C * C, dw_out, 1, nullptr, accumulate, 1.f);
  TDR_CHECK_LAUNCH();
  reduce_parts_kernel<<<1, 256, 0, stream>>>(dtemp_part, B, heads, dtemperature, 1, nullptr, accumulate, 1.f);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" size_t tdr_mdta_bwd_workspace_bytes(int B, int C, int heads) {
  const size_t c = heads > 0 ? (size_t)(C / heads) : 0;
  return ((size_t)B * C * C + (size_t)B * heads * c * c + (size_t)B * heads) * sizeof(float);
}

extern "C" int tdr_scale_add_f32(const float* x, long long x_ld, const float* y, long long y_ld, long long rows, int C,
                                 const float* scale_ptr, float scale, float* out, long long out_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(x && out && rows > 0 && C > 0 && C % 4 == 0 && x_ld % 4 == 0 && y_ld % 4 == 0 && out_ld % 4 == 0,
                "tdr_scale_add_f32: bad arguments");
  scale_add_kernel<<<grid_for(rows * (C / 4), 256, 16), 256, 0, stream>>>(x, x_ld, y, y_ld, rows, C, scale_ptr, scale, out, out_ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_dot_f32(const float* x, long long x_ld, const float* y, long long y_ld, long long rows, int C,
                           float* out, int accumulate, float* workspace, cudaStream_t stream) {
  TDR_CHECK_ARG(x && y && out && workspace && rows > 0 && C > 0 && C % 4 == 0 && x_ld % 4 == 0 && y_ld % 4 == 0,
                "tdr_dot_f32: bad arguments");
  const int nblk = grid_for(rows * (C / 4), 256, 2);
  dot_kernel<<<nblk, 256, 0, stream>>>(x, x_ld, y, y_ld, rows, C, workspace);
  TDR_CHECK_LAUNCH();
  reduce_parts_kernel<<<1, 32, 0, stream>>>(workspace, nblk, 1, out, 1, nullptr, accumulate, 1.f);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_pixel_shuffle_nhwc(const void* in_bf16, long long in_ld, int B, int H, int W, int C, int mode,
                                      void* out_bf16, long long out_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(in_bf16 && out_bf16 && B > 0 && H > 0 && W > 0 && C > 0, "tdr_pixel_shuffle_nhwc: bad arguments");
  TDR_CHECK_ARG(mode == 1 ? (H % 2 == 0 && W % 2 == 0) : (mode == 2 && C % 4 == 0), "tdr_pixel_shuffle_nhwc: bad mode/shape");
  const bool aligned = in_ld % 8 == 0 && out_ld % 8 == 0 && ((uintptr_t)in_bf16 & 15) == 0 && ((uintptr_t)out_bf16 & 15) == 0 &&
                       (long long)B * H * W * C < (1ll << 31);
  if (mode == 1 && aligned && C % 8 == 0)
    pixel_unshuffle_vec_kernel<<<grid_for((long long)B * (H / 2) * (W / 2) * (C / 8), 256, 16), 256, 0, stream>>>(
        reinterpret_cast<const uint16_t*>(in_bf16), in_ld, B, H, W, C, reinterpret_cast<uint16_t*>(out_bf16), out_ld);
  else if (mode == 2 && aligned && C % 32 == 0)
    pixel_shuffle_vec_kernel<<<grid_for((long long)B * H * W * (C / 32), 256, 16), 256, 0, stream>>>(
        reinterpret_cast<const uint16_t*>(in_bf16), in_ld, B, H, W, C, reinterpret_cast<uint16_t*>(out_bf16), out_ld);
  else
    pixel_shuffle_kernel<<<grid_for((long long)B * H * W * C, 256, 16), 256, 0, stream>>>(
        reinterpret_cast<const bf16*>(in_bf16), in_ld, B, H, W, C, mode, reinterpret_cast<bf16*>(out_bf16), out_ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_relu_mask(const void* y_bf16, long long y_ld, const void* dy_bf16, long long dy_ld, long long rows,
                             int C, void* out_bf16, long long out_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(y_bf16 && dy_bf16 && out_bf16 && rows > 0 && C > 0 && C % 8 == 0 && y_ld % 8 == 0 && dy_ld % 8 == 0 &&
                    out_ld % 8 == 0,
                "tdr_relu_mask: bad arguments");
  relu_mask_kernel<<<grid_for(rows * (C / 8), 256, 16), 256, 0, stream>>>(
      reinterpret_cast<const bf16*>(y_bf16), y_ld, reinterpret_cast<const bf16*>(dy_bf16), dy_ld, rows, C,
      reinterpret_cast<bf16*>(out_bf16), out_ld);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_naf_scaled_conv_bwd(const float* raw, int nb, int Co, int C, const float* W, const float* bias,
                                       const float* scale, const float* colsum_dy, const float* s, float* dW,
                                       float* dbias, float* dscale, cudaStream_t stream) {
  TDR_CHECK_ARG(raw && W && scale && colsum_dy && dW && dscale && nb > 0 && Co > 0 && C > 0,
                "tdr_naf_scaled_conv_bwd: bad arguments");
  naf_scaled_conv_bwd_kernel<<<Co, 256, 0, stream>>>(raw, nb, Co, C, W, bias, scale, colsum_dy, s, dW, dbias, dscale);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_naf_sca_bwd(const float* raw, int B, int Co, int C, const float* w3, const float* scale,
                               const float* mean, const float* w_sca, long long P, float* dw_sca, float* db_sca,
                               float* dg_add, float* workspace /* B*C floats */, cudaStream_t stream) {
  TDR_CHECK_ARG(raw && w3 && scale && mean && w_sca && dw_sca && dg_add && workspace && B > 0 && Co > 0 && C > 0 && P > 0,
                "tdr_naf_sca_bwd: bad arguments");
  naf_sca_ds_kernel<<<dim3(tdr_cdiv(C, 32), B), dim3(32, 8), 0, stream>>>(raw, B, Co, C, w3, scale, workspace);
  TDR_CHECK_LAUNCH();
  naf_sca_dgadd_kernel<<<dim3(tdr_cdiv(C, 32), B), dim3(32, 8), 0, stream>>>(workspace, w_sca, C, 1.f / (float)P, dg_add);
  TDR_CHECK_LAUNCH();
  naf_sca_bwd_kernel<<<C, 256, 0, stream>>>(workspace, mean, B, C, dw_sca, db_sca);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
