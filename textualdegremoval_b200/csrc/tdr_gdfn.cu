// tdr_gdfn_tail: the second half of Restormer's gated-dconv feed-forward (GDFN) in ONE kernel,
// /root/reference/models/archs/network_restormer_guided_arch.py:236-240 (FeedForward.forward) + the residual add :329 /
// the Res-fusion epilogue :345-353:
//
//     x1, x2 = dwconv3x3(hidden).chunk(2)          hidden = project_in(norm2(x)),  2*hp channels
//     out    = project_out(gelu(x1) * x2) + residual(s)
//
// Unfused this is a depthwise+gate kernel that WRITES the gated tensor g (hp channels, 16-bit) and a 1x1 conv kernel
// that READS it back: 4*hp B/pixel of HBM traffic (21 % of a transformer block's bytes) and ~50 launches per forward for
// nothing -- the gate is pure CUDA-core work and the tensor pipe idles meanwhile.  Here the gated values never leave the
// SM: 8 compute warps run the depthwise stencil + exact GELU gate on a TMA-staged haloed tile and write the result
// straight into a SWIZZLE_128B K-major shared-memory tile, which one thread feeds to tcgen05.mma against the matching
// 64-channel slice of W_out (TMA-loaded next to the input slice); the [128 px x C] accumulator lives in TMEM across all
// hp/64 slices, and the same 8 warps then add the residual(s) and store the fp32 rows.
//
// Mapping (one persistent CTA per SM, 22 warps; a work item = 8 rows x 16 cols = 128 pixels):
//   warp 0      TMA producer: per 64-channel slice kc two haloed [10 x 18 px x 64 ch] tiles of `hidden` (the x1 and x2
//               halves; out-of-image rows / cols / channels are zero-filled = the conv padding) + the [C x 64] slice of
//               W_out (SWIZZLE_128B), 2..4-stage mbarrier ring (3 stages at C = 96: two loads in flight per SM);
//   warps 6..21 warp = one pixel column of the tile, lane = one channel pair of the slice: sliding-window stencil over the
//               10 staged rows in packed fp32 (FFMA2), gate, 4 B store of the gated pair into the A tile (a warp writes
//               one 128 B swizzle row); then fence.proxy.async + arrive.  16 warps (4 per scheduler), ~90 registers:
//               the stencil is latency-bound, with 8 warps of 4-channel threads it ran 2.4x slower;
//   warp 1      one thread issues 4 x tcgen05.mma (M = 128, N = C, K = 16) per slice; tcgen05.commit frees the A tile
//               and the ring stage, the last slice's commit publishes the accumulator;
//   warps 2..5  epilogue, one warp per TMEM lane quadrant, on the OTHER of two TMEM accumulators while the stencil warps
//               already work on the next tile: residual rows are fetched first (before the accumulator is even ready),
//               then tcgen05.ld (thread = pixel row) -> bias / alpha -> staging tile -> coalesced fp32 stores (8 lanes per
//               pixel row).  (With the stencil warps doing the epilogue themselves the kernel spent 40 % of its time in
//               it: ncu showed the serialised residual loads and the tile-end barrier on top.)
#include <stdlib.h>
#include <string.h>

#include "tdr_stencil.cuh"

namespace {

constexpr int kTH = 8, kTW = 16;                               // pixel tile (M = 128)
constexpr int kHaloBytes = (kTH + 2) * (kTW + 2) * 64 * 2;     // 23040: one haloed 64-channel tile
constexpr int kATile = 128 * 64 * 2;                           // 16 KiB
constexpr int kMaxStages = 4;
constexpr int kComputeWarps = 16;                              // 4 per scheduler: the stencil is latency-bound
constexpr int kEpiWarps = 4;                                   // one per TMEM lane quadrant, overlapped with the next tile
constexpr int kThreads = 32 * (2 + kEpiWarps + kComputeWarps);
constexpr int kEpiStage = 32 * 128;                            // per warp: 32 rows x 32 fp32 columns

struct GdfnArgs {
  int B, H, W, hp, C;              // hp = channels per half of `hidden` (padded hidden width), C = output channels
  int tiles_x, tiles_y, n_tiles, nchunks;
  int stage_bytes;                 // 2 * kHaloBytes + C * 128
  int stages;                      // ring depth (2..4): as many as fit -- ONE stage in flight per SM (2-deep ring) caps the
                                   // kernel at stage_bytes / load latency, far below the stencil's own speed
  const float* wt;                 // depthwise taps fp32 [9][2*hp]
  const float* dw_bias;            // [2*hp] or null
  const float* bias;               // project_out bias [C] or null
  float alpha, res1_scale;
  const float* scale_ptr;          // optional device scalar g multiplying alpha and res1_scale
  const float* res1; long long res1_ld;
  const float* res2; long long res2_ld;
  float* out;        long long out_ld;
  uint32_t tmem_cols;
  int half_in;                     // hidden / W_out are IEEE fp16 (else bf16)
};

template <bool HALF>
__global__ void __launch_bounds__(kThreads, 1)
gdfn_tail_kernel(const __grid_constant__ TdrTensorMap map_a, const __grid_constant__ TdrTensorMap map_b,
                 const __grid_constant__ TdrTensorMap map_w, const GdfnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kStages = a.stages;
  uint8_t* ring = smem;                                        // stages x stage_bytes (multiples of 1024)
  uint8_t* atile = ring + kStages * a.stage_bytes;             // 2 x 16 KiB
  uint8_t* epi = atile + 2 * kATile;                           // epilogue staging: kEpiWarps x 4 KiB
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + kEpiWarps * kEpiStage);
  uint64_t* full = bars;             // [kMaxStages]  TMA landed
  uint64_t* empty = bars + 4;        // [kMaxStages]  stage consumed (MMA commit)
  uint64_t* a_full = bars + 8;       // [2]        gated A tile written (16 warp arrivals)
  uint64_t* a_empty = bars + 10;     // [2]        A tile consumed (MMA commit)
  uint64_t* tfull = bars + 12;       // [2] accumulator complete
  uint64_t* tempty = bars + 14;      // [2] accumulator drained (4 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_w);
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], kComputeWarps); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull[s], 1); mbar_init(&tempty[s], kEpiWarps); }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int per_img = a.tiles_y * a.tiles_x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
        const int b = t / per_img, r = t % per_img;
        const int y0 = (r / a.tiles_x) * kTH, x0 = (r % a.tiles_x) * kTW;
        for (int kc = 0; kc < a.nchunks; ++kc) {
          mbar_wait_sleep(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], (uint32_t)a.stage_bytes);
          uint8_t* dst = ring + stage * a.stage_bytes;
          tma_load_4d(dst, &map_a, &full[stage], kc * 64, x0 - 1, y0 - 1, b);
          tma_load_4d(dst + kHaloBytes, &map_b, &full[stage], kc * 64, x0 - 1, y0 - 1, b);
          tma_load_3d(dst + 2 * kHaloBytes, &map_w, &full[stage], kc * 64, 0, 0);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(128, a.C, 0, 0, a.half_in, a.half_in);
    int stage = 0, cnt = 0, it = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int ab = it & 1;                           // accumulator buffer: the epilogue of tile it overlaps tile it + 1
      const uint32_t d_tmem = tmem_base + ab * a.C;
      mbar_wait_sleep(&tempty[ab], ((it >> 1) & 1) ^ 1);     // the epilogue two tiles back has drained this accumulator
      tc_fence_after();
      for (int kc = 0; kc < a.nchunks; ++kc, ++cnt) {
        const int buf = cnt & 1;
        mbar_wait_sleep(&a_full[buf], (cnt >> 1) & 1, 100);   // gated tile written (the compute warps waited on full[stage])
        mbar_wait(&full[stage], phase);                // W slice visible to this thread too
        tc_fence_after();
        if (elect_one()) {
          const uint64_t da = umma_desc_sw128(smem_u32(atile + buf * kATile), 0, 1024);
          const uint64_t db = umma_desc_sw128(smem_u32(ring + stage * a.stage_bytes + 2 * kHaloBytes), 0, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0);
          umma_commit(&a_empty[buf]);
          umma_commit(&empty[stage]);
          if (kc == a.nchunks - 1) umma_commit(&tfull[ab]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 2 + kEpiWarps) {
    // ===================== stencil + gate =====================
    const int cw = warp - 2 - kEpiWarps;           // 0..15
    const int cp = lane;                           // channel PAIR within the 64-channel slice (32 pairs)
    const int xl = cw;                             // column within the tile: a warp = one pixel column, all 64 channels
    int stage = 0, cnt = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x) {
      for (int kc = 0; kc < a.nchunks; ++kc, ++cnt) {
        const int buf = cnt & 1;
        // taps of this slice (L1-resident: 9 x 2hp floats), one channel pair of each half
        const int c0 = kc * 64 + cp * 2;
        const bool c_ok = c0 < a.hp;
        f2 w[2][9], bv[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int tp = 0; tp < 9; ++tp) {
            float2 v = make_float2(0.f, 0.f);
            if (c_ok) v = __ldg(reinterpret_cast<const float2*>(a.wt + (size_t)tp * 2 * a.hp + h * a.hp + c0));
            w[h][tp] = pk2(v.x, v.y);
          }
          float2 bb = make_float2(0.f, 0.f);
          if (a.dw_bias && c_ok) bb = __ldg(reinterpret_cast<const float2*>(a.dw_bias + h * a.hp + c0));
          bv[h] = pk2(bb.x, bb.y);
        }
        mbar_wait(&full[stage], phase);
        mbar_wait(&a_empty[buf], ((cnt >> 1) & 1) ^ 1);
        const uint32_t tile_s = smem_u32(ring + stage * a.stage_bytes) + xl * 128 + cp * 4;   // explicit ld/st.shared: a
        const uint32_t a_dst = smem_u32(atile + buf * kATile);                                // generic LD costs latency
        f2 acc[3][2];
#pragma unroll
        for (int k = 0; k < 3; ++k) { acc[k][0] = bv[0]; acc[k][1] = bv[1]; }
#pragma unroll
        for (int i = 0; i < kTH + 2; ++i) {                  // staged row i = input row y0 - 1 + i
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              const uint32_t rv = lds32(tile_s + h * kHaloBytes + (i * (kTW + 2) + kx) * 128);
              const f2 v = x2_to_f2<HALF>(rv);
#pragma unroll
              for (int k = 0; k < 3; ++k)                    // staged row i feeds output rows i - 2 + k inside the tile
                if (i + k >= 2 && i + k < kTH + 2)
                  acc[(i + k) % 3][h] = fma2(w[h][(2 - k) * 3 + kx], v, acc[(i + k) % 3][h]);
            }
          if (i >= 2) {                                      // output row i - 2 (slot i % 3) is complete: gate + A tile
            const int m = (i - 2) * kTW + xl;                // row of the 128-pixel UMMA tile
            float p0, p1;
            upk2(mul2(gelu2(acc[i % 3][0]), acc[i % 3][1]), p0, p1);
            // K-major SWIZZLE_128B: row m = 128 B (64 channels), 16 B units XOR-ed with (m & 7); the warp writes one row
            sts32(a_dst + m * 128 + ((((cp >> 2) ^ (m & 7))) << 4) + (cp & 3) * 4, pack2t<HALF>(p0, p1));
            acc[i % 3][0] = bv[0];
            acc[i % 3][1] = bv[1];
          }
        }
        fence_proxy_async();                                 // generic-proxy smem writes -> visible to tcgen05.mma
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[buf]);
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  }
  if (warp >= 2 && warp < 2 + kEpiWarps) {
    // ===================== epilogue (overlaps the next tile's stencil): out = alpha * (acc + bias) + r1s * res1 + res2
    const int quad = warp & 3;                     // TMEM lane quadrant this warp may read (warps 2..5 -> 2, 3, 0, 1)
    const float g = a.scale_ptr ? *a.scale_ptr : 1.f;
    const float alpha = a.alpha * g, r1s = a.res1_scale * g;
    uint8_t* stg = epi + (warp - 2) * kEpiStage;
    const int nsb = (a.C + 31) >> 5;               // 32-column sub-blocks
    int it = 0;
    for (int t = blockIdx.x; t < a.n_tiles; t += gridDim.x, ++it) {
      const int b = t / per_img, r = t % per_img;
      const int y0 = (r / a.tiles_x) * kTH, x0 = (r % a.tiles_x) * kTW;
      const int ab = it & 1;
      const int m = quad * 32 + lane;
      const int oy = y0 + m / kTW, ox = x0 + m % kTW;
      const bool valid = oy < a.H && ox < a.W;
      const long long pix = ((long long)b * a.H + oy) * a.W + ox;
      // phase 2 handles 4 pixel rows per step, 8 lanes (128 B) per row: this lane's rows rl = i8 * 4 + (lane >> 3)
      long long prow[8];
      bool vrow[8];
#pragma unroll
      for (int i8 = 0; i8 < 8; ++i8) {
        const int rl = i8 * 4 + (lane >> 3);
        vrow[i8] = __shfl_sync(0xffffffffu, (int)valid, rl) != 0;
        prow[i8] = __shfl_sync(0xffffffffu, pix, rl);
      }
      const int ch = lane & 7;
      const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + ab * a.C;
      mbar_wait_sleep(&tfull[ab], (it >> 1) & 1, 400);
      tc_fence_after();
      for (int sb = 0; sb < nsb; ++sb) {
        const int cs = sb * 32;
        const int col = cs + ch * 4;
        const bool c_ok = col < a.C;
        {
          uint32_t raw[2][16];
          tmem_ld16(t_base + cs, raw[0]);
          tmem_ld16(t_base + cs + 16, raw[1]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int hh = 0; hh < 4; ++hh) {
              float v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                float x = __uint_as_float(raw[j][hh * 4 + i]);
                const int cc = cs + j * 16 + hh * 4 + i;
                if (a.bias && cc < a.C) x += a.bias[cc];
                v[i] = x * alpha;
              }
              *reinterpret_cast<float4*>(stg + lane * 128 + ((((j * 4 + hh) & 7) ^ (lane & 7)) << 4)) =
                  make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        __syncwarp();
        // 4 pixel rows per step, 8 lanes (128 B) per row.  The residual rows of a group of 4 steps are read BEFORE any of its
        // stores (`out` may alias res2: the compiler would otherwise serialise every load behind the previous store)
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
          float4 q2[4], q1[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i8 = grp * 4 + k;
            q2[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            q1[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vrow[i8] && c_ok) {
              if (a.res2) q2[k] = *reinterpret_cast<const float4*>(a.res2 + prow[i8] * a.res2_ld + col);
              if (a.res1) q1[k] = *reinterpret_cast<const float4*>(a.res1 + prow[i8] * a.res1_ld + col);
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i8 = grp * 4 + k;
            const int rl = i8 * 4 + (lane >> 3);
            if (vrow[i8] && c_ok) {
              float4 o = *reinterpret_cast<const float4*>(stg + rl * 128 + ((ch ^ (rl & 7)) << 4));
              o.x += r1s * q1[k].x + q2[k].x; o.y += r1s * q1[k].y + q2[k].y;
              o.z += r1s * q1[k].z + q2[k].z; o.w += r1s * q1[k].w + q2[k].w;
              *reinterpret_cast<float4*>(a.out + prow[i8] * a.out_ld + col) = o;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[ab]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

}  // namespace

static bool gdfn_ok(const tdr_gdfn_tail_desc* d) {
  return d && d->hidden && d->dw_weight && d->w_out && d->out && d->B > 0 && d->H > 0 && d->W > 0 && d->hp > 0 &&
         d->hp % 8 == 0 && d->C >= 16 && d->C <= 192 && d->C % 16 == 0 && d->hidden_ld % 8 == 0 && d->hidden_ld >= 2 * d->hp &&
         d->w_ld % 8 == 0 && d->w_ld >= d->hp && ((uintptr_t)d->hidden & 15) == 0 && ((uintptr_t)d->w_out & 15) == 0 &&
         ((uintptr_t)d->dw_weight & 15) == 0 && (!d->dw_bias || ((uintptr_t)d->dw_bias & 15) == 0) &&
         ((uintptr_t)d->out & 15) == 0 && d->out_ld % 4 == 0 && (!d->res1 || (((uintptr_t)d->res1 & 15) == 0 && d->res1_ld % 4 == 0)) &&
         (!d->res2 || (((uintptr_t)d->res2 & 15) == 0 && d->res2_ld % 4 == 0)) && (d->hp * 2) % 8 == 0;
}

extern "C" int tdr_gdfn_tail_supported(const tdr_gdfn_tail_desc* d) { return gdfn_ok(d) ? 1 : 0; }

extern "C" int tdr_gdfn_tail(const tdr_gdfn_tail_desc* d, cudaStream_t stream) {
  TDR_CHECK_ARG(gdfn_ok(d), "tdr_gdfn_tail: unsupported arguments (need 16 <= C <= 192, C %% 16 == 0, hp %% 8 == 0, "
                "16 B-aligned rows; see tdr_gdfn_tail_supported)");
  GdfnArgs a;
  a.B = d->B; a.H = d->H; a.W = d->W; a.hp = d->hp; a.C = d->C;
  a.tiles_x = tdr_cdiv(d->W, kTW); a.tiles_y = tdr_cdiv(d->H, kTH);
  a.n_tiles = d->B * a.tiles_x * a.tiles_y;
  a.nchunks = tdr_cdiv(d->hp, 64);
  a.stage_bytes = 2 * kHaloBytes + d->C * 128;
  a.stage_bytes = (a.stage_bytes + 1023) / 1024 * 1024;
  a.wt = d->dw_weight; a.dw_bias = d->dw_bias; a.bias = d->bias;
  a.alpha = d->alpha; a.res1_scale = d->res1_scale; a.scale_ptr = d->scale_ptr;
  a.res1 = d->res1; a.res1_ld = d->res1_ld; a.res2 = d->res2; a.res2_ld = d->res2_ld;
  a.out = d->out; a.out_ld = d->out_ld;
  a.half_in = d->fp16 ? 1 : 0;
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * d->C)) cols <<= 1;            // two accumulators: the epilogue of a tile overlaps the next tile
  a.tmem_cols = cols;
  TdrTensorMap map_a, map_b, map_w;
  {
    // the two halves of `hidden` as separate tensors of hp channels each: channels beyond hp are zero-filled
    const uint64_t dims[4] = {(uint64_t)d->hp, (uint64_t)d->W, (uint64_t)d->H, (uint64_t)d->B};
    const uint64_t strides[3] = {(uint64_t)d->hidden_ld * 2, (uint64_t)d->hidden_ld * 2 * d->W,
                                 (uint64_t)d->hidden_ld * 2 * d->W * d->H};
    const uint32_t box[4] = {64, (uint32_t)(kTW + 2), (uint32_t)(kTH + 2), 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    int rc = tdr_make_tensor_map_bf16_noswizzle(&map_a, d->hidden, 4, dims, strides, box, es);
    if (rc) return rc;
    rc = tdr_make_tensor_map_bf16_noswizzle(&map_b, reinterpret_cast<const uint16_t*>(d->hidden) + d->hp, 4, dims, strides,
                                            box, es);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)d->hp, (uint64_t)d->C, 1};
    const uint64_t strides[2] = {(uint64_t)d->w_ld * 2, (uint64_t)d->w_ld * 2 * d->C};
    const uint32_t box[3] = {64, (uint32_t)d->C, 1};
    const uint32_t es[3] = {1, 1, 1};
    int rc = tdr_make_tensor_map_bf16(&map_w, d->w_out, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  a.stages = (int)((227 * 1024 - 1024 - 2 * kATile - kEpiWarps * kEpiStage - 256) / a.stage_bytes);
  if (a.stages > kMaxStages) a.stages = kMaxStages;
  if (const char* e = getenv("TDR_GDFN_STAGES")) {                               // tuning knob (experiments only)
    const int v = atoi(e);
    if (v >= 2 && v <= a.stages) a.stages = v;
  }
  TDR_CHECK_ARG(a.stages >= 2, "tdr_gdfn_tail: shared-memory plan does not fit");
  const size_t smem = 1024 + (size_t)a.stages * a.stage_bytes + 2 * kATile + (size_t)kEpiWarps * kEpiStage + 256;
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(gdfn_tail_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    TDR_CHECK_CUDA(cudaFuncSetAttribute(gdfn_tail_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int grid = a.n_tiles < tdr_num_sms() ? a.n_tiles : tdr_num_sms();
  if (a.half_in) gdfn_tail_kernel<true><<<grid, kThreads, smem, stream>>>(map_a, map_b, map_w, a);
  else gdfn_tail_kernel<false><<<grid, kThreads, smem, stream>>>(map_a, map_b, map_w, a);
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
