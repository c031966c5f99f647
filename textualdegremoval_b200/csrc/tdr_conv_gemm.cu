// tdr_conv_gemm: implicit-GEMM convolution on NHWC bf16 activations with tcgen05 (UMMA) + TMA.
//
//   out[b, oy, ox, co] = epilogue( sum_{ky,kx,ci} in[b, oy*s - pad + ky*dil, ox*s - pad + kx*dil, ci] * W[t][co][ci] )
//
// This one kernel carries every dense contraction of the restoration nets (reference ops, all NCHW
// nn.Conv2d in /root/reference/models/archs/network_restormer_guided_arch.py):
//   1x1 convs  qkv :252, project_out :254, project_in :229, project_out :234, reduce_chan :613,:619
//   3x3 convs  Encoder/ResidualBlock :34-49,:100-134 (stride 1 and 2, bias, ReLU, residual),
//              Downsample/Upsample :372-391 (PixelUnshuffle/PixelShuffle folded into the store map)
//   attn @ v followed by project_out :272-276 as ONE 1x1 conv with per-sample weights W_out*blockdiag(attn)
//   MASA correlations (search :674-696, search_org :654-672) as dilated 3x3 convs with per-sample filters.
//
// Mapping to the hardware (one persistent CTA per SM, 10 warps):
//   warp 0      TMA producer: per (tap, 64-channel chunk) one 4-D box [64ch x TW x TH x 1] of the input
//               (out-of-bounds rows/cols/channels are zero-filled by TMA = conv padding for free) and one
//               3-D box [64ch x BN x 1] of the weights, both SWIZZLE_128B, into a 4-stage mbarrier ring.
//   warp 1      MMA issuer: one elected thread issues 4 x tcgen05.mma (M=128 pixels, N=BN, K=16) per stage,
//               fp32 accumulators in TMEM, double-buffered (2 x BN columns) so the epilogue of tile i
//               overlaps the main loop of tile i+1.  tcgen05.commit releases smem stages / publishes TMEM.
//   warps 2..9  epilogue: tcgen05.ld (thread = pixel row) -> bias / ReLU / row-scale / alpha / two residuals
//               -> 128-bit stores, fp32 and/or bf16, plain or pixel-(un)shuffled addressing.
#include <stdlib.h>
#include <string.h>

#include "tdr_common.cuh"
#include "tdr_stencil.cuh"

namespace {

constexpr int kMaxStages = 8;
constexpr int kTileM = 128;
constexpr int kChunkK = 64;                       // bf16 elements = 128 B = one swizzle row
constexpr int kABytes = kTileM * kChunkK * 2;     // 16 KiB
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kEpiStageBytes = 32 * 128;          // per epilogue warp: 32 rows x 128 B, 16 B chunks XOR-swizzled

struct ConvGemmArgs {
  int B, OH, OW;
  int Ci, Co;
  int KH, KW, stride, pad, dil;
  int TH, TW, tiles_y, tiles_x;
  int BN, n_tiles, kchunks;
  int w_batched;                // weights tensor map's 3rd coordinate = b*taps + tap
  const int* origin;            // optional [B][3] = (image, y0, x0)
  int total_tiles;              // m_tiles * n_tiles
  int m_tiles;                  // pixel tiles
  int by_pixel;                 // item order, see conv_item(): unequal n tiles (Co = 288 -> 192 + 96) must not pile up on
                                // half of the CTAs (148 is even)
  uint32_t tmem_cols;
  int nacc;                     // TMEM accumulator buffers in flight (2..4): nacc * BN <= 512 columns
  // epilogue
  const float* bias;            // [Co] or null
  const float* rowscale;        // [B*OH*OW] or null
  float alpha, res1_scale;
  const float* scale_ptr;       // optional device scalar multiplying alpha and res1_scale
  int act;                      // 0 none, 1 ReLU, 2 exact GELU
  const float* res1;  long long res1_ld;
  const void* res2;   long long res2_ld;  int res2_bf16;
  float* out_f32;     long long out_f32_ld;
  bf16* out_bf16;     long long out_bf16_ld;
  int store_mode;               // 0 plain, 1 pixel-unshuffle(2), 2 pixel-shuffle(2)
  int epi_mode;                 // 0 generic staged stores; 1 TMA epilogue (single-dtype output, residuals via TMA)
  int out_is_f32;               // epi_mode 1: output (and res2) element type
  int halo;                     // 1: stride-1 KxK conv whose weights fit in shared memory: they are loaded ONCE per CTA,
                                //    and ONE haloed A tile per channel chunk is streamed; taps are descriptor offsets
                                //    into it (A is fetched once instead of KH*KW times, B never again)
  int resident_b;               // 1: 1x1 conv whose whole weight matrix fits in shared memory: every (n tile, chunk) slice
                                //    is loaded ONCE per CTA, the ring holds A chunks only (2-3 pixel tiles of look-ahead
                                //    instead of ~1.5 items), and the A chunks of a pixel tile are loaded once and multiplied
                                //    with all n tiles (items are pixel-tile major: by_pixel = 1)
  int halo_bytes;               // ring slot of one haloed A tile: (TH + (KH-1)*dil) rows x halo_pitch px x 128 B, rounded to 1 KB
  int halo_pitch;               // pixels per staged image row: 8 + (KW-1)*dil (exactly what the taps touch)
  int halo_tx;                  // bytes one haloed box delivers (unrounded)
  int stages;                   // smem pipeline depth; in halo mode: depth of the haloed-A ring
  int epi_bufs;                 // staging buffers per epilogue warp (1 or 2; 4 with the fused LayerNorm)
  int alt_items;                // TMA epilogue (no LN): the two warp groups take alternate items
  int pair;                     // streamed-weight multi-tap convs (C >= 96 dense 3x3): TWO pixel tiles share every weight
                                // tile of the ring (M = 256 per B fetch, two TMEM accumulators): the per-tile weight
                                // re-fetch was ~2/3 of the L2 -> SM traffic, which bounds these convs
  int pair_tiles;               // ceil(m_tiles / 2) * n_tiles
  // division by n_tiles / tiles per image / tiles_x as multiply-high + shift (every role recomputes its item's
  // coordinates per item; the epilogue warps spent 13 % of their samples in the integer-division sequences)
  uint32_t fd_nt_m, fd_nt_s, fd_tpi_m, fd_tpi_s, fd_tx_m, fd_tx_s;
  // fused LayerNorm of the output rows (variant bit 3): ln_out = LN(out) in bf16 through map_o2
  int in_fp16;                  // A and W operands are IEEE fp16 (MASA feature encoder)
  int out_fp16;                 // the 16-bit output is IEEE fp16
  int ln_mode;                  // 1 WithBias, 2 BiasFree (as tdr_rownorm)
  float ln_eps;
  const float* ln_w;
  const float* ln_b;
};

// it-th work item of this CTA -> (pixel tile, n tile).  by_pixel: CTA c owns pixel tiles c, c + grid, ... and walks ALL n tiles
// of each in turn (balanced even when the n tiles are unequal, and the A tile is re-used from L2); otherwise (fewer pixel
// tiles than SMs: ViT / mapper linears) items are dealt round-robin so that every SM gets work.
// n / d for 0 <= n < 2^31 with (m, s) from tdr_fast_div_setup(d)
__device__ __forceinline__ int fdiv(int n, uint32_t m, uint32_t s) {
  return m ? (int)(__umulhi((uint32_t)n, m) >> s) : n;
}
__device__ __forceinline__ bool conv_item(const ConvGemmArgs& a, int it, int& mt, int& nt) {
  if (a.pair) {                                   // items 2q, 2q + 1 = the two pixel tiles of pair q (same n tile); an odd
    const int tile = blockIdx.x + (it >> 1) * gridDim.x;      // tail yields a ghost tile past the last image: its TMA loads
    if (tile >= a.pair_tiles) return false;                   // are zero fill and its TMA stores are clipped
    const int mp = fdiv(tile, a.fd_nt_m, a.fd_nt_s);
    nt = tile - mp * a.n_tiles;
    mt = 2 * mp + (it & 1);
    return true;
  }
  if (a.by_pixel) {
    const int q = fdiv(it, a.fd_nt_m, a.fd_nt_s);
    mt = blockIdx.x + q * gridDim.x;
    nt = it - q * a.n_tiles;
    return mt < a.m_tiles;
  }
  const int tile = blockIdx.x + it * gridDim.x;
  mt = fdiv(tile, a.fd_nt_m, a.fd_nt_s);
  nt = tile - mt * a.n_tiles;
  return tile < a.total_tiles;
}
// pixel tile -> (image, tile row, tile column)
__device__ __forceinline__ void conv_tile_coords(const ConvGemmArgs& a, int mt, int tiles_per_img, int& b, int& ty, int& tx) {
  b = fdiv(mt, a.fd_tpi_m, a.fd_tpi_s);
  const int r = mt - b * tiles_per_img;
  ty = fdiv(r, a.fd_tx_m, a.fd_tx_s);
  tx = r - ty * a.tiles_x;
}

// exact GELU on 16 epilogue values: the packed logistic-polynomial form of tdr_stencil.cuh (|error| <= 2 fp32 ulp) -- erff
// per element made the GELU epilogue of the ViT fc1 GEMMs cost as much as the GEMM itself (170 vs 92 us at CLIP-H fc1)
__device__ __forceinline__ void gelu16(float* v) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    float a, b;
    upk2(gelu2(pk2(v[i], v[i + 1])), a, b);
    v[i] = a;
    v[i + 1] = b;
  }
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return fmaxf(x, 0.f);
  if (act == 2) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  return x;
}

// V < 0: generic epilogue (any combination of outputs / residuals / pixel (un)shuffle).
// V >= 0: TMA epilogue specialised at compile time: bit 0 = fp32 output, bit 1 = res2 tile, bit 2 = res1 tile,
//         bit 3 = fused LayerNorm of the output rows (only with bits 0 and 1, Co <= 96);
//         bit 4 = the 16-bit output is IEEE fp16 (only V == 16: no residual tiles).
template <int V>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ TdrTensorMap map_a, const __grid_constant__ TdrTensorMap map_w,
                 const __grid_constant__ TdrTensorMap map_o, const __grid_constant__ TdrTensorMap map_r2,
                 const __grid_constant__ TdrTensorMap map_r1, const __grid_constant__ TdrTensorMap map_o2,
                 const ConvGemmArgs a) {
  constexpr bool kTma = V >= 0;
  constexpr bool kF32 = kTma && (V & 1), kR2 = kTma && (V & 2), kR1 = kTma && (V & 4), kLN = kTma && (V & 8);
  constexpr bool kH16 = kTma && (V & 16);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [A stages][B stages][barriers][tmem ptr]; base rounded up to 1024 B for SWIZZLE_128B
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_bytes = a.BN * kChunkK * 2;
  uint8_t* smem_a = smem;
  const int kStages = a.stages;
  const int a_stage_bytes = a.pair ? 2 * kABytes : kABytes;
  const int a_ring_bytes = a.halo ? kStages * a.halo_bytes : kStages * a_stage_bytes;
  uint8_t* smem_b = smem + a_ring_bytes;
  const int b_ring_bytes = a.halo ? a.KH * a.KW * a.kchunks * b_bytes
                                  : (a.resident_b ? a.n_tiles * a.kchunks * b_bytes : kStages * b_bytes);
  uint8_t* smem_epi = smem_b + b_ring_bytes;                   // kEpiWarps x epi_bufs x 4 KiB staging tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + kEpiWarps * a.epi_bufs * kEpiStageBytes);
  uint64_t* full = bars;                                       // [kMaxStages]
  uint64_t* empty = bars + kMaxStages;                         // [kMaxStages]
  uint64_t* tfull = bars + 2 * kMaxStages;                     // [4]
  uint64_t* tempty = bars + 2 * kMaxStages + 4;                // [4]
  uint64_t* rbar = bars + 2 * kMaxStages + 8;                  // [kEpiWarps][2] residual-tile barriers
  uint64_t* wfull = rbar + 2 * kEpiWarps;                      // halo / resident-B mode: the weights have landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(wfull + 1);
  uint64_t* lnbar = bars + 48;                                 // kLN: [4 quadrants][2 sets][3 sub-blocks] residual tiles
  float2* ln_xch = reinterpret_cast<float2*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [2][kEpiWarps][32] (kLN only)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_w);
    if (kTma) {
      tma_prefetch_desc(&map_o);
      if (kR2) tma_prefetch_desc(&map_r2);
      if (kR1) tma_prefetch_desc(&map_r1);
      if (kLN) tma_prefetch_desc(&map_o2);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2 * kEpiWarps; ++i) mbar_init(&rbar[i], 1);
    mbar_init(wfull, 1);
    if (kLN)
      for (int i = 0; i < 24; ++i) mbar_init(&lnbar[i], 1);
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], (kTma && !kLN && a.alt_items) ? kEpiWarps / 2 : kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc(tmem_ptr, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_wait();          // everything above overlapped the previous kernel's tail; global memory is touched from here on
  pdl_launch();

  const int taps = a.KH * a.KW;
  const int ksteps = taps * a.kchunks;
  const int tiles_per_img = a.tiles_y * a.tiles_x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (a.halo) {                            // resident weights: every (chunk, tap) slice once, one barrier
        mbar_expect_tx(wfull, (uint32_t)(taps * a.kchunks * b_bytes));
        for (int kc = 0; kc < a.kchunks; ++kc)
          for (int tap = 0; tap < taps; ++tap)
            tma_load_3d(smem_b + (kc * taps + tap) * b_bytes, &map_w, wfull, kc * kChunkK, 0, tap);
      }
      if (a.resident_b) {                      // resident weights: every (n tile, chunk) slice once, one barrier
        mbar_expect_tx(wfull, (uint32_t)(a.n_tiles * a.kchunks * b_bytes));
        for (int nt = 0; nt < a.n_tiles; ++nt)
          for (int kc = 0; kc < a.kchunks; ++kc)
            tma_load_3d(smem_b + (nt * a.kchunks + kc) * b_bytes, &map_w, wfull, kc * kChunkK, nt * a.BN, 0);
      }
      for (int sq = 0;; ++sq) {
        int mt, nt;
        if (!conv_item(a, sq, mt, nt)) break;
        if (a.resident_b && nt != 0) continue;   // the A chunks of a pixel tile are loaded once for all its n tiles
        int b, tyi, txi;
        conv_tile_coords(a, mt, tiles_per_img, b, tyi, txi);
        const int oy0 = tyi * a.TH;
        const int ox0 = txi * a.TW;
        int img = b, org_y = 0, org_x = 0;
        if (a.origin) { img = a.origin[3 * b]; org_y = a.origin[3 * b + 1]; org_x = a.origin[3 * b + 2]; }
        if (a.halo) {
          for (int kc = 0; kc < a.kchunks; ++kc) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], a.halo_tx);
            tma_load_4d(smem_a + stage * a.halo_bytes, &map_a, &full[stage], kc * kChunkK, org_x + ox0 - a.pad,
                        org_y + oy0 - a.pad, img);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          continue;
        }
        if (a.pair) {
          if (sq & 1) continue;                  // one producer pass per pair: item sq + 1 is the second pixel tile
          int b1, ty1, tx1;
          conv_tile_coords(a, mt + 1, tiles_per_img, b1, ty1, tx1);
          // (tap, kc, ky, kx) advance as counters: four runtime divisions per stage were most of the producer's work
          for (int ks = 0, tap = 0, kc = 0, ky = 0, kx = 0; ks < ksteps; ++ks) {
            mbar_wait(&empty[stage], phase ^ 1);
            mbar_expect_tx(&full[stage], 2 * kABytes + b_bytes);
            uint8_t* sa = smem_a + stage * a_stage_bytes;
            tma_load_4d(sa, &map_a, &full[stage], kc * kChunkK, ox0 * a.stride - a.pad + kx * a.dil,
                        oy0 * a.stride - a.pad + ky * a.dil, b);
            tma_load_4d(sa + kABytes, &map_a, &full[stage], kc * kChunkK, tx1 * a.TW * a.stride - a.pad + kx * a.dil,
                        ty1 * a.TH * a.stride - a.pad + ky * a.dil, b1);
            tma_load_3d(smem_b + stage * b_bytes, &map_w, &full[stage], kc * kChunkK, nt * a.BN, tap);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
            if (++kc == a.kchunks) { kc = 0; ++tap; if (++kx == a.KW) { kx = 0; ++ky; } }
          }
          continue;
        }
        for (int ks = 0, tap = 0, kc = 0, ky = 0, kx = 0; ks < ksteps; ++ks) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], a.resident_b ? kABytes : kABytes + b_bytes);
          tma_load_4d(smem_a + stage * kABytes, &map_a, &full[stage], kc * kChunkK,
                      org_x + ox0 * a.stride - a.pad + kx * a.dil, org_y + oy0 * a.stride - a.pad + ky * a.dil, img);
          if (!a.resident_b)
            tma_load_3d(smem_b + stage * b_bytes, &map_w, &full[stage], kc * kChunkK, nt * a.BN,
                        (a.w_batched ? b * taps : 0) + tap);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
          if (++kc == a.kchunks) { kc = 0; ++tap; if (++kx == a.KW) { kx = 0; ++ky; } }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc_bf16(kTileM, a.BN, 0, 0, a.in_fp16, a.in_fp16);
    // K = 16 steps that hold real channels in the LAST 64-channel chunk (the rest of the chunk is TMA zero fill: Ci = 48
    // needs 3 of the 4 steps, Ci = 96 needs 4 + 2): skipping the all-zero steps saves a quarter of the MMAs at Ci = 48 / 96
    const int klast = (a.Ci - (a.kchunks - 1) * kChunkK + 15) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    // accumulator index / barrier parity of item `it` (= it % nacc, (it / nacc) & 1) kept as counters: the runtime
    // divisions sat between two items' MMAs on the issuing warp
    int acc_c = 0;
    uint32_t acc_ph = 0;
    auto acc_next = [&]() { if (++acc_c == a.nacc) { acc_c = 0; acc_ph ^= 1; } };
    if (a.resident_b) {
      // pixel-tile major: the kchunks A stages of a pixel tile stay in the ring while every n tile is multiplied with
      // them; they are released by the commit that follows the LAST n tile's MMAs
      mbar_wait(wfull, 0);
      for (int pt = 0;; ++pt) {
        if (blockIdx.x + pt * (int)gridDim.x >= a.m_tiles) break;
        const int st0 = stage;
        const uint32_t ph0 = phase;
        for (int nt = 0; nt < a.n_tiles; ++nt, ++it, acc_next()) {
          const int acc = acc_c;
          const uint32_t acc_phase = acc_ph;
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * a.BN;
          int st = st0;
          uint32_t ph = ph0;
          for (int kc = 0; kc < a.kchunks; ++kc) {
            if (nt == 0) {
              mbar_wait(&full[st], ph);
              tc_fence_after();
            }
            if (elect_one()) {
              const uint64_t da = umma_desc_sw128(smem_u32(smem_a + st * kABytes), 0, 1024);
              const uint64_t db = umma_desc_sw128(smem_u32(smem_b + (nt * a.kchunks + kc) * b_bytes), 0, 1024);
#pragma unroll
              for (int k = 0; k < kChunkK / 16; ++k)
                if (kc < a.kchunks - 1 || k < klast) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (kc | k) != 0);
              if (nt == a.n_tiles - 1) umma_commit(&empty[st]);
              if (kc == a.kchunks - 1) umma_commit(&tfull[acc]);
            }
            __syncwarp();
            if (++st == kStages) { st = 0; ph ^= 1; }
          }
          if (nt == a.n_tiles - 1) { stage = st; phase = ph; }
        }
      }
    }
    for (; !a.resident_b; ++it, acc_next()) {
      int mt_, nt_;
      if (!conv_item(a, it, mt_, nt_)) break;
      const int acc = acc_c;
      const uint32_t acc_phase = acc_ph;
      if (a.pair && (it & 1)) continue;          // the pair's second item was issued with the first (accumulator acc)
      mbar_wait(&tempty[acc], acc_phase ^ 1);          // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * a.BN;
      if (a.halo) {
        if (it == 0) mbar_wait(wfull, 0);
        for (int kc = 0; kc < a.kchunks; ++kc) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            // tap (ky, kx) = the same haloed tile shifted by (ky*dil) rows of halo_pitch px and (kx*dil) px; every
            // 8-pixel UMMA row group is one image-row segment, groups are halo_pitch * 128 B apart.  The 128 B
            // swizzle is a function of the absolute smem address, so 128 B-aligned group starts need no fix-up
            // (verified on device: descriptor base-offset must stay 0).  The staged row holds exactly the 8 + (KW-1)
            // dil pixels the taps touch (a 16-px row made the L2 -> SM traffic 2.25x the useful bytes).
            // The single issuing thread is instruction-bound (ncu: tensor pipe 23 % busy at N = 64), so descriptors
            // are built ONCE per stage and advanced by adding to their 14-bit start-address field (16 B units).
            const uint64_t da0 = umma_desc_sw128(smem_u32(smem_a + stage * a.halo_bytes), 0, (uint32_t)a.halo_pitch * 128);
            uint64_t db = umma_desc_sw128(smem_u32(smem_b + kc * taps * b_bytes), 0, 1024);
            const uint32_t b_step = (uint32_t)b_bytes >> 4;
            uint32_t accum = kc != 0;
            const int nk = kc < a.kchunks - 1 ? kChunkK / 16 : klast;
            for (int ky = 0; ky < a.KH; ++ky) {
              uint64_t da = da0 + (uint64_t)((ky * a.dil * a.halo_pitch) << 3);
              for (int kx = 0; kx < a.KW; ++kx) {
#pragma unroll
                for (int k = 0; k < kChunkK / 16; ++k) {
                  if (k < nk) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, accum);
                  accum = 1;
                }
                da += (uint64_t)(a.dil << 3);
                db += b_step;
              }
            }
            umma_commit(&empty[stage]);
            if (kc == a.kchunks - 1) umma_commit(&tfull[acc]);
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      if (a.pair) {
        const int acc1 = acc + 1 == a.nacc ? 0 : acc + 1;      // the pair's second item (it + 1)
        const uint32_t acc1_phase = acc1 == 0 ? acc_phase ^ 1 : acc_phase;
        mbar_wait(&tempty[acc1], acc1_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem1 = tmem_base + acc1 * a.BN;
        for (int ks = 0, kc = 0; ks < ksteps; ++ks) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const int nk = kc < a.kchunks - 1 ? kChunkK / 16 : klast;
          if (++kc == a.kchunks) kc = 0;
          if (elect_one()) {
            const uint64_t da0 = umma_desc_sw128(smem_u32(smem_a + stage * a_stage_bytes), 0, 1024);
            const uint64_t da1 = umma_desc_sw128(smem_u32(smem_a + stage * a_stage_bytes + kABytes), 0, 1024);
            const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * b_bytes), 0, 1024);
#pragma unroll
            for (int k = 0; k < kChunkK / 16; ++k)
              if (k < nk) {
                umma_bf16(d_tmem, da0 + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
                umma_bf16(d_tmem1, da1 + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
              }
            umma_commit(&empty[stage]);
            if (ks == ksteps - 1) {
              umma_commit(&tfull[acc]);
              umma_commit(&tfull[acc1]);
            }
          }
          __syncwarp();
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      for (int ks = 0, kc = 0; ks < ksteps; ++ks) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        const int nk = kc < a.kchunks - 1 ? kChunkK / 16 : klast;     // k-step ks = (tap, chunk kc), chunk fastest
        if (++kc == a.kchunks) kc = 0;
        if (elect_one()) {
          const uint64_t da = umma_desc_sw128(smem_u32(smem_a + stage * kABytes), 0, 1024);
          const uint64_t db = umma_desc_sw128(smem_u32(smem_b + stage * b_bytes), 0, 1024);
#pragma unroll
          for (int k = 0; k < kChunkK / 16; ++k)                  // K advance = 32 B = 2 units of the address field
            if (k < nk) umma_bf16(d_tmem, da + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
          umma_commit(&empty[stage]);                   // smem stage reusable once these MMAs retire
          if (ks == ksteps - 1) umma_commit(&tfull[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue =====================
    const int ew = warp - 2;                 // 0..7
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int half = ew >> 2;                // which half of the columns (warps 2-5: 0, 6-9: 1)
    const int m = quad * 32 + lane;          // row of the tile = TMEM lane
    const int my = m / a.TW, mx = m % a.TW;
    const float g = a.scale_ptr ? *a.scale_ptr : 1.f;
    const float alpha = a.alpha * g, r1s = a.res1_scale * g;
    const int ncol16 = a.BN / 16;
    const int c_begin = (ncol16 * half) / 2, c_end = (ncol16 * (half + 1)) / 2;
    uint32_t rphase[2] = {0, 0};             // parity of this warp's residual-tile barriers
    int store_cnt = 0;
    // kLN: residual tiles of item `t` -> staging set (t & 1) of this quadrant's pool (lane 0 only)
    auto ln_prefetch = [&](int t) {
      int mt2, nt2;
      if (!conv_item(a, t, mt2, nt2)) return;
      int b2, ty2, tx2;
      conv_tile_coords(a, mt2, tiles_per_img, b2, ty2, tx2);
      const int bw = a.TW < 32 ? a.TW : 32;
      const int tx = tx2 * a.TW + (a.TW > 32 ? quad * 32 : 0);
      const int ty = ty2 * a.TH + (a.TW > 32 ? 0 : quad * (32 / bw));
      const int set = t & 1, n_sb2 = (a.Co + 31) >> 5;
      for (int sb = half; sb < n_sb2; sb += 2) {
        uint64_t* const bar = &lnbar[quad * 6 + set * 3 + sb];
        mbar_expect_tx(bar, kEpiStageBytes);
        tma_load_4d(smem_epi + (quad * 8 + set * 3 + sb) * kEpiStageBytes, &map_r2, bar, sb * 32, tx, ty, b2);
      }
    };
    if (kLN && lane == 0) ln_prefetch(0);
    int it = 0;
    int acc_c = 0;                             // it % nacc and (it / nacc) & 1 as counters
    uint32_t acc_ph = 0;
    for (;; ++it, acc_c = (acc_c + 1 == a.nacc ? 0 : acc_c + 1), acc_ph ^= (acc_c == 0)) {
      int mt, nt;
      if (!conv_item(a, it, mt, nt)) break;
      // TMA epilogue without LayerNorm: the two warps of a TMEM lane quadrant take ALTERNATE items (each does all the
      // sub-blocks of its items) instead of splitting every item, so that their fence / store / barrier latencies overlap
      if (kTma && !kLN && a.alt_items && ((it & 1) != half)) continue;
      const int acc = acc_c;
      const uint32_t acc_phase = acc_ph;
      int b, tyi, txi;
      conv_tile_coords(a, mt, tiles_per_img, b, tyi, txi);
      const int oy = tyi * a.TH + my;
      const int ox = txi * a.TW + mx;
      const bool valid = (oy < a.OH) && (ox < a.OW);
      const long long pix = ((long long)b * a.OH + oy) * a.OW + ox;
      const float rs = (valid && a.rowscale) ? a.rowscale[pix] : 1.f;

      const uint32_t t_base = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * a.BN;
      if constexpr (kLN) {
        // TMA epilogue + LayerNorm of the finished rows (Co <= 96, one N tile, <= 3 fp32 sub-blocks of 32 columns).
        // The two warps of a TMEM lane quadrant share a pool of staging tiles: two SETS of one fp32 tile per sub-block
        // (residual in, finished row out -- it stays in shared memory after its TMA store was issued) and one bf16 tile
        // per 64 columns.  The residual tiles of item t + 1 are fetched into the other set while item t is processed:
        // with a single set the HBM latency of the residual sat on every item's critical path.
        //   pass 1  sub-block sb belongs to warp half (sb & 1): residual add, per-row sum / sum of squares, fp32 store;
        //   A       partial sums cross through shared memory, 64-thread named barrier (also publishes the fp32 tiles);
        //   pass 2  re-read staged rows, normalise, write the bf16 tile; with an odd sub-block count the last one moves
        //           to half 1 so that both warps carry the same number of passes;
        //   B       second named barrier, then one lane per bf16 tile hands it to TMA.
        const int set = it & 1;
        uint8_t* const pool = smem_epi + quad * 8 * kEpiStageBytes;   // [set * 3 + sb] fp32 tiles, [6 + j] bf16 tiles
        uint8_t* const fp = pool + set * 3 * kEpiStageBytes;
        uint64_t* const rb = lnbar + quad * 6 + set * 3;
        const bool plain = !a.bias && !a.rowscale && !a.act && alpha == 1.f;
        const int box_w = a.TW < 32 ? a.TW : 32;
        const int tx0 = txi * a.TW + (a.TW > 32 ? quad * 32 : 0);
        const int ty0 = tyi * a.TH + (a.TW > 32 ? 0 : quad * (32 / box_w));
        const int n_sb = (a.Co + 31) >> 5;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        float s1 = 0.f, s2 = 0.f;
        int n_p1 = 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int sb = half + 2 * k;
          if (sb < n_sb) {
            ++n_p1;
            const int cs = sb * 32;
            const uint32_t stg_s = smem_u32(fp + sb * kEpiStageBytes);
            uint32_t raw[2][16];
            tmem_ld16(t_base + cs, raw[0]);
            tmem_ld16(t_base + cs + 16, raw[1]);
            tmem_ld_wait();
            mbar_wait(&rb[sb], (it >> 1) & 1);
#pragma unroll
            for (int q4 = 0; q4 < 2; ++q4) {
              const int col0 = cs + q4 * 16;
              float v[16];
              if (plain) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[q4][i]);
              } else {
                float bb[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (a.bias && col0 + i4 * 4 < a.Co) t = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + i4 * 4));
                  bb[i4 * 4] = t.x; bb[i4 * 4 + 1] = t.y; bb[i4 * 4 + 2] = t.z; bb[i4 * 4 + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(raw[q4][i]), rs, bb[i]);
                if (a.act == 1) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (alpha != 1.f) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] *= alpha;
                }
              }
#pragma unroll
              for (int hh = 0; hh < 4; ++hh) {
                const int off = lane * 128 + ((((q4 * 4 + hh) & 7) ^ (lane & 7)) << 4);
                const uint4 t = lds128(stg_s + off);
                float4 o = make_float4(v[hh * 4] + __uint_as_float(t.x), v[hh * 4 + 1] + __uint_as_float(t.y),
                                       v[hh * 4 + 2] + __uint_as_float(t.z), v[hh * 4 + 3] + __uint_as_float(t.w));
                // columns >= Co are exact zeros (zero-filled weights and residual, bias skipped): no mask needed
                s1 += (o.x + o.y) + (o.z + o.w);
                s2 = fmaf(o.x, o.x, s2); s2 = fmaf(o.y, o.y, s2); s2 = fmaf(o.z, o.z, s2); s2 = fmaf(o.w, o.w, s2);
                sts128(stg_s + off, __float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w));
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&map_o, fp + sb * kEpiStageBytes, cs, tx0, ty0, b);
              tma_store_commit();
            }
          }
        }
        if (lane == 0) {
          // everything this lane stored before this item (the other set's fp32 tiles, its bf16 tile) has been read out;
          // the partner finished reading the other set before barrier B of the previous item -> refill it for item + 1
          if (n_p1 == 2) tma_store_wait_read2();
          else if (n_p1 == 1) tma_store_wait_read1();
          else tma_store_wait_read();
          ln_prefetch(it + 1);
        }
        // A: row statistics, mine + the partner warp's (same quadrant, other sub-blocks)
        float2* const xq = ln_xch + set * (kEpiWarps * 32);
        xq[ew * 32 + lane] = make_float2(s1, s2);
        __syncwarp();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        const float2 other = xq[(ew ^ 4) * 32 + lane];
        s1 += other.x;
        s2 += other.y;
        const float inv_c = 1.f / (float)a.Co;
        const float mean = s1 * inv_c;
        const float rstd = rsqrtf(fmaxf(fmaf(-mean, mean, s2 * inv_c), 0.f) + a.ln_eps);
        const float sub = a.ln_mode == 1 ? mean : 0.f;             // BiasFree divides x itself (R:184-186)
        // pass 2: sub-block sb is normalised by warp half (sb & 1), except that the last of an odd count goes to half 1
#pragma unroll
        for (int sb = 0; sb < 3; ++sb) {
          const int owner = (sb == n_sb - 1 && (n_sb & 1)) ? 1 : (sb & 1);
          if (sb < n_sb && owner == half) {
            const uint32_t stg_s = smem_u32(fp + sb * kEpiStageBytes);
            const uint32_t stg16 = smem_u32(pool + (6 + (sb >> 1)) * kEpiStageBytes);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {                       // 8 columns = two fp32 chunks -> one bf16 chunk
              const int col = sb * 32 + c8 * 8;
              float y[8];
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const uint4 t = lds128(stg_s + lane * 128 + ((((c8 * 2 + h2) & 7) ^ (lane & 7)) << 4));
                float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (col + h2 * 4 < a.Co) {
                  w4 = __ldg(reinterpret_cast<const float4*>(a.ln_w + col + h2 * 4));
                  if (a.ln_b) b4 = __ldg(reinterpret_cast<const float4*>(a.ln_b + col + h2 * 4));
                }
                y[h2 * 4] = fmaf((__uint_as_float(t.x) - sub) * rstd, w4.x, b4.x);
                y[h2 * 4 + 1] = fmaf((__uint_as_float(t.y) - sub) * rstd, w4.y, b4.y);
                y[h2 * 4 + 2] = fmaf((__uint_as_float(t.z) - sub) * rstd, w4.z, b4.z);
                y[h2 * 4 + 3] = fmaf((__uint_as_float(t.w) - sub) * rstd, w4.w, b4.w);
              }
              sts128(stg16 + lane * 128 + ((((sb & 1) * 4 + c8) ^ (lane & 7)) << 4), pack2t<kH16>(y[0], y[1]),
                     pack2t<kH16>(y[2], y[3]), pack2t<kH16>(y[4], y[5]), pack2t<kH16>(y[6], y[7]));
            }
          }
        }
        // B: both warps' bf16 chunks are in place (and visible to the async proxy) -> one lane per bf16 tile stores it
        fence_proxy_async();
        __syncwarp();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
        if (lane == 0 && half * 64 < a.Co) {
          tma_store_4d(&map_o2, pool + (6 + half) * kEpiStageBytes, half * 64, tx0, ty0, b);
          tma_store_commit();
        }
        __syncwarp();
      } else if constexpr (kTma) {
        // TMA epilogue.  Sub-blocks of 128 B per pixel row (64 bf16 / 32 fp32 columns) alternate between the two
        // warps of a TMEM lane quadrant.  Residual tiles are TMA-loaded into the same SWIZZLE_128B staging tile
        // (prefetched before the accumulator is even ready), updated in place by phase 1 (thread = pixel row), and
        // the tile is handed back to the TMA engine, which clips against Co / OW / OH and writes full 128 B lines.
        constexpr int sbc = kF32 ? 32 : 64;                      // columns per sub-block
        const int nsb = a.BN / sbc;
        uint8_t* const buf = smem_epi + ew * a.epi_bufs * kEpiStageBytes;
        uint64_t* const rb = rbar + ew * 2;
        constexpr bool has_r2 = kR2, has_r1 = kR1;
        const bool dbuf = a.epi_bufs == 2 && !has_r1;
        const bool plain = !a.bias && !a.rowscale && !a.act && alpha == 1.f;
        const int box_w = a.TW < 32 ? a.TW : 32;                 // pixels per staged image row
        const int tx0 = txi * a.TW + (a.TW > 32 ? quad * 32 : 0);
        const int ty0 = tyi * a.TH + (a.TW > 32 ? 0 : quad * (32 / box_w));
        const int sb0 = a.alt_items ? 0 : half, sbs = a.alt_items ? 1 : 2;   // my sub-blocks: sb0, sb0 + sbs, ...
        int n_my = 0;
        for (int sb = sb0; sb < nsb && nt * a.BN + sb * sbc < a.Co; sb += sbs) ++n_my;
        // lane 0 only: queue the residual tile(s) of my k-th sub-block
        auto issue_res = [&](int k) {
          const int j = dbuf ? (k & 1) : 0;
          const int col = nt * a.BN + (sb0 + sbs * k) * sbc;
          tma_store_wait_read();                                 // earlier stores have finished reading the buffers
          mbar_expect_tx(&rb[j], has_r1 ? 2 * kEpiStageBytes : kEpiStageBytes);
          tma_load_4d(buf + j * kEpiStageBytes, &map_r2, &rb[j], col, tx0, ty0, b);
          if (has_r1) tma_load_4d(buf + kEpiStageBytes, &map_r1, &rb[j], col, tx0, ty0, b);
        };
        if (has_r2 && lane == 0 && n_my > 0) {
          issue_res(0);
          if (dbuf && n_my > 1) issue_res(1);
        }
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        for (int k = 0; k < n_my; ++k) {
          // staging buffer: residual tiles were queued into (k & 1); without residuals consecutive stores simply
          // alternate (across tiles too), so wait_group.read 1 always covers the buffer about to be overwritten
          int j = 0;
          if (has_r2) j = dbuf ? (k & 1) : 0;
          else if (a.epi_bufs > 1) { j = store_cnt; if (++store_cnt == a.epi_bufs) store_cnt = 0; }   // round robin
          const int cs = (sb0 + sbs * k) * sbc;
          uint8_t* const stg = buf + j * kEpiStageBytes;
          uint32_t raw[4][16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4)
            if (q4 * 16 < sbc) tmem_ld16(t_base + cs + q4 * 16, raw[q4]);
          tmem_ld_wait();
          if (has_r2) {
            mbar_wait(&rb[j], rphase[j]);
            rphase[j] ^= 1;
          } else {
            if (lane == 0) {             // the staging buffer about to be overwritten was handed to TMA epi_bufs stores ago
              if (a.epi_bufs >= 4) tma_store_wait_read3();
              else if (a.epi_bufs == 3) tma_store_wait_read2();
              else if (a.epi_bufs == 2) tma_store_wait_read1();
              else tma_store_wait_read();
            }
            __syncwarp();
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            if (q4 * 16 < sbc) {
              const int col0 = nt * a.BN + cs + q4 * 16;
              float v[16];
              if (plain) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[q4][i]);
              } else {
                float bb[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                  float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (a.bias && col0 + i4 * 4 < a.Co) t = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + i4 * 4));
                  bb[i4 * 4] = t.x; bb[i4 * 4 + 1] = t.y; bb[i4 * 4 + 2] = t.z; bb[i4 * 4 + 3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaf(__uint_as_float(raw[q4][i]), rs, bb[i]);
                // activation switch hoisted out of the element loop (GELU's erf must not be if-converted into
                // every element of the ReLU / linear paths)
                if (a.act == 1) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (a.act == 2) {
#pragma unroll
                  gelu16(v);
                }
                if (alpha != 1.f) {
#pragma unroll
                  for (int i = 0; i < 16; ++i) v[i] *= alpha;
                }
              }
              const uint32_t stg_s = smem_u32(stg);
              if constexpr (kF32) {
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const int off = lane * 128 + ((((q4 * 4 + hh) & 7) ^ (lane & 7)) << 4);
                  float4 o = make_float4(v[hh * 4], v[hh * 4 + 1], v[hh * 4 + 2], v[hh * 4 + 3]);
                  if (has_r2) {
                    const uint4 t = lds128(stg_s + off);
                    o.x += __uint_as_float(t.x); o.y += __uint_as_float(t.y);
                    o.z += __uint_as_float(t.z); o.w += __uint_as_float(t.w);
                  }
                  if (has_r1) {
                    const uint4 t = lds128(smem_u32(buf + kEpiStageBytes) + off);
                    o.x = fmaf(r1s, __uint_as_float(t.x), o.x); o.y = fmaf(r1s, __uint_as_float(t.y), o.y);
                    o.z = fmaf(r1s, __uint_as_float(t.z), o.z); o.w = fmaf(r1s, __uint_as_float(t.w), o.w);
                  }
                  sts128(stg_s + off, __float_as_uint(o.x), __float_as_uint(o.y), __float_as_uint(o.z), __float_as_uint(o.w));
                }
              } else {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int off = lane * 128 + (((q4 * 2 + hh) ^ (lane & 7)) << 4);
                  if (has_r2) {
                    const uint4 t = lds128(stg_s + off);
                    const uint32_t tw[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      v[hh * 8 + 2 * i] += __uint_as_float(tw[i] << 16);
                      v[hh * 8 + 2 * i + 1] += __uint_as_float(tw[i] & 0xffff0000u);
                    }
                  }
                  if constexpr (kH16)
                    sts128(stg_s + off, pack2h(v[hh * 8], v[hh * 8 + 1]), pack2h(v[hh * 8 + 2], v[hh * 8 + 3]),
                           pack2h(v[hh * 8 + 4], v[hh * 8 + 5]), pack2h(v[hh * 8 + 6], v[hh * 8 + 7]));
                  else
                    sts128(stg_s + off, pack2(v[hh * 8], v[hh * 8 + 1]), pack2(v[hh * 8 + 2], v[hh * 8 + 3]),
                           pack2(v[hh * 8 + 4], v[hh * 8 + 5]), pack2(v[hh * 8 + 6], v[hh * 8 + 7]));
                }
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&map_o, stg, nt * a.BN + cs, tx0, ty0, b);
            tma_store_commit();
            if (has_r2) {
              if (dbuf) { if (k + 2 < n_my) issue_res(k + 2); }
              else if (k + 1 < n_my) issue_res(k + 1);
            }
          }
          __syncwarp();
        }
      } else if (a.store_mode == 0) {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        // Coalesced stores through a per-warp staging tile: phase 1 (thread = pixel row) applies the column-wise
        // epilogue and writes 16 B chunks, XOR-swizzled by (row & 7); phase 2 re-reads them with 8 lanes per row so
        // that every global access of the warp is 4 full 128 B lines (residual loads included).
        uint8_t* stg = smem_epi + ew * kEpiStageBytes;
        const bool direct16 = a.out_bf16 && !a.out_f32 && !a.res1 && !a.res2;
        const int sub_cols = direct16 ? 64 : 32;
        const int col_hi = c_end * 16;
        for (int cs = c_begin * 16; cs < col_hi; cs += sub_cols) {
          const int ncols = min(sub_cols, col_hi - cs);
          const int n16 = ncols >> 4;
          uint32_t raw[4][16];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (j < n16) tmem_ld16(t_base + cs + j * 16, raw[j]);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (j < n16) {
              const int col0 = nt * a.BN + cs + j * 16;
              float v[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float x = __uint_as_float(raw[j][i]) * rs;
                if (a.bias && col0 + i < a.Co) x += a.bias[col0 + i];
                v[i] = x;
              }
              if (a.act == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
              } else if (a.act == 2) {
#pragma unroll
                gelu16(v);
              }
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] *= alpha;
              if (direct16) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                  const int ch = j * 2 + hh;
                  const float* vv = v + hh * 8;
                  if (a.out_fp16)
                    *reinterpret_cast<uint4*>(stg + lane * 128 + ((ch ^ (lane & 7)) << 4)) =
                        make_uint4(pack2h(vv[0], vv[1]), pack2h(vv[2], vv[3]), pack2h(vv[4], vv[5]), pack2h(vv[6], vv[7]));
                  else
                    *reinterpret_cast<bf16x8*>(stg + lane * 128 + ((ch ^ (lane & 7)) << 4)) = pack8(vv);
                }
              } else {
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                  const int ch = j * 4 + hh;        // j < 2 here (32 columns per sub-block)
                  *reinterpret_cast<float4*>(stg + lane * 128 + (((ch & 7) ^ (lane & 7)) << 4)) =
                      make_float4(v[hh * 4], v[hh * 4 + 1], v[hh * 4 + 2], v[hh * 4 + 3]);
                }
              }
            }
          }
          __syncwarp();
#pragma unroll
          for (int i8 = 0; i8 < 8; ++i8) {
            const int rl = i8 * 4 + (lane >> 3), ch = lane & 7;
            const int v_r = __shfl_sync(0xffffffffu, (int)valid, rl);
            const long long pix_r = __shfl_sync(0xffffffffu, pix, rl);
            const uint4 val = *reinterpret_cast<const uint4*>(stg + rl * 128 + ((ch ^ (rl & 7)) << 4));
            if (direct16) {
              const int col = nt * a.BN + cs + ch * 8;
              if (v_r && ch * 8 < ncols && col < a.Co)
                *reinterpret_cast<uint4*>(a.out_bf16 + pix_r * a.out_bf16_ld + col) = val;
            } else {
              const int col = nt * a.BN + cs + ch * 4;
              if (v_r && ch * 4 < ncols && col < a.Co) {
                float4 o = *reinterpret_cast<const float4*>(&val);
                if (a.res1) {
                  const float4 r = *reinterpret_cast<const float4*>(a.res1 + pix_r * a.res1_ld + col);
                  o.x += r1s * r.x; o.y += r1s * r.y; o.z += r1s * r.z; o.w += r1s * r.w;
                }
                if (a.res2) {
                  if (a.res2_bf16) {
                    const uint2 r = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(a.res2) +
                                                                    pix_r * a.res2_ld + col);
                    o.x += __uint_as_float(r.x << 16); o.y += __uint_as_float(r.x & 0xffff0000u);
                    o.z += __uint_as_float(r.y << 16); o.w += __uint_as_float(r.y & 0xffff0000u);
                  } else {
                    const float4 r = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.res2) +
                                                                      pix_r * a.res2_ld + col);
                    o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w;
                  }
                }
                if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + pix_r * a.out_f32_ld + col) = o;
                if (a.out_bf16) {
                  uint2 pk;
                  pk.x = a.out_fp16 ? pack2h(o.x, o.y) : pack2(o.x, o.y);
                  pk.y = a.out_fp16 ? pack2h(o.z, o.w) : pack2(o.z, o.w);
                  *reinterpret_cast<uint2*>(a.out_bf16 + pix_r * a.out_bf16_ld + col) = pk;
                }
              }
            }
          }
          __syncwarp();
        }
      } else {
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        for (int c16 = c_begin; c16 < c_end; ++c16) {
          uint32_t raw[16];
          tmem_ld16(t_base + c16 * 16, raw);
          tmem_ld_wait();
          const int col0 = nt * a.BN + c16 * 16;
          if (valid && col0 < a.Co) {
            float v[16];
  #pragma unroll
            for (int i = 0; i < 16; ++i) {
              float x = __uint_as_float(raw[i]) * rs;
              if (a.bias && col0 + i < a.Co) x += a.bias[col0 + i];
              v[i] = x;
            }
            if (a.act == 1) {
  #pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (a.act == 2) {
  #pragma unroll
              gelu16(v);
            }
  #pragma unroll
            for (int i = 0; i < 16; ++i) v[i] *= alpha;
            if (a.store_mode == 1) {
              // PixelUnshuffle(2): out[b, oy/2, ox/2, co*4 + (oy&1)*2 + (ox&1)]
              const long long row = ((long long)b * (a.OH >> 1) + (oy >> 1)) * (a.OW >> 1) + (ox >> 1);
              const int sub = ((oy & 1) << 1) | (ox & 1);
  #pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int co = col0 + i;
                if (co < a.Co) {
                  if (a.out_f32) a.out_f32[row * a.out_f32_ld + co * 4 + sub] = v[i];
                  if (a.out_bf16)
                    reinterpret_cast<uint16_t*>(a.out_bf16)[row * a.out_bf16_ld + co * 4 + sub] = pack1r(v[i], a.out_fp16);
                }
              }
            } else {
              // PixelShuffle(2): out[b, 2*oy + i, 2*ox + j, co/4] with (i, j) = ((co%4)/2, co%2)
  #pragma unroll
              for (int sub = 0; sub < 4; ++sub) {
                const long long row = ((long long)b * (a.OH * 2) + (oy * 2 + (sub >> 1))) * (a.OW * 2) + ox * 2 + (sub & 1);
                const int cq = col0 >> 2;        // 4 consecutive output channels
                if (col0 < a.Co) {
                  float4 o = make_float4(v[sub], v[4 + sub], v[8 + sub], v[12 + sub]);
                  if (a.res2) {               // fp32 skip connection added at the shuffled location (N:372-373)
                    const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(a.res2) +
                                                                      row * a.res2_ld + cq);
                    o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
                  }
                  if (a.out_f32) *reinterpret_cast<float4*>(a.out_f32 + row * a.out_f32_ld + cq) = o;
                  if (a.out_bf16) {
                    uint2 pk;
                    pk.x = pack2r(o.x, o.y, a.out_fp16);
                    pk.y = pack2r(o.z, o.w, a.out_fp16);
                    *reinterpret_cast<uint2*>(a.out_bf16 + row * a.out_bf16_ld + cq) = pk;
                  }
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (kTma && lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Plain SIMT restatement of the same op (one thread per output element).  Used by the GPU tests to
// cross-check the tcgen05 kernel on device and selectable with impl=1 for debugging; never chosen
// implicitly.
// ------------------------------------------------------------------------------------------------
__global__ void conv_gemm_simt_kernel(const bf16* __restrict__ in, long long in_ld, int H, int W,
                                      const bf16* __restrict__ w, long long w_ld, long long w_bstride,
                                      const ConvGemmArgs a) {
  const long long total = (long long)a.B * a.OH * a.OW * a.Co;
  const int taps = a.KH * a.KW;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % a.Co);
    const long long pix = idx / a.Co;
    const int ox = (int)(pix % a.OW);
    const int oy = (int)((pix / a.OW) % a.OH);
    const int b = (int)(pix / ((long long)a.OW * a.OH));
    float acc = 0.f;
    int img = b, org_y = 0, org_x = 0;
    if (a.origin) { img = a.origin[3 * b]; org_y = a.origin[3 * b + 1]; org_x = a.origin[3 * b + 2]; }
    for (int tap = 0; tap < taps; ++tap) {
      const int iy = org_y + oy * a.stride - a.pad + (tap / a.KW) * a.dil;
      const int ix = org_x + ox * a.stride - a.pad + (tap % a.KW) * a.dil;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;     // H, W = full image extent (TMA zero-fill rule)
      const bf16* ip = in + (((long long)img * H + iy) * W + ix) * in_ld;
      const bf16* wp = w_bstride ? w + (long long)b * w_bstride + (long long)co * w_ld
                                 : w + ((long long)((a.w_batched ? b * taps : 0) + tap) * a.Co + co) * w_ld;
      for (int ci = 0; ci < a.Ci; ++ci) acc += __bfloat162float(ip[ci]) * __bfloat162float(wp[ci]);
    }
    float x = acc * (a.rowscale ? a.rowscale[pix] : 1.f);
    if (a.bias) x += a.bias[co];
    x = apply_act(x, a.act);
    const float g = a.scale_ptr ? *a.scale_ptr : 1.f;
    x *= a.alpha * g;
    long long row = pix;
    int col = co;
    if (a.store_mode == 1) {
      row = ((long long)b * (a.OH >> 1) + (oy >> 1)) * (a.OW >> 1) + (ox >> 1);
      col = co * 4 + ((oy & 1) << 1) + (ox & 1);
    } else if (a.store_mode == 2) {
      const int sub = co & 3;
      row = ((long long)b * (a.OH * 2) + oy * 2 + (sub >> 1)) * (a.OW * 2) + ox * 2 + (sub & 1);
      col = co >> 2;
    }
    if (a.res1) x += a.res1_scale * g * a.res1[row * a.res1_ld + col];
    if (a.res2)
      x += a.res2_bf16 ? __bfloat162float(reinterpret_cast<const bf16*>(a.res2)[row * a.res2_ld + col])
                       : reinterpret_cast<const float*>(a.res2)[row * a.res2_ld + col];
    if (a.out_f32) a.out_f32[row * a.out_f32_ld + col] = x;
    if (a.out_bf16) a.out_bf16[row * a.out_bf16_ld + col] = __float2bfloat16(x);
  }
}

}  // namespace

// fused output LayerNorm: what the kLN epilogue covers (see include/tdr_sm100.h)
static bool conv_ln_ok(const tdr_conv_gemm_desc* d) {
  return d && (d->ln_mode == 1 || d->ln_mode == 2) && d->impl == 0 && d->store_mode == 0 && d->KH == 1 && d->KW == 1 &&
         d->out_f32 && !d->out_bf16 && d->res2 && !d->res2_bf16 && !d->res1 && d->act != 2 && d->Co <= 96 &&
         d->Co % 8 == 0 && d->ln_weight && d->ln_out_bf16 && ((uintptr_t)d->ln_out_bf16 & 15) == 0 &&
         d->ln_out_ld % 8 == 0 && ((uintptr_t)d->out_f32 & 15) == 0 && d->out_f32_ld % 4 == 0 &&
         ((uintptr_t)d->res2 & 15) == 0 && d->res2_ld % 4 == 0 && ((uintptr_t)d->ln_weight & 15) == 0 &&
         (!d->ln_bias || ((uintptr_t)d->ln_bias & 15) == 0) && getenv("TDR_CONV_EPI") == nullptr &&
         getenv("TDR_CONV_BN") == nullptr;
}

extern "C" int tdr_conv_gemm_ln_supported(const tdr_conv_gemm_desc* d) { return conv_ln_ok(d) ? 1 : 0; }

extern "C" int tdr_conv_gemm(const tdr_conv_gemm_desc* d, cudaStream_t stream) {
  TDR_CHECK_ARG(d != nullptr, "tdr_conv_gemm: null descriptor");
  TDR_CHECK_ARG(d->in && d->weight, "tdr_conv_gemm: null input/weight");
  TDR_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->Ci > 0 && d->Co > 0, "tdr_conv_gemm: bad dims");
  TDR_CHECK_ARG(d->KH >= 1 && d->KW >= 1 && d->stride >= 1 && d->stride <= 2 && d->dil >= 1,
                "tdr_conv_gemm: bad filter geometry");
  TDR_CHECK_ARG(d->in_ld % 8 == 0 && d->w_ld % 8 == 0, "tdr_conv_gemm: in_ld/w_ld must be multiples of 8 (16 B rows)");
  TDR_CHECK_ARG(((uintptr_t)d->in & 15) == 0 && ((uintptr_t)d->weight & 15) == 0, "tdr_conv_gemm: 16 B alignment");
  TDR_CHECK_ARG(d->out_f32 || d->out_bf16, "tdr_conv_gemm: no output");
  TDR_CHECK_ARG(d->store_mode >= 0 && d->store_mode <= 2, "tdr_conv_gemm: bad store_mode");
  const int OH = (d->H + 2 * d->pad - d->dil * (d->KH - 1) - 1) / d->stride + 1;
  const int OW = (d->W + 2 * d->pad - d->dil * (d->KW - 1) - 1) / d->stride + 1;
  TDR_CHECK_ARG(OH > 0 && OW > 0, "tdr_conv_gemm: empty output");
  const int img_h = d->origin ? d->img_h : d->H, img_w = d->origin ? d->img_w : d->W;
  const int n_img = d->origin ? d->n_images : d->B;
  TDR_CHECK_ARG(img_h > 0 && img_w > 0 && n_img > 0, "tdr_conv_gemm: bad image extent for origin mode");
  if (d->store_mode == 0) {
    TDR_CHECK_ARG(d->Co % 8 == 0, "tdr_conv_gemm: Co must be a multiple of 8 (got %d)", d->Co);
  } else if (d->store_mode == 1) {
    TDR_CHECK_ARG(OH % 2 == 0 && OW % 2 == 0, "tdr_conv_gemm: pixel-unshuffle needs even output size");
    TDR_CHECK_ARG(!d->res1 && !d->res2, "tdr_conv_gemm: residuals unsupported with pixel (un)shuffle");
  } else {
    TDR_CHECK_ARG(d->Co % 16 == 0, "tdr_conv_gemm: pixel-shuffle needs Co %% 16 == 0");
    TDR_CHECK_ARG(!d->res1 && (!d->res2 || !d->res2_bf16), "tdr_conv_gemm: pixel-shuffle supports an fp32 res2 only");
  }

  ConvGemmArgs a;
  a.B = d->B; a.OH = OH; a.OW = OW; a.Ci = d->Ci; a.Co = d->Co;
  a.KH = d->KH; a.KW = d->KW; a.stride = d->stride; a.pad = d->pad; a.dil = d->dil;
  a.w_batched = d->w_batched; a.origin = d->origin;
  a.bias = d->bias; a.rowscale = d->rowscale; a.alpha = d->alpha; a.res1_scale = d->res1_scale; a.scale_ptr = d->scale_ptr; a.act = d->act;
  a.res1 = d->res1; a.res1_ld = d->res1_ld; a.res2 = d->res2; a.res2_ld = d->res2_ld; a.res2_bf16 = d->res2_bf16;
  a.out_f32 = d->out_f32; a.out_f32_ld = d->out_f32_ld;
  a.out_bf16 = reinterpret_cast<bf16*>(d->out_bf16); a.out_bf16_ld = d->out_bf16_ld;
  a.store_mode = d->store_mode;
  a.in_fp16 = d->in_fp16 ? 1 : 0; a.out_fp16 = d->out_fp16 ? 1 : 0;
  TDR_CHECK_ARG(!d->out_fp16 || ((d->out_bf16 || d->ln_out_bf16) && !d->res2_bf16),
                "tdr_conv_gemm: out_fp16 needs a 16-bit output (out_bf16 / ln_out_bf16) and no bf16 residual");
  TDR_CHECK_ARG(!d->in_fp16 || d->impl == 0, "tdr_conv_gemm: the SIMT restatement reads bf16 operands only");
  const bool want_ln = d->ln_mode != 0;
  TDR_CHECK_ARG(!want_ln || conv_ln_ok(d), "tdr_conv_gemm: fused LayerNorm needs a 1x1 op with fp32 output + fp32 res2, "
                "no res1, Co <= 96 and 16 B-aligned rows (ln_mode %d, Co %d)", d->ln_mode, d->Co);
  a.ln_mode = d->ln_mode; a.ln_eps = d->ln_eps; a.ln_w = d->ln_weight; a.ln_b = d->ln_mode == 1 ? d->ln_bias : nullptr;
  // spatial tile: 128 output pixels as TH x TW
  a.TW = OW >= 16 ? 16 : 8;
  a.TH = kTileM / a.TW;
  if (OH == 1) { a.TW = 128; a.TH = 1; }         // flat [rows, C] GEMM view
  // halo mode: stride-1 multi-tap convs fetch one haloed [TH+(KH-1)dil] x 16 px tile per channel chunk
  a.halo = 0;                                    // decided below, once the N tiling is known
  a.halo_bytes = 0;
  a.tiles_x = tdr_cdiv(OW, a.TW);
  a.tiles_y = tdr_cdiv(OH, a.TH);
  {
    // TMA epilogue: exactly one output dtype, res2 (if any) of that dtype, res1 only with fp32 output, 16 B aligned
    // rows everywhere.  Everything else goes through the generic (slower) staged path.
    const bool one_out = (d->out_bf16 != nullptr) != (d->out_f32 != nullptr);
    const bool f32o = d->out_f32 != nullptr;
    const long long old_ = f32o ? d->out_f32_ld : d->out_bf16_ld;
    const void* optr = f32o ? (const void*)d->out_f32 : d->out_bf16;
    const int per16 = f32o ? 4 : 8;
    bool ok = d->impl == 0 && d->store_mode == 0 && one_out && ((uintptr_t)optr & 15) == 0 && old_ % per16 == 0;
    if (ok && d->res2) ok = (d->res2_bf16 != 0) == !f32o && ((uintptr_t)d->res2 & 15) == 0 && d->res2_ld % per16 == 0;
    if (ok && d->res1) ok = f32o && d->res2 && ((uintptr_t)d->res1 & 15) == 0 && d->res1_ld % 4 == 0;
    if (const char* e = getenv("TDR_CONV_EPI")) {                                // tuning knob (experiments only)
      if (atoi(e) == 0) ok = false;
    }
    a.epi_mode = ok ? 1 : 0;
    a.out_is_f32 = f32o ? 1 : 0;
  }
  if (a.epi_mode == 1) {
    // N tiles are multiples of the sub-block width (64 bf16 / 32 fp32 columns, <= 256).  Cost model: padded columns
    // plus ~48 columns' worth of fixed work (A re-fetch, pipeline ramp) per extra N tile.
    const int sbc = a.out_is_f32 ? 32 : 64;
    int best_bn = 256, best_cost = 1 << 30;
    for (int bn = sbc; bn <= 256; bn += sbc) {
      const int nt = tdr_cdiv(d->Co, bn);
      const int cost = nt * bn + 48 * nt;
      if (cost < best_cost || (cost == best_cost && bn > best_bn)) { best_cost = cost; best_bn = bn; }
    }
    if (const char* e = getenv("TDR_CONV_BN")) {                                 // tuning knob (experiments only)
      const int v = atoi(e);
      if (v >= sbc && v <= 256 && v % sbc == 0) best_bn = v;
    }
    a.BN = best_bn;
    a.n_tiles = tdr_cdiv(d->Co, a.BN);
    a.epi_bufs = 2;
    a.stages = a.BN <= 192 ? 4 : 3;
    if (!d->res2 && !d->res1 && a.BN > 192) { a.epi_bufs = 1; a.stages = 4; }   // no residual tiles: favour depth
    if (const char* e = getenv("TDR_CONV_EPIBUFS")) {                            // tuning knobs (experiments only)
      const int v = atoi(e);
      if ((v == 1 && !d->res1) || v == 2 || ((v == 3 || v == 4) && !d->res1 && !d->res2)) a.epi_bufs = v;
    }
    if (const char* e = getenv("TDR_CONV_STAGES")) {
      const int v = atoi(e);
      if (v >= 2 && v <= 4) a.stages = v;
    }
    if (want_ln) a.epi_bufs = 4;                 // per TMEM quadrant (2 warps): 2 sets x 3 fp32 tiles + 2 bf16 tiles
  } else {
    // N tiling: equal tiles of at most 256 columns, multiples of 16
    const int co16 = tdr_cdiv(d->Co, 16) * 16;
    a.n_tiles = tdr_cdiv(co16, 256);
    a.BN = tdr_cdiv(tdr_cdiv(co16, a.n_tiles), 16) * 16;
    a.epi_bufs = 1;
    a.stages = 4;
  }
  a.kchunks = tdr_cdiv(d->Ci, kChunkK);
  a.resident_b = 0;
  auto smem_need = [&]() {
    const size_t bt = (size_t)a.BN * kChunkK * 2;
    const size_t ring = a.halo ? (size_t)a.stages * a.halo_bytes + (size_t)d->KH * d->KW * a.kchunks * bt
                               : (a.resident_b ? (size_t)a.stages * kABytes + (size_t)a.n_tiles * a.kchunks * bt
                                               : (size_t)a.stages * ((a.pair ? 2 : 1) * kABytes + bt));
    return 1024 + ring + (size_t)kEpiWarps * a.epi_bufs * kEpiStageBytes + 1024 + (want_ln ? 2 * kEpiWarps * 32 * 8 : 0);
  };
  // Resident-B mode (1x1 convs with shared weights and many pixel tiles per CTA): the whole weight matrix stays in shared
  // memory, the ring carries A chunks only.  BN = the multiple of the sub-block width with the fewest padded columns;
  // accepted when at least max(4, kchunks + 2) A stages fit next to the weights and the epilogue staging.
  if (a.epi_mode == 1 && d->KH * d->KW == 1 && !d->w_batched && d->impl == 0 && d->stride == 1 &&
      (long long)d->B * a.tiles_y * a.tiles_x >= 2LL * tdr_num_sms() && getenv("TDR_CONV_NO_RESIDENT") == nullptr &&
      getenv("TDR_CONV_BN") == nullptr && getenv("TDR_CONV_STAGES") == nullptr) {
    const int sbc = a.out_is_f32 ? 32 : 64;
    int bn_best = 0, pad_best = 1 << 30;
    for (int bn = sbc; bn <= 256; bn += sbc) {
      const int nt = tdr_cdiv(d->Co, bn);
      const int pad = nt * bn + 48 * nt;                        // same cost model as the streamed-weights plan above
      if (pad < pad_best || (pad == pad_best && bn > bn_best)) { pad_best = pad; bn_best = bn; }
    }
    const int nt = tdr_cdiv(d->Co, bn_best);
    const size_t bbytes = (size_t)nt * a.kchunks * bn_best * kChunkK * 2;
    const int tries[2] = {want_ln ? 4 : 2, want_ln ? 4 : (d->res1 ? 2 : 1)};
    for (int t = 0; t < 2 && !a.resident_b; ++t) {
      const long long avail = 227LL * 1024 - 1024 - 1024 - (want_ln ? 2 * kEpiWarps * 32 * 8 : 0) - (long long)bbytes -
                              (long long)kEpiWarps * tries[t] * kEpiStageBytes;
      int S = (int)(avail / kABytes);
      if (S > kMaxStages) S = kMaxStages;
      const int need = a.kchunks + 2 > 4 ? a.kchunks + 2 : 4;
      if (avail > 0 && S >= need && (!want_ln || nt == 1)) {
        a.resident_b = 1;
        a.BN = bn_best; a.n_tiles = nt; a.stages = S; a.epi_bufs = tries[t];
      }
    }
  }
  // Halo mode: stride-1 multi-tap conv, shared (not per-sample) weights that fit in shared memory next to a 3-deep
  // ring of haloed [TH + (KH-1)dil] x 16 px A tiles, a single N tile.  TW = 8 so that every 8-row UMMA group is one
  // image-row segment of the haloed tile.
  if (d->impl == 0 && d->stride == 1 && d->KH * d->KW > 1 && (d->KW - 1) * d->dil <= 8 && OH > 1 && !d->w_batched &&
      a.n_tiles == 1 && getenv("TDR_CONV_NO_HALO") == nullptr) {
    const int pitch = 8 + (d->KW - 1) * d->dil;
    const int tx = (16 + (d->KH - 1) * d->dil) * pitch * 128;
    const int halo_bytes = (tx + 1023) & ~1023;
    const size_t fixed = 1024 + (size_t)d->KH * d->KW * a.kchunks * a.BN * kChunkK * 2 +
                         (size_t)kEpiWarps * (d->res1 ? 2 : 1) * kEpiStageBytes + 1024;
    if (fixed + (size_t)3 * halo_bytes <= 227 * 1024) {
      a.halo = 1;
      a.halo_bytes = halo_bytes;
      a.halo_pitch = pitch;
      a.halo_tx = tx;
      a.TW = 8; a.TH = 16;
      a.tiles_x = tdr_cdiv(OW, a.TW); a.tiles_y = tdr_cdiv(OH, a.TH);
      int st = (int)((227 * 1024 - fixed) / halo_bytes);     // as deep a ring of haloed tiles as fits (3..6)
      a.stages = st > 6 ? 6 : st;
      a.epi_bufs = d->res1 ? 2 : 1;
    }
  }
  // Pair mode: streamed-weight multi-tap convs with many pixel tiles per CTA (dense 3x3 at C >= 96, stride 1 or 2).
  a.pair = 0;
  if (a.epi_mode == 1 && !want_ln && !a.halo && !a.resident_b && d->impl == 0 && (d->KH * d->KW > 1 || getenv("TDR_CONV_PAIR_1X1") != nullptr) && !d->w_batched &&
      !d->origin && !d->rowscale && (a.n_tiles == 1 || d->Co % a.BN == 0) && 2 * a.BN <= 512 &&
      (long long)d->B * a.tiles_y * a.tiles_x * a.n_tiles >= 2LL * tdr_num_sms() && getenv("TDR_CONV_NO_PAIR") == nullptr) {
    // BN = 256 leaves two accumulators, both owned by the pair: no MMA / epilogue overlap.  With short K (Ci <= 256) that
    // loss outweighs the saved traffic (256->256 @128^2: 126 -> 144 us), with long K it does not (512->512: 128 -> 114 us)
    a.pair = (a.BN > 192 && a.kchunks < 6) ? 0 : 1;
    if (a.pair && a.stages > 4) a.stages = 4;
  }
  const bool bufs_forced = getenv("TDR_CONV_EPIBUFS") != nullptr;
  while (smem_need() > 227 * 1024) {
    if ((bufs_forced || want_ln) && a.stages > 2) --a.stages;
    else if (a.epi_bufs >= 2 && !d->res1 && !want_ln) --a.epi_bufs;
    else if (a.stages > 2) --a.stages;
    else break;
  }
  TDR_CHECK_ARG(smem_need() <= 227 * 1024, "tdr_conv_gemm: shared-memory plan does not fit");
  TDR_CHECK_ARG(!want_ln || (a.epi_mode == 1 && a.n_tiles == 1 && !a.halo && a.epi_bufs == 4),
                "tdr_conv_gemm: fused LayerNorm plan not available for this shape");
  a.m_tiles = d->B * a.tiles_y * a.tiles_x;
  a.total_tiles = a.m_tiles * a.n_tiles;
  // by_pixel only pays when the n tiles are UNEQUAL (its scheduling granule is n_tiles items, so the tail gets coarser) and
  // there are many pixel tiles per SM
  a.by_pixel = (a.n_tiles > 1 && d->Co % a.BN != 0 && a.m_tiles >= 8 * tdr_num_sms()) ? 1 : 0;
  if (a.resident_b) a.by_pixel = 1;
  a.pair_tiles = ((a.m_tiles + 1) / 2) * a.n_tiles;
  if (a.pair && (a.by_pixel || a.stages < 3)) a.pair = 0;
  // accumulator ring: as many BN-column buffers as fit in the 512 TMEM columns (2..4) -- a deeper ring keeps more
  // tiles between the MMA issuer and the epilogue in flight
  a.nacc = 512 / a.BN;
  if (a.nacc > 4) a.nacc = 4;
  if (a.nacc < 2) a.nacc = 2;
  if (const char* e = getenv("TDR_CONV_NACC")) {                                 // tuning knob (experiments only)
    const int v = atoi(e);
    if (v >= 2 && v <= 4 && v * a.BN <= 512) a.nacc = v;
  }
  uint32_t cols = 32;
  while (cols < (uint32_t)(a.nacc * a.BN)) cols <<= 1;
  a.tmem_cols = cols;
  // (same-box bench step: 50.92 ms split items, 50.35 ms alternate for BN <= 192 only, 50.11 ms alternate everywhere)
  a.alt_items = (a.epi_mode == 1 && !want_ln) ? 1 : 0;
  tdr_fast_div_setup(a.n_tiles, &a.fd_nt_m, &a.fd_nt_s);
  tdr_fast_div_setup(a.tiles_y * a.tiles_x, &a.fd_tpi_m, &a.fd_tpi_s);
  tdr_fast_div_setup(a.tiles_x, &a.fd_tx_m, &a.fd_tx_s);
  if (const char* e = getenv("TDR_CONV_ALT")) a.alt_items = a.alt_items && atoi(e) != 0;   // A/B knob

  if (d->impl == 1) {
    const long long total = (long long)a.B * OH * OW * a.Co;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    conv_gemm_simt_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const bf16*>(d->in), d->in_ld, img_h, img_w,
                                                      reinterpret_cast<const bf16*>(d->weight), d->w_ld,
                                                      d->w_batched ? d->w_batch_stride : 0, a);
    TDR_CHECK_LAUNCH();
    return TDR_OK;
  }

  TDR_CHECK_ARG(a.TW * d->stride <= 256 && a.TH * d->stride <= 256, "tdr_conv_gemm: TMA box too large");
  TdrTensorMap map_a, map_w, map_o, map_r2, map_r1, map_o2;
  memset(&map_o2, 0, sizeof(map_o2));
  memset(&map_o, 0, sizeof(map_o));
  memset(&map_r2, 0, sizeof(map_r2));
  memset(&map_r1, 0, sizeof(map_r1));
  if (a.epi_mode == 1) {
    const int box_w = a.TW < 32 ? a.TW : 32;
    const int esz = a.out_is_f32 ? 4 : 2;
    const uint64_t dims[4] = {(uint64_t)d->Co, (uint64_t)OW, (uint64_t)OH, (uint64_t)d->B};
    const uint32_t box[4] = {(uint32_t)(128 / esz), (uint32_t)box_w, (uint32_t)(32 / box_w), 1};
    const uint32_t es[4] = {1, 1, 1, 1};
    auto mk = [&](TdrTensorMap* m, const void* ptr, long long ld) {
      const uint64_t strides[3] = {(uint64_t)ld * esz, (uint64_t)ld * esz * OW, (uint64_t)ld * esz * OW * OH};
      return a.out_is_f32 ? tdr_make_tensor_map_f32(m, ptr, 4, dims, strides, box, es)
                          : tdr_make_tensor_map_bf16(m, ptr, 4, dims, strides, box, es);
    };
    int rc = a.out_is_f32 ? mk(&map_o, d->out_f32, d->out_f32_ld) : mk(&map_o, d->out_bf16, d->out_bf16_ld);
    if (rc) return rc;
    if (d->res2 && (rc = mk(&map_r2, d->res2, d->res2_ld))) return rc;
    if (d->res1 && (rc = mk(&map_r1, d->res1, d->res1_ld))) return rc;
    if (want_ln) {
      const uint32_t box16[4] = {64, (uint32_t)box_w, (uint32_t)(32 / box_w), 1};
      const uint64_t st16[3] = {(uint64_t)d->ln_out_ld * 2, (uint64_t)d->ln_out_ld * 2 * OW,
                                (uint64_t)d->ln_out_ld * 2 * OW * OH};
      if ((rc = tdr_make_tensor_map_bf16(&map_o2, d->ln_out_bf16, 4, dims, st16, box16, es))) return rc;
    }
  }
  {
    const uint64_t dims[4] = {(uint64_t)d->Ci, (uint64_t)img_w, (uint64_t)img_h, (uint64_t)n_img};
    const uint64_t strides[3] = {(uint64_t)d->in_ld * 2, (uint64_t)d->in_ld * 2 * img_w,
                                 (uint64_t)d->in_ld * 2 * img_w * img_h};
    uint32_t box[4] = {(uint32_t)kChunkK, (uint32_t)(a.TW * d->stride), (uint32_t)(a.TH * d->stride), 1};
    if (a.halo) { box[1] = (uint32_t)a.halo_pitch; box[2] = (uint32_t)(a.TH + (d->KH - 1) * d->dil); }
    const uint32_t es[4] = {1, (uint32_t)d->stride, (uint32_t)d->stride, 1};
    int rc = tdr_make_tensor_map_bf16(&map_a, d->in, 4, dims, strides, box, es);
    if (rc) return rc;
  }
  {
    const int taps = d->KH * d->KW;
    const uint64_t dims[3] = {(uint64_t)d->Ci, (uint64_t)d->Co, (uint64_t)taps * (d->w_batched ? d->B : 1)};
    uint64_t tap_stride = (uint64_t)d->w_ld * 2 * d->Co;
    if (d->w_batched && d->w_batch_stride) {
      TDR_CHECK_ARG(taps == 1 && d->w_batch_stride % 8 == 0, "tdr_conv_gemm: w_batch_stride needs a 1x1 op and 16 B rows");
      tap_stride = (uint64_t)d->w_batch_stride * 2;
    }
    const uint64_t strides[2] = {(uint64_t)d->w_ld * 2, tap_stride};
    const uint32_t box[3] = {(uint32_t)kChunkK, (uint32_t)a.BN, 1};
    const uint32_t es[3] = {1, 1, 1};
    int rc = tdr_make_tensor_map_bf16(&map_w, d->weight, 3, dims, strides, box, es);
    if (rc) return rc;
  }
  const size_t smem = smem_need();
  if (getenv("TDR_CONV_DEBUG"))
    fprintf(stderr, "conv_gemm plan: Ci %d Co %d k%d BN %d n_tiles %d stages %d epi_bufs %d halo %d resident %d pair %d by_pixel %d nacc %d\n",
            d->Ci, d->Co, d->KH, a.BN, a.n_tiles, a.stages, a.epi_bufs, a.halo, a.resident_b, a.pair, a.by_pixel, a.nacc);
  int grid = a.total_tiles < tdr_num_sms() ? a.total_tiles : tdr_num_sms();
  if (a.by_pixel && grid > a.m_tiles) grid = a.m_tiles;
  if (a.pair && grid > a.pair_tiles) grid = a.pair_tiles;
  int variant = a.epi_mode == 1 ? (a.out_is_f32 | (d->res2 ? 2 : 0) | (d->res1 ? 4 : 0) | (want_ln ? 8 : 0)) : -1;
  if (variant == 0 && a.out_fp16) variant = 16;
  if (variant == 11 && a.out_fp16) variant = 27;
#define TDR_LAUNCH_CONV(VV)                                                                                        \
  do {                                                                                                             \
    static bool attr_set = false;                                                                                  \
    if (!attr_set) {                                                                                               \
      TDR_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<VV>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                                          227 * 1024));                                                            \
      attr_set = true;                                                                                             \
    }                                                                                                              \
    TDR_CHECK_CUDA(tdr_launch_pdl(conv_gemm_kernel<VV>, dim3(grid), dim3(kThreads), smem, stream, map_a, map_w,    \
                                  map_o, map_r2, map_r1, map_o2, a));                                              \
  } while (0)
  switch (variant) {
    case 0: TDR_LAUNCH_CONV(0); break;
    case 1: TDR_LAUNCH_CONV(1); break;
    case 2: TDR_LAUNCH_CONV(2); break;
    case 3: TDR_LAUNCH_CONV(3); break;
    case 7: TDR_LAUNCH_CONV(7); break;
    case 11: TDR_LAUNCH_CONV(11); break;
    case 16: TDR_LAUNCH_CONV(16); break;
    case 27: TDR_LAUNCH_CONV(27); break;
    default: TDR_LAUNCH_CONV(-1); break;
  }
#undef TDR_LAUNCH_CONV
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
