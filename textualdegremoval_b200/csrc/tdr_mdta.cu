// MDTA (multi-Dconv-head transposed attention) score path, /root/reference/models/archs/network_restormer_guided_arch.py:262-276.
//
//   q, k : [heads, c, P] per sample (c = 48 or 96 channels per head, P = H*W pixels, up to 262144)
//   attn = softmax( normalize(q) @ normalize(k)^T * temperature )           -- a c x c matrix per head
//   out  = project_out( attn @ v )
//
// The contraction runs over PIXELS, so with NHWC activations both operands are "MN-major" for UMMA: a TMA box of
// [64 channels x 128 pixels] lands in shared memory as 128 rows (K) of 128 B (64 channels of M/N), SWIZZLE_128B, and is
// consumed directly by tcgen05.mma with a_major = b_major = MN.  Three products share each staged tile:
//   D0 = q^T k (the Gram), D1 = q^T q, D2 = k^T k  -- the diagonals of D1/D2 are the squared L2 norms that
// F.normalize (:266-267) needs, so q and k are read exactly once and no separate norm pass exists.
// The pixel axis is split over CTAs; each writes an fp32 partial (deterministic two-stage reduction, no atomics).
// tdr_mdta_weff then reduces, applies softmax and folds attn into project_out:  Weff = W_out * blockdiag(attn), so that
// `attn @ v` + project_out (:272-276) becomes one tdr_conv_gemm with per-sample weights.
#include "tdr_common.cuh"

namespace {

constexpr int kPixTile = 128;              // pixels (K) per pipeline stage
constexpr int kBoxBytes = 64 * kPixTile * 2;

struct GramPlan {
  int c, M, N, boxes, stages, nchunks, tiles_per_chunk, tiles_total;
  int box_ch, box_bytes;  // channels / bytes per TMA box: 64 ch SWIZZLE_128B, or 32 ch SWIZZLE_64B when that tiles c exactly
                          // (c = 96: three 32-channel boxes read exactly q and k; two 64-channel boxes over-read 1.33x --
                          // ncu measured 570 MB of DRAM reads for 403 MB of q, k)
  uint32_t tmem_cols;
  int generic;            // heads wider than 128 channels (PromptIR's prompt-interaction blocks, c = 176): SIMT Gram
};

constexpr int kGenChunkPix = 2048;     // pixels per partial of the generic Gram (a function of P only)

static int make_plan(int B, long long P, int C, int heads, GramPlan* p) {
  if (heads <= 0 || C % heads != 0) return -1;
  p->c = C / heads;
  p->generic = 0;
  if (p->c % 8 == 0 && p->c > 128 && p->c <= 224) {       // 224: the fold kernel keeps c x c floats in shared memory
    // Wide heads occur only at the coarse levels (a few thousand pixels): a shared-memory tiled SIMT Gram is enough.
    p->generic = 1;
    p->M = p->N = p->c; p->boxes = 0; p->stages = 0; p->tmem_cols = 0;
    p->tiles_total = (int)((P + kGenChunkPix - 1) / kGenChunkPix);
    p->tiles_per_chunk = 1;
    p->nchunks = p->tiles_total;
    return 0;
  }
  if (p->c % 8 != 0 || p->c > 128) return -1;
  if (p->c > 64 && p->c % 16 != 0) return -1;
  p->M = p->c <= 64 ? 64 : 128;
  p->N = p->c;
  p->boxes = p->c <= 64 ? 1 : 2;
  p->stages = p->boxes == 1 ? 4 : 3;
  p->box_ch = 64;
  p->box_bytes = kBoxBytes;
  // Measured and NOT adopted (opt-in knob): the exact-width boxes cut the DRAM reads but the kernel gets slower (156 vs 138 us
  // at C = 96, 512^2, batch 4): 1.5x the TMA requests on 64-byte rows cost more than the saved 1.33x over-read.
  if (p->c == 96 && getenv("TDR_GRAM_SW64") != nullptr) {
    p->box_ch = 32;
    p->box_bytes = 32 * kPixTile * 2;
    p->boxes = 3;
    p->stages = 4;
  }
  p->tiles_total = (int)((P + kPixTile - 1) / kPixTile);
  // The split of the pixel axis is a function of (P, heads) ONLY, never of the batch size: the fp32 partial sums are
  // then combined in the same order whether a sample is run alone or inside a batch, so a sample's output does not
  // depend on its batch mates (a different summation order moves the Gram by ~1e-7, which flips bf16 roundings of
  // Weff downstream and showed up as a 4.5e-3 batch-size dependence of the 512x512 output).  num_sms / heads chunks
  // per (sample, head) make every sample one full wave of CTAs.
  (void)B;
  int want = tdr_num_sms() / heads;
  if (want < 1) want = 1;
  if (want > p->tiles_total) want = p->tiles_total;
  p->tiles_per_chunk = (p->tiles_total + want - 1) / want;
  p->nchunks = (p->tiles_total + p->tiles_per_chunk - 1) / p->tiles_per_chunk;
  uint32_t cols = 32;
  while (cols < (uint32_t)(3 * p->N)) cols <<= 1;
  p->tmem_cols = cols;
  return 0;
}

struct GramArgs {
  int B, C, heads;
  long long P;
  GramPlan plan;
  float* partials;
  int fp16;                     // q, k are IEEE fp16 (inference forward) instead of bf16
};

__global__ void __launch_bounds__(192, 1) mdta_gram_kernel(const __grid_constant__ TdrTensorMap map, const GramArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const GramPlan& pl = a.plan;
  const int stage_bytes = 2 * pl.boxes * pl.box_bytes;       // q boxes then k boxes
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + pl.stages * stage_bytes + pl.box_bytes);   // after the over-read pad
  uint64_t* full = bars;
  uint64_t* empty = bars + pl.stages;
  uint64_t* tfull = bars + 2 * pl.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * pl.stages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int tile0 = chunk * pl.tiles_per_chunk;
  int ntiles = pl.tiles_total - tile0;
  if (ntiles > pl.tiles_per_chunk) ntiles = pl.tiles_per_chunk;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map);
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
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int pix0 = (tile0 + t) * kPixTile;
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_expect_tx(&full[stage], stage_bytes);
        uint8_t* base = smem + stage * stage_bytes;
        for (int j = 0; j < pl.boxes; ++j) {
          tma_load_3d(base + j * pl.box_bytes, &map, &full[stage], h * pl.c + pl.box_ch * j, pix0, b);
          tma_load_3d(base + (pl.boxes + j) * pl.box_bytes, &map, &full[stage], a.C + h * pl.c + pl.box_ch * j, pix0, b);
        }
        if (++stage == pl.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(pl.M, pl.N, 1, 1, a.fp16, a.fp16);
    int stage = 0;
    uint32_t phase = 0;
    for (int t = 0; t < ntiles; ++t) {
      mbar_wait(&full[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sq = smem_u32(smem + stage * stage_bytes);
        const uint32_t sk = sq + pl.boxes * pl.box_bytes;
        const bool sw64 = pl.box_ch == 32;
        // MN-major operands: rows (pixels) of 128 B (64 B with SWIZZLE_64B), an 8-row group is 1024 (512) B = SBO, the
        // channel chunks are box_bytes apart (LBO); 16 pixels (one K step) = 2048 (1024) B = 128 (64) address units
        const uint64_t dq0 = sw64 ? umma_desc_sw64(sq, pl.box_bytes, 512) : umma_desc_sw128(sq, pl.box_bytes, 1024);
        const uint64_t dk0 = sw64 ? umma_desc_sw64(sk, pl.box_bytes, 512) : umma_desc_sw128(sk, pl.box_bytes, 1024);
        const uint32_t kstep = sw64 ? 64 : 128;
#pragma unroll
        for (int ks = 0; ks < kPixTile / 16; ++ks) {
          const uint64_t dq = dq0 + ks * kstep, dk = dk0 + ks * kstep;
          const uint32_t accum = (t | ks) != 0;
          umma_bf16(tmem_base + 0 * pl.N, dq, dk, idesc, accum);
          umma_bf16(tmem_base + 1 * pl.N, dq, dq, idesc, accum);
          umma_bf16(tmem_base + 2 * pl.N, dk, dk, idesc, accum);
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
    const int row = pl.M == 128 ? quad * 32 + lane : quad * 16 + lane;
    const bool row_ok = (pl.M == 128 || lane < 16) && row < pl.c;
    float* out = a.partials + ((size_t)(b * a.heads + h) * pl.nchunks + chunk) * (size_t)(pl.c * pl.c + 2 * pl.c);
    mbar_wait(tfull, 0);
    tc_fence_after();
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const int nc16 = (pl.N + 15) / 16;
    float dq = 0.f, dk = 0.f;
    for (int c16 = 0; c16 < nc16; ++c16) {
      uint32_t g[16], qq[16], kk[16];
      tmem_ld16(t_lane + 0 * pl.N + c16 * 16, g);
      tmem_ld16(t_lane + 1 * pl.N + c16 * 16, qq);
      tmem_ld16(t_lane + 2 * pl.N + c16 * 16, kk);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int col = c16 * 16 + i;
        if (row_ok && col < pl.c) out[row * pl.c + col] = __uint_as_float(g[i]);
        if (col == row) {
          dq = __uint_as_float(qq[i]);
          dk = __uint_as_float(kk[i]);
        }
      }
    }
    if (row_ok) {
      out[pl.c * pl.c + row] = dq;
      out[pl.c * pl.c + pl.c + row] = dk;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, pl.tmem_cols);
  }
}

// Generic Gram for heads wider than 128 channels: grid (32x32 tiles of the c x c Gram, pixel chunk, B * heads), 256 threads.
// Same partial layout as the tcgen05 kernel ([c*c Gram | c sum q^2 | c sum k^2] per chunk), fp32 FMA over 16-bit operands.
__global__ void __launch_bounds__(256) mdta_gram_generic_kernel(const uint16_t* __restrict__ qkv, long long ld, int C,
                                                                int heads, long long P, int nchunks,
                                                                float* __restrict__ partials, int fp16) {
  const int c = C / heads;
  const int nt = (c + 31) >> 5;
  const int ti = blockIdx.x / nt, tj = blockIdx.x % nt;
  const int chunk = blockIdx.y;
  const int b = blockIdx.z / heads, h = blockIdx.z % heads;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;            // thread -> column tx, rows ty*4 .. ty*4+3
  __shared__ float sq[32][33], sk[32][33];
  const long long p0 = (long long)chunk * kGenChunkPix;
  const long long p1 = p0 + kGenChunkPix < P ? p0 + kGenChunkPix : P;
  const uint16_t* base = qkv + (size_t)b * P * ld;
  const int qi = h * c + ti * 32, kj = C + h * c + tj * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float nq = 0.f, nk = 0.f;                                          // ty == 0 threads: sum of squares of column tx
  for (long long pp = p0; pp < p1; pp += 32) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {                                     // 32 pixels x 32 channels of q and of k
      const int pr = ty * 4 + r;
      const long long pix = pp + pr;
      float vq = 0.f, vk = 0.f;
      if (pix < p1) {
        float d;
        if (ti * 32 + tx < c) unpack2r(base[pix * ld + qi + tx], vq, d, fp16);
        if (tj * 32 + tx < c) unpack2r(base[pix * ld + kj + tx], vk, d, fp16);
      }
      sq[pr][tx] = vq;
      sk[pr][tx] = vk;
    }
    __syncthreads();
#pragma unroll 8
    for (int pr = 0; pr < 32; ++pr) {
      const float kv = sk[pr][tx];
#pragma unroll
      for (int r = 0; r < 4; ++r) acc[r] = fmaf(sq[pr][ty * 4 + r], kv, acc[r]);
    }
    if (ty == 0) {
#pragma unroll 8
      for (int pr = 0; pr < 32; ++pr) {
        nq = fmaf(sq[pr][tx], sq[pr][tx], nq);
        nk = fmaf(sk[pr][tx], sk[pr][tx], nk);
      }
    }
    __syncthreads();
  }
  float* out = partials + ((size_t)(b * heads + h) * nchunks + chunk) * ((size_t)c * c + 2 * c);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = ti * 32 + ty * 4 + r, j = tj * 32 + tx;
    if (i < c && j < c) out[(size_t)i * c + j] = acc[r];
  }
  if (ty == 0) {
    if (tj == 0 && ti * 32 + tx < c) out[(size_t)c * c + ti * 32 + tx] = nq;
    if (ti == 0 && tj * 32 + tx < c) out[(size_t)c * c + c + tj * 32 + tx] = nk;
  }
}

// Stage 1 of the finalize: grid (c, heads, B), one CTA of 8 warps per attention row i.  The per-chunk partial Grams are
// reduced in a FIXED two-level order (warp w sums chunks w, w+8, w+16, ... in order; the 8 warp sums are then added in warp
// order), so the result is deterministic and a function of (P, heads) only; the 8 warps keep 8 x 4 independent L2 loads
// in flight instead of one warp walking all ~148 chunks.  Then the F.normalize denominators, temperature and the softmax.
template <int T>   // T * 32 >= c columns per attention row: 4 (c <= 128) or 8 (wide heads)
__global__ void __launch_bounds__(256) mdta_softmax_kernel(const float* __restrict__ partials, int C, int heads,
                                                           int nchunks, const float* __restrict__ temperature,
                                                           float* __restrict__ attn, float* __restrict__ shat_out,
                                                           const float* __restrict__ topk_w) {
  const int c = C / heads;
  const int h = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x;
  const size_t psz = (size_t)c * c + 2 * c;
  const float* base = partials + (size_t)(b * heads + h) * nchunks * psz;
  __shared__ float sg[8][T][32], sk[8][T][32], sq[8];
  pdl_wait();
  pdl_launch();
  float g[T], nk[T];
#pragma unroll
  for (int t = 0; t < T; ++t) g[t] = nk[t] = 0.f;
  float nq = 0.f;
  for (int ch0 = warp; ch0 < nchunks; ch0 += 32) {          // 4 chunks of this warp's residue class at a time
    float tg[4][T], tk[4][T], tq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int ch = ch0 + 8 * u;
      const bool in = ch < nchunks;
      const float* p = base + (size_t)(in ? ch : 0) * psz;
#pragma unroll
      for (int t = 0; t < T; ++t) {
        const int j = lane + 32 * t;
        const bool ok = in && j < c;
        tg[u][t] = ok ? p[i * c + j] : 0.f;
        tk[u][t] = ok ? p[c * c + c + j] : 0.f;
      }
      tq[u] = in ? p[c * c + i] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int t = 0; t < T; ++t) {
        g[t] += tg[u][t];
        nk[t] += tk[u][t];
      }
      nq += tq[u];
    }
  }
#pragma unroll
  for (int t = 0; t < T; ++t) {
    sg[warp][t][lane] = g[t];
    sk[warp][t][lane] = nk[t];
  }
  if (lane == 0) sq[warp] = nq;
  __syncthreads();
  if (warp != 0) return;
  nq = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) g[t] = nk[t] = 0.f;
  for (int w = 0; w < 8; ++w) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      g[t] += sg[w][t][lane];
      nk[t] += sk[w][t][lane];
    }
    nq += sq[w];
  }
  const float nqi = fmaxf(sqrtf(fmaxf(nq, 0.f)), 1e-12f);
  const float temp = temperature[h];
  float mx = -INFINITY;
  // training: keep the normalised Gram and the norms ([c*c | nq(c) | nk(c)] per (sample, head)) for tdr_mdta_bwd
  float* so = shat_out ? shat_out + (size_t)(b * heads + h) * psz : nullptr;
  if (so && lane == 0) so[(size_t)c * c + i] = nqi;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int j = lane + 32 * t;
    if (j < c) {
      const float nkj = fmaxf(sqrtf(fmaxf(nk[t], 0.f)), 1e-12f);
      const float sh = g[t] / (nqi * nkj);
      if (so) {
        so[(size_t)i * c + j] = sh;
        if (i == 0) so[(size_t)c * c + c + j] = nkj;
      }
      g[t] = sh * temp;
      mx = fmaxf(mx, g[t]);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float* out = attn + ((size_t)(b * heads + h) * c + i) * c;
  if (topk_w) {
    // Top-k sparse attention (DRSformer TKSA, network_drsformer_guided_arch.py:296-327): four softmaxes over the
    // int(c/2), int(2c/3), int(3c/4), int(4c/5) largest entries of the row, mixed with the learnable attn1..attn4.
    int rank[T];
#pragma unroll
    for (int t = 0; t < T; ++t) rank[t] = 0;
    for (int tt = 0; tt < T; ++tt) {
      for (int src = 0; src < 32; ++src) {
        const int jj = src + 32 * tt;
        if (jj >= c) break;
        float mine = 0.f;
#pragma unroll
        for (int t = 0; t < T; ++t) mine = t == tt ? g[t] : mine;
        const float v = __shfl_sync(0xffffffffu, mine, src);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const int j = lane + 32 * t;
          rank[t] += (j < c && (v > g[t] || (v == g[t] && jj < j))) ? 1 : 0;
        }
      }
    }
    const int kk[4] = {(int)(c / 2.0), (int)(c * 2 / 3.0), (int)(c * 3 / 4.0), (int)(c * 4 / 5.0)};
    float e[T], s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int j = lane + 32 * t;
      e[t] = j < c ? __expf(g[t] - mx) : 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) s4[q] += (j < c && rank[t] < kk[q]) ? e[t] : 0.f;
    }
    float coef[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) coef[q] = topk_w[q] / warp_sum(s4[q]);
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const int j = lane + 32 * t;
      if (j < c) {
        float wsum = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) wsum += rank[t] < kk[q] ? coef[q] : 0.f;
        out[j] = e[t] * wsum;
      }
    }
    return;
  }
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int j = lane + 32 * t;
    if (j < c) {
      g[t] = __expf(g[t] - mx);
      sum += g[t];
    }
  }
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int j = lane + 32 * t;
    if (j < c) out[j] = g[t] * inv;
  }
}

// Stage 2: grid (ceil(C/kFoldRows), heads, B).  Weff[b][co][h*c + j] = sum_i W_out[co][h*c + i] * attn[b,h,i,j]
constexpr int kFoldRows = 8;      // output rows per CTA: small, so that even one (sample, head) spreads over C/8 CTAs
__global__ void __launch_bounds__(256) mdta_fold_kernel(const float* __restrict__ attn, int C, int heads,
                                                        const float* __restrict__ w_out, bf16* __restrict__ weff,
                                                        long long weff_ld, bf16* __restrict__ weff_t, int fp16) {
  extern __shared__ float sm[];
  const int c = C / heads;
  const int h = blockIdx.y, b = blockIdx.z;
  const int co0 = blockIdx.x * kFoldRows;
  float* a = sm;                   // [c][c]
  float* wsm = sm + c * c;         // [kFoldRows][c]
  const float* src = attn + (size_t)(b * heads + h) * c * c;
  // w_out is a parameter (never written by the softmax kernel that precedes this one in the stream): staged while
  // that kernel drains
  for (int t = threadIdx.x; t < kFoldRows * c; t += blockDim.x) {
    const int co = co0 + t / c;
    wsm[t] = co < C ? w_out[(size_t)co * C + h * c + t % c] : 0.f;
  }
  pdl_wait();
  pdl_launch();
  for (int t = threadIdx.x; t < (c * c) >> 2; t += blockDim.x)          // c % 8 == 0: float4 staging
    reinterpret_cast<float4*>(a)[t] = reinterpret_cast<const float4*>(src)[t];
  __syncthreads();
  for (int idx = threadIdx.x; idx < kFoldRows * c; idx += blockDim.x) {
    const int r = idx / c, j = idx % c;
    if (co0 + r >= C) break;
    const float* wr = wsm + r * c;
    float s = 0.f;
#pragma unroll 4
    for (int i = 0; i < c; ++i) s = fmaf(wr[i], a[i * c + j], s);
    reinterpret_cast<uint16_t*>(weff)[((size_t)b * C + co0 + r) * weff_ld + h * c + j] = pack1r(s, fp16);
    if (weff_t) weff_t[((size_t)b * C + h * c + j) * weff_ld + co0 + r] = __float2bfloat16(s);   // transposed (dgrad)
  }
}

}  // namespace

extern "C" size_t tdr_mdta_partials_bytes(int B, long long P, int C, int heads) {
  GramPlan p;
  if (B <= 0 || P <= 0 || make_plan(B, P, C, heads, &p)) return 0;
  return (size_t)B * heads * p.nchunks * ((size_t)p.c * p.c + 2 * p.c) * sizeof(float);
}

extern "C" int tdr_mdta_gram(const void* qkv_bf16, long long ld, int B, long long P, int C, int heads, float* partials,
                             int fp16, cudaStream_t stream) {
  TDR_CHECK_ARG(qkv_bf16 && partials && B > 0 && P > 0, "tdr_mdta_gram: bad arguments");
  TDR_CHECK_ARG(ld % 8 == 0 && ld >= 3 * C, "tdr_mdta_gram: ld must be a multiple of 8 and >= 3C");
  GramArgs a;
  TDR_CHECK_ARG(make_plan(B, P, C, heads, &a.plan) == 0,
                "tdr_mdta_gram: unsupported head width (C=%d heads=%d; need c%%8==0, c<=128, c%%16==0 if c>64)", C, heads);
  a.B = B; a.C = C; a.heads = heads; a.P = P; a.partials = partials; a.fp16 = fp16 ? 1 : 0;
  if (a.plan.generic) {
    const int nt = (a.plan.c + 31) / 32;
    TDR_CHECK_ARG((long long)B * heads <= 65535 && a.plan.nchunks <= 65535, "tdr_mdta_gram: grid too large");
    dim3 grid(nt * nt, a.plan.nchunks, B * heads);
    mdta_gram_generic_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(qkv_bf16), ld, C, heads, P,
                                                       a.plan.nchunks, partials, a.fp16);
    TDR_CHECK_LAUNCH();
    return TDR_OK;
  }
  TdrTensorMap map;
  const uint64_t dims[3] = {(uint64_t)(3 * C), (uint64_t)P, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)ld * 2 * (uint64_t)P};
  const uint32_t box[3] = {(uint32_t)a.plan.box_ch, (uint32_t)kPixTile, 1};
  const uint32_t es[3] = {1, 1, 1};
  int rc = a.plan.box_ch == 32 ? tdr_make_tensor_map_bf16_sw64(&map, qkv_bf16, 3, dims, strides, box, es)
                               : tdr_make_tensor_map_bf16(&map, qkv_bf16, 3, dims, strides, box, es);
  if (rc) return rc;
  // + one box: with three 32-channel boxes the M = 128 operand reads a fourth (unused) chunk past k's last box
  const size_t smem = 1024 + (size_t)a.plan.stages * 2 * a.plan.boxes * a.plan.box_bytes + 256 + a.plan.box_bytes;
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(mdta_gram_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid(a.plan.nchunks, heads, B);
  TDR_CHECK_CUDA(tdr_launch_pdl(mdta_gram_kernel, grid, dim3(192), smem, stream, map, a));
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}

extern "C" int tdr_mdta_weff(const float* partials, int B, long long P, int C, int heads, const float* temperature,
                             const float* w_out, void* weff_bf16, long long weff_ld, float* attn_ws,
                             void* weff_t_bf16, float* shat_out, int fp16, const float* topk_w, cudaStream_t stream) {
  TDR_CHECK_ARG(partials && temperature && w_out && weff_bf16 && attn_ws, "tdr_mdta_weff: null pointer");
  GramPlan p;
  TDR_CHECK_ARG(make_plan(B, P, C, heads, &p) == 0, "tdr_mdta_weff: unsupported head width");
  TDR_CHECK_ARG(weff_ld >= C && weff_ld % 8 == 0, "tdr_mdta_weff: bad weff_ld");
  {
    dim3 grid(p.c, heads, B);
    if (p.c <= 128)
      TDR_CHECK_CUDA(tdr_launch_pdl(mdta_softmax_kernel<4>, grid, dim3(256), 0, stream, partials, C, heads, p.nchunks,
                                    temperature, attn_ws, shat_out, topk_w));
    else
      TDR_CHECK_CUDA(tdr_launch_pdl(mdta_softmax_kernel<8>, grid, dim3(256), 0, stream, partials, C, heads, p.nchunks,
                                    temperature, attn_ws, shat_out, topk_w));
    TDR_CHECK_LAUNCH();
  }
  const size_t smem = ((size_t)p.c * p.c + kFoldRows * p.c) * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(mdta_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  dim3 grid((C + kFoldRows - 1) / kFoldRows, heads, B);
  TDR_CHECK_CUDA(tdr_launch_pdl(mdta_fold_kernel, grid, dim3(256), smem, stream, attn_ws, C, heads, w_out,
                                reinterpret_cast<bf16*>(weff_bf16), weff_ld, reinterpret_cast<bf16*>(weff_t_bf16), fp16));
  TDR_CHECK_LAUNCH();
  return TDR_OK;
}
