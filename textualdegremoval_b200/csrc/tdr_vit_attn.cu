// Fused multi-head self-attention of the frozen ViT encoders (forward only):
//     out[b, i, h*hd : (h+1)*hd] = softmax_j(scale * q_i . k_j) v_j
// for DINOv2 ViT-B/14 (/root/reference/models/dino/attention.py:52-68, hd = 64, N = 1370 at 518 x 518) and the CLIP
// ViT-H/14 vision tower (transformers CLIPAttention called from scripts/train/main_train_tr_mapping.py:780, hd = 80,
// N = 257).  It replaces the materialised-score schedule (2 x heads tdr_conv_gemm launches + tdr_softmax_rows +
// tdr_vit_transpose_v per layer, an fp32 [heads, B, N, N] tensor through HBM) with ONE launch per layer:
//
//   grid (ceil(N/128), heads, B), 128 threads.  The CTA owns 128 query rows of one (sample, head) and walks the keys in
//   tiles of 128:  TMA stages Q once and K_j / V_j per tile straight from the packed qkv rows [B, N, 3D] (SWIZZLE_128B;
//   rows past N are the TMA zero fill), S = Q K_j^T is one tcgen05.mma chain into TMEM (128 x 128 fp32), thread r reads
//   row r of S (tcgen05.ld), keeps the running max / sum of the online softmax in registers, writes P = exp2(c (s - m))
//   as the bf16 K-major A operand into shared memory, P V_j is a second tcgen05.mma chain (V is consumed MN-major, as it
//   lies in memory -- no transposed copy) into its own TMEM columns, and the thread folds it into its fp32 output row
//   o = o * alpha + PV.  S_{j+1} is issued right behind P V_j so that it runs under the accumulate of tile j; two CTAs per
//   SM (hd = 64) cover the remaining bubbles.  HBM traffic is the qkv rows (K, V re-read from L2 per query tile) and out.
#include "tdr_common.cuh"

namespace {

constexpr int kQT = 128;                 // query rows per CTA
constexpr int kKT = 128;                 // keys per tile
constexpr int kBox = kKT * 128;          // one TMA box: 128 rows x 64 bf16 = 16 KB

struct VitAttnArgs {
  int B, N, heads, hd, D;
  float c;                                // scale * log2(e)
  bf16* out;
  long long out_ld;
};

__device__ __forceinline__ float ex2f(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Online-softmax update of one thread's query row with one tile of scores (TMEM row at `trow`): running max / sum, the
// rescale factor alpha of what was accumulated so far, and P = exp2(c s - c m) as the bf16 K-major A operand in shared
// memory (SWIZZLE_128B: 16-byte chunk q of row r sits at q ^ (r & 7)).  MASK: keys >= valid do not exist (last tile).
template <bool MASK>
__device__ __forceinline__ void softmax_tile(uint32_t trow, int valid, float c, float& m_run, float& l_run, float& alpha,
                                             uint32_t p_row, int sw) {
  // pass 1: row maximum over the keys of the tile
  float mt = -INFINITY;
  for (int c0 = 0; c0 < valid; c0 += 32) {
    uint32_t s[32];
    tmem_ld16(trow + c0, s);
    tmem_ld16(trow + c0 + 16, s + 16);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (!MASK || c0 + i < valid) mt = fmaxf(mt, __uint_as_float(s[i]));
  }
  const float m_new = fmaxf(m_run, mt);
  alpha = ex2f((m_run - m_new) * c);                       // first tile: exp2(-inf) = 0
  m_run = m_new;
  const float mc = m_new * c;
  float lsum = 0.f;
  // pass 2: probabilities
  const int vend = (valid + 15) & ~15;
  for (int c0 = 0; c0 < vend; c0 += 32) {
    uint32_t s[32];
    tmem_ld16(trow + c0, s);
    tmem_ld16(trow + c0 + 16, s + 16);
    tmem_ld_wait();
    uint32_t pk[16];
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const float p0 = (!MASK || c0 + i < valid) ? ex2f(fmaf(__uint_as_float(s[i]), c, -mc)) : 0.f;
      const float p1 = (!MASK || c0 + i + 1 < valid) ? ex2f(fmaf(__uint_as_float(s[i + 1]), c, -mc)) : 0.f;
      lsum += p0 + p1;
      pk[i >> 1] = pack2(p0, p1);
    }
    const uint32_t base = p_row + (c0 >> 6) * kBox;
    const int q8 = (c0 & 63) >> 3;                          // first 16-byte chunk of this 32-key group within the row
#pragma unroll
    for (int g = 0; g < 4; ++g)
      sts128(base + (((q8 + g) ^ sw) << 4), pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
  }
  l_run = l_run * alpha + lsum;
}

template <int HD>
__global__ void __launch_bounds__(128, HD <= 64 ? 2 : 1) vit_attn_kernel(const __grid_constant__ TdrTensorMap map,
                                                                         const VitAttnArgs a) {
  constexpr int NCH = (HD + 63) / 64;                      // 64-wide column chunks of a head
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + NCH * kBox;
  uint8_t* sV = sK + NCH * kBox;
  uint8_t* sP = sV + NCH * kBox;                           // 2 chunks of 64 keys
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kBox);
  uint64_t *bar_q = bars, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_o = bars + 4;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5;
  const int q0 = blockIdx.x * kQT, h = blockIdx.y, b = blockIdx.z;
  const int nkv = (a.N + kKT - 1) / kKT;

  if (tid == 0) {
    tma_prefetch_desc(&map);
    for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(tmem_ptr, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tS = tmem, tO = tmem + 128;

  auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col0, int row0) {      // thread 0
    mbar_expect_tx(bar, NCH * kBox);
#pragma unroll
    for (int c = 0; c < NCH; ++c) tma_load_3d(dst + c * kBox, &map, bar, col0 + 64 * c, row0, b);
  };
  auto issue_s = [&](int j) {                                                  // thread 0: S = Q K_j^T
    int n = a.N - j * kKT;
    n = n >= kKT ? kKT : (n + 15) & ~15;
    const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      const uint64_t da = umma_desc_sw128(smem_u32(sQ + (ks >> 2) * kBox), 0, 1024) + 2 * (ks & 3);
      const uint64_t db = umma_desc_sw128(smem_u32(sK + (ks >> 2) * kBox), 0, 1024) + 2 * (ks & 3);
      umma_bf16(tS, da, db, idesc, ks != 0);
    }
    umma_commit(bar_s);
  };

  const bool leader = warp == 0 && elect_one();           // one lane of warp 0 issues every TMA load and MMA
  if (leader) {
    load_tile(sQ, bar_q, h * HD, q0);
    load_tile(sK, bar_k, a.D + h * HD, 0);
    load_tile(sV, bar_v, 2 * a.D + h * HD, 0);
    mbar_wait(bar_q, 0);
    mbar_wait(bar_k, 0);
    tc_fence_after();
    issue_s(0);
  }
  __syncwarp();

  const bool active = q0 + warp * 32 < a.N;                // warp-uniform: any valid query row in this warp?
  const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
  const int r = tid;                                        // query row of this thread within the tile
  const uint32_t p_row = smem_u32(sP) + r * 128;
  const int sw = r & 7;

  float o[HD];
#pragma unroll
  for (int i = 0; i < HD; ++i) o[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < nkv; ++j) {
    const int valid = min(kKT, a.N - j * kKT);
    mbar_wait(bar_s, j & 1);
    tc_fence_after();
    if (leader && j + 1 < nkv) load_tile(sK, bar_k, a.D + h * HD, (j + 1) * kKT);   // S_j is done with K_j
    float alpha = 1.f;
    if (active) {
      // full tiles (all but the last) skip the key mask: the compare / select pairs are a third of the softmax issue slots
      if (valid == kKT) softmax_tile<false>(tS + lane_off, valid, a.c, m_run, l_run, alpha, p_row, sw);
      else softmax_tile<true>(tS + lane_off, valid, a.c, m_run, l_run, alpha, p_row, sw);
    }
    fence_proxy_async();                                    // P (generic proxy) -> tcgen05.mma (async proxy)
    tc_fence_before();
    __syncthreads();                                        // P complete, every thread is done reading S_j
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      mbar_wait(bar_v, j & 1);
      const int ksteps = (valid + 15) >> 4;
      const uint32_t idesc = umma_idesc_bf16(128, HD, 0, 1);
      const uint64_t dv0 = umma_desc_sw128(smem_u32(sV), kBox, 1024);
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t da = umma_desc_sw128(smem_u32(sP + (ks >> 2) * kBox), 0, 1024) + 2 * (ks & 3);
        umma_bf16(tO, da, dv0 + 128 * ks, idesc, ks != 0);
      }
      umma_commit(bar_o);
      if (j + 1 < nkv) {                                    // S_{j+1} runs under the accumulate of tile j
        mbar_wait(bar_k, (j + 1) & 1);
        issue_s(j + 1);
      }
    }
    __syncwarp();
    mbar_wait(bar_o, j & 1);
    tc_fence_after();
    if (leader && j + 1 < nkv) load_tile(sV, bar_v, 2 * a.D + h * HD, (j + 1) * kKT);   // P V_j is done with V_j
    if (active) {
#pragma unroll
      for (int c0 = 0; c0 < HD; c0 += 16) {
        uint32_t pv[16];
        tmem_ld16(tO + lane_off + c0, pv);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[c0 + i] = fmaf(o[c0 + i], alpha, __uint_as_float(pv[i]));
      }
    }
    tc_fence_before();                                      // the next P V (after the next __syncthreads) overwrites tO
  }

  if (active && q0 + r < a.N) {
    const float inv = 1.f / l_run;
    bf16* dst = a.out + ((long long)b * a.N + q0 + r) * a.out_ld + h * HD;
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += 8) {
      uint4 v;
      v.x = pack2(o[c0] * inv, o[c0 + 1] * inv);
      v.y = pack2(o[c0 + 2] * inv, o[c0 + 3] * inv);
      v.z = pack2(o[c0 + 4] * inv, o[c0 + 5] * inv);
      v.w = pack2(o[c0 + 6] * inv, o[c0 + 7] * inv);
      *reinterpret_cast<uint4*>(dst + c0) = v;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

template <int HD>
int launch(const TdrTensorMap& map, const VitAttnArgs& a, cudaStream_t stream) {
  constexpr int NCH = (HD + 63) / 64;
  const size_t smem = (size_t)(3 * NCH + 2) * kBox + 64 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    TDR_CHECK_CUDA(cudaFuncSetAttribute(vit_attn_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const dim3 grid((unsigned)tdr_cdiv(a.N, kQT), (unsigned)a.heads, (unsigned)a.B);
  vit_attn_kernel<HD><<<grid, 128, smem, stream>>>(map, a);
  TDR_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int tdr_vit_attention_supported(int hd) { return hd == 16 || hd == 32 || hd == 64 || hd == 80; }

// qkv: bf16 rows [B, N, ld] with q | k | v at columns 0 | D | 2D (D = heads * hd, head h at h * hd inside each);
// out: bf16 rows [B, N, out_ld], head h written at columns [h * hd, (h + 1) * hd).
extern "C" int tdr_vit_attention(const void* qkv_bf16, long long ld, int B, int N, int heads, int hd, float scale,
                                 void* out_bf16, long long out_ld, cudaStream_t stream) {
  TDR_CHECK_ARG(qkv_bf16 && out_bf16 && B > 0 && N > 0 && heads > 0, "tdr_vit_attention: bad arguments");
  TDR_CHECK_ARG(tdr_vit_attention_supported(hd), "tdr_vit_attention: head dim %d (16, 32, 64 and 80 are built)", hd);
  const int D = heads * hd;
  TDR_CHECK_ARG(ld >= 3LL * D && ld % 8 == 0 && out_ld >= D && out_ld % 8 == 0 && B <= 65535 && heads <= 65535,
                "tdr_vit_attention: row strides %lld / %lld for D = %d", ld, out_ld, D);
  TDR_CHECK_ARG(((uintptr_t)qkv_bf16 & 15) == 0 && ((uintptr_t)out_bf16 & 15) == 0, "tdr_vit_attention: 16-byte alignment");
  TdrTensorMap map;
  const uint64_t dims[3] = {(uint64_t)(3 * D), (uint64_t)N, (uint64_t)B};
  const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)N * (uint64_t)ld * 2};
  const uint32_t box[3] = {64, (uint32_t)kKT, 1}, es[3] = {1, 1, 1};
  int rc = tdr_make_tensor_map_bf16(&map, qkv_bf16, 3, dims, strides, box, es);
  if (rc) return rc;
  VitAttnArgs a;
  a.B = B; a.N = N; a.heads = heads; a.hd = hd; a.D = D;
  a.c = scale * 1.4426950408889634f;
  a.out = reinterpret_cast<bf16*>(out_bf16);
  a.out_ld = out_ld;
  cudaStream_t st = stream;
  switch (hd) {
    case 16: return launch<16>(map, a, st);
    case 32: return launch<32>(map, a, st);
    case 64: return launch<64>(map, a, st);
    default: return launch<80>(map, a, st);
  }
}
