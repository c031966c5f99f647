// Common device/host helpers for the sm_100a kernels: error reporting, bf16 packing, and thin
// inline-PTX wrappers for mbarrier / TMA (cp.async.bulk.tensor) / tcgen05 (UMMA + TMEM).
// Everything here is written for sm_100a only; there is no other code path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/tdr_sm100.h"

// ------------------------------------------------------------------------------------------------
// host-side error plumbing (C-ABI: every entry point returns 0 or a negative TDR_E* code)
// ------------------------------------------------------------------------------------------------
void tdr_set_error(const char* fmt, ...);

#define TDR_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      tdr_set_error(__VA_ARGS__);                \
      return TDR_EINVAL;                         \
    }                                            \
  } while (0)

#define TDR_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      tdr_set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, cudaGetErrorName(_e),  \
                    cudaGetErrorString(_e));                                              \
      return TDR_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define TDR_CHECK_LAUNCH() TDR_CHECK_CUDA(cudaGetLastError())

static inline int tdr_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
// (m, s) with n / d == __umulhi(n, m) >> s for every 0 <= n < 2^31; m == 0 encodes d == 1 (quotient = n)
static inline void tdr_fast_div_setup(int d, uint32_t* m, uint32_t* s) {
  if (d <= 1) { *m = 0; *s = 0; return; }
  uint32_t l = 0;
  while ((1u << l) < (uint32_t)d) ++l;                       // ceil(log2 d)
  *m = (uint32_t)((((uint64_t)1 << (31 + l)) + (uint64_t)d - 1) / (uint64_t)d);
  *s = l - 1;
}

int tdr_num_sms();
// Programmatic dependent launch: 1 lets the hot kernels' prologues (barrier init, TMEM allocation, tensor-map prefetch,
// CTA launch) overlap the tail of the previous kernel in the stream.  Off by default (TDR_PDL=1 / tdr_set_pdl(1) turn
// it on): measured neutral on the forward step and 2 % slower on the training step (profiles/r02b_pdl_ab.txt).
int tdr_pdl_enabled();

// ------------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

struct __align__(16) bf16x8 {
  bf162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float* f) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float* f) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  bf162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// IEEE fp16 pair, round to nearest, saturating at +-65504 (one F2FP.SATFINITE): fp16 operands must never become inf
__device__ __forceinline__ uint32_t pack2h(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
// 16-bit operand format selected at compile time (H = IEEE fp16, else bf16) or at run time
template <bool H>
__device__ __forceinline__ uint32_t pack2t(float a, float b) {
  if constexpr (H) return pack2h(a, b);
  else return pack2(a, b);
}
__device__ __forceinline__ uint32_t pack2r(float a, float b, int fp16) { return fp16 ? pack2h(a, b) : pack2(a, b); }
template <bool H>
__device__ __forceinline__ void unpack2t(uint32_t u, float& a, float& b) {
  if constexpr (H) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u));
    a = f.x; b = f.y;
  } else {
    a = __uint_as_float(u << 16); b = __uint_as_float(u & 0xffff0000u);
  }
}
__device__ __forceinline__ void unpack2r(uint32_t u, float& a, float& b, int fp16) {
  if (fp16) unpack2t<true>(u, a, b);
  else unpack2t<false>(u, a, b);
}

// 8 consecutive 16-bit values (bf16, or IEEE fp16 when `fp16`) <-> fp32, through one 128-bit access
__device__ __forceinline__ void unpack8r(const void* p, float* f, int fp16) {
  const uint4 v = *reinterpret_cast<const uint4*>(p);
  unpack2r(v.x, f[0], f[1], fp16); unpack2r(v.y, f[2], f[3], fp16);
  unpack2r(v.z, f[4], f[5], fp16); unpack2r(v.w, f[6], f[7], fp16);
}
__device__ __forceinline__ void pack8r(void* p, const float* f, int fp16) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack2r(f[0], f[1], fp16), pack2r(f[2], f[3], fp16), pack2r(f[4], f[5], fp16),
                                            pack2r(f[6], f[7], fp16));
}
__device__ __forceinline__ uint16_t pack1r(float a, int fp16) { return (uint16_t)(pack2r(a, 0.f, fp16) & 0xffffu); }

// Programmatic dependent launch.  Contract for every kernel launched through tdr_launch_pdl: pdl_wait() comes before the
// first access to global memory that is not a launch parameter (it returns once the previous kernel in the stream has
// finished and its writes are visible), and pdl_launch() comes after it and after every dynamic resource (TMEM) is
// held, so that the next kernel's CTAs may be placed as soon as this kernel's CTAs free their SMs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// explicit shared-space 128-bit accesses (a generic-pointer store of a small struct may be split into 32-bit ST.E)
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
  return r;
}

__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
  uint32_t r;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(r) : "r"(saddr) : "memory");
  return r;
}
__device__ __forceinline__ void sts32(uint32_t saddr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) {
      printf("tdr: mbarrier timeout block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// Same, for roles that usually wait long (a whole tile): back off between polls so that the spinning warp does not take
// issue slots from the working warps of its scheduler.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity, unsigned ns = 200) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(ns);
    if (++spins > (1u << 24)) {
      printf("tdr: mbarrier timeout (sleeping wait) block %d thread %d parity %u\n", blockIdx.x, threadIdx.x, parity);
      __trap();
    }
  }
}

// ---- TMA ---------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const void* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// L2 prefetch of a tensor-map box (no shared memory involved): lets a deep software prefetch distance hide DRAM latency
// when the smem ring itself cannot hold enough bytes in flight
__device__ __forceinline__ void tma_prefetch_4d(const void* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1),
               "r"(c2), "r"(c3)
               : "memory");
}

// TMA store (shared -> global), bulk async-group completion
__device__ __forceinline__ void tma_store_4d(const void* map, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read3() { asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (TMA) before a bulk store reads them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 / TMEM ----------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// One lane of the (converged) warp, chosen by the hardware.  Unlike `lane == 0` the compiler knows that exactly one thread
// runs the guarded region, so register operands of tcgen05.mma / TMA instructions move to uniform registers with a plain
// R2UR instead of an ELECT / R2UR.BROADCAST / BRA.U.ANY waterfall loop per instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; one thread issues for the CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 16 consecutive fp32 columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp): SWIZZLE_128B, version 1.
//   K-major operand : rows of 128 B (64 bf16 of K); 8-row groups 1024 B apart  -> SBO = 1024, LBO unused.
//   MN-major operand: rows of 128 B (64 bf16 of M/N) indexed by K; 8 K-rows = 1024 B atom -> SBO = 1024
//                     between K-groups, LBO = byte stride between 64-element M/N chunks.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
// Same with SWIZZLE_64B (layout type 4): rows of 64 B (32 bf16); an 8-row group is 512 B.
__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)4 << 61;   // SWIZZLE_64B
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32 (cute::UMMA::InstrDescriptor).
// a_fp16 / b_fp16: that operand holds IEEE fp16 instead of bf16 (kind::f16 takes either, per operand, at the same rate)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major, int a_fp16 = 0,
                                                       int b_fp16 = 0) {
  return (1u << 4)                          // D = f32
         | ((a_fp16 ? 0u : 1u) << 7)        // A = bf16 (1) / fp16 (0)
         | ((b_fp16 ? 0u : 1u) << 10)       // B = bf16 (1) / fp16 (0)
         | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
#endif  // __CUDACC__

#ifdef __CUDACC__
// <<<grid, block, smem, stream>>> with the programmatic-stream-serialization attribute (see pdl_wait above)
template <typename... KArgs, typename... Args>
static inline cudaError_t tdr_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = tdr_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// host: build a CUtensorMap (bf16, SWIZZLE_128B, zero OOB fill).  dims/strides innermost first;
// strides_bytes has rank-1 entries (stride of dim 1..rank-1).  Returns 0 or TDR_E*.
struct TdrTensorMap {
  alignas(64) unsigned char bytes[128];
};
int tdr_make_tensor_map_bf16(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);
// bf16, SWIZZLE_64B (box inner dimension <= 32 elements = 64 B)
int tdr_make_tensor_map_bf16_sw64(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                                  const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);
// bf16, SWIZZLE_NONE (plain row-major box in shared memory)
int tdr_make_tensor_map_bf16_noswizzle(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                                       const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);
// same for fp32 data (box inner dim <= 32 elements = 128 B)
int tdr_make_tensor_map_f32(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides);
