// Packed fp32x2 helpers (FFMA2 on sm_100) and the GELU evaluations shared by the depthwise-stencil kernels
// (tdr_pointwise.cu: tdr_dwconv3x3*, tdr_gdfn.cu: the fused GDFN tail).
#pragma once
#include "tdr_common.cuh"

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
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// two bf16 packed in a 32-bit word -> fp32 pair (exact)
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ f2 bf2_to_f2(uint32_t u) { return pk2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }
// two 16-bit values (bf16, or IEEE fp16 when HALF) packed in a 32-bit word -> fp32 pair (exact)
template <bool HALF>
__device__ __forceinline__ f2 x2_to_f2(uint32_t u) {
  float a, b;
  unpack2t<HALF>(u, a, b);
  return pk2(a, b);
}

__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// Exact GELU x * Phi(x) (F.gelu default, R:239) on a pair.  Phi(x) = 1 / (1 + exp(-p(x))) with p the odd degree-13
// minimax fit of logit(Phi) on |x| <= 7 (|gelu error| <= 7e-8 in exact arithmetic, ~1e-6 = 2 fp32 ulp of x with
// MUFU.EX2 / MUFU.RCP; the fit was made in tools/gelu_fit.py).  Being odd it needs no |x| / sign select, so a pair costs
// 10 packed ops + 4 MUFU; beyond the fit range p keeps growing, so the result saturates to x (x > 0) or -0 (x < 0).
// The coefficients below are -log2(e) * c_k so that ex2(q) = exp(-p).
__device__ __forceinline__ f2 gelu2(f2 x) {
#define TDR_C2(v) pk2(v, v)
  const f2 x2 = mul2(x, x);
  f2 q = fma2(x2, TDR_C2(-5.2338826606046496e-09f), TDR_C2(3.8601649521297077e-07f));
  q = fma2(q, x2, TDR_C2(-1.1466740943433251e-05f));
  q = fma2(q, x2, TDR_C2(0.000159485251060687f));
  q = fma2(q, x2, TDR_C2(9.529500675853342e-05f));
  q = fma2(q, x2, TDR_C2(-0.10483819246292114f));
  q = fma2(q, x2, TDR_C2(-2.3022074699401855f));
  float q0, q1;
  upk2(mul2(q, x), q0, q1);
  float d0, d1;
  upk2(add2(pk2(ex2_approx(q0), ex2_approx(q1)), TDR_C2(1.f)), d0, d1);   // 1 + exp(-p); +inf -> rcp gives 0
#undef TDR_C2
  return mul2(x, pk2(rcp_approx(d0), rcp_approx(d1)));
}

// GELU and its derivative on a pair, sharing one erf evaluation: with z = |x|/sqrt2, e = exp(-z^2) = exp(-x^2/2) is both
// the A&S tail factor and (times 1/sqrt(2 pi)) the normal pdf, so gelu'(x) = Phi(x) + x pdf(x) costs two more FMAs.
__device__ __forceinline__ void gelu2_grad(f2 x, f2& gelu, f2& dgelu) {
  float x0, x1;
  upk2(x, x0, x1);
  const float z0 = fabsf(x0) * 0.70710678118654752f, z1 = fabsf(x1) * 0.70710678118654752f;
  const f2 z = pk2(z0, z1);
  float d0, d1;
  upk2(fma2(pk2(0.3275911f, 0.3275911f), z, pk2(1.f, 1.f)), d0, d1);
  const f2 t = pk2(rcp_approx(d0), rcp_approx(d1));
  f2 p = fma2(t, pk2(1.061405429f, 1.061405429f), pk2(-1.453152027f, -1.453152027f));
  p = fma2(p, t, pk2(1.421413741f, 1.421413741f));
  p = fma2(p, t, pk2(-0.284496736f, -0.284496736f));
  p = fma2(p, t, pk2(0.254829592f, 0.254829592f));
  p = mul2(p, t);
  float e0, e1;
  upk2(mul2(z, mul2(z, pk2(-1.4426950408889634f, -1.4426950408889634f))), e0, e1);
  const f2 ex = pk2(ex2_approx(e0), ex2_approx(e1));                 // exp(-x^2 / 2)
  float q0, q1;
  upk2(mul2(p, ex), q0, q1);                                          // q = 1 - erf(|x|/sqrt2)
  const f2 cdf = pk2(x0 >= 0.f ? 1.f - 0.5f * q0 : 0.5f * q0, x1 >= 0.f ? 1.f - 0.5f * q1 : 0.5f * q1);
  gelu = mul2(x, cdf);
  dgelu = fma2(mul2(x, pk2(0.3989422804014327f, 0.3989422804014327f)), ex, cdf);
}

