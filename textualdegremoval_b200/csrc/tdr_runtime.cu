// Host-side runtime glue of libtdr_sm100.so: thread-local error string, device check, and CUtensorMap
// construction through the driver entry point (resolved at run time so the library has no link-time
// dependency on libcuda and can be built on a box without a GPU).
#include <cuda.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include "tdr_common.cuh"

static thread_local char g_err[512] = "";

void tdr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* tdr_last_error(void) { return g_err; }
extern "C" int tdr_version(void) { return 100; }

int tdr_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

static int g_pdl = -1;
int tdr_pdl_enabled() {
  if (g_pdl < 0) {
    const char* e = getenv("TDR_PDL");
    g_pdl = (e && e[0] == '1') ? 1 : 0;      // opt-in: measured neutral on the forward step, see DESIGN.md
  }
  return g_pdl;
}
extern "C" int tdr_set_pdl(int on) {
  const int prev = tdr_pdl_enabled();
  g_pdl = on ? 1 : 0;
  return prev;
}

extern "C" int tdr_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  TDR_CHECK_CUDA(cudaGetDevice(&dev));
  TDR_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TDR_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    tdr_set_error("libtdr_sm100 needs an sm_100a device (B200); found sm_%d%d", major, minor);
    return TDR_ENOSUP;
  }
  return TDR_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

static int make_map(TdrTensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B);

int tdr_make_tensor_map_bf16_noswizzle(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                                       const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_map(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, elem_strides,
                  CU_TENSOR_MAP_SWIZZLE_NONE);
}

int tdr_make_tensor_map_bf16_sw64(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                                  const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_map(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, elem_strides,
                  CU_TENSOR_MAP_SWIZZLE_64B);
}

int tdr_make_tensor_map_bf16(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                             const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_map(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, elem_strides);
}
int tdr_make_tensor_map_f32(TdrTensorMap* out, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides) {
  return make_map(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, elem_strides);
}

static int make_map(TdrTensorMap* out, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                    CUtensorMapSwizzle sw) {
  static_assert(sizeof(CUtensorMap) == sizeof(TdrTensorMap), "tensor map size");
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    tdr_set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return TDR_ECUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = elem_strides[i];
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapL2promotion promo =
      sw == CU_TENSOR_MAP_SWIZZLE_NONE ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  static const char* promo_env = getenv("TDR_TMA_L2PROMO");                    // tuning knob (experiments only)
  if (promo_env && sw != CU_TENSOR_MAP_SWIZZLE_NONE) {
    const int v = atoi(promo_env);
    promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                   : (v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                              : (v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B));
  }
  CUresult r = enc(reinterpret_cast<CUtensorMap*>(out), dt, (cuuint32_t)rank,
                   const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    tdr_set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu %llu %llu %llu] stride0 %llu box [%u %u %u %u]",
                  (int)r, rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                  (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
                  (unsigned long long)(rank > 1 ? strides_bytes[0] : 0), box[0], rank > 1 ? box[1] : 0,
                  rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return TDR_ECUDA;
  }
  return TDR_OK;
}

extern "C" void tdr_conv_gemm_desc_layout(int* out) {
  out[0] = (int)sizeof(tdr_conv_gemm_desc);
  out[1] = (int)offsetof(tdr_conv_gemm_desc, weight);
  out[2] = (int)offsetof(tdr_conv_gemm_desc, origin);
  out[3] = (int)offsetof(tdr_conv_gemm_desc, bias);
  out[4] = (int)offsetof(tdr_conv_gemm_desc, scale_ptr);
  out[5] = (int)offsetof(tdr_conv_gemm_desc, res1);
  out[6] = (int)offsetof(tdr_conv_gemm_desc, res2);
  out[7] = (int)offsetof(tdr_conv_gemm_desc, out_f32);
  out[8] = (int)offsetof(tdr_conv_gemm_desc, out_bf16);
  out[9] = (int)offsetof(tdr_conv_gemm_desc, impl);
  out[10] = (int)offsetof(tdr_conv_gemm_desc, ln_mode);
  out[11] = (int)offsetof(tdr_conv_gemm_desc, ln_weight);
  out[12] = (int)offsetof(tdr_conv_gemm_desc, ln_out_bf16);
}
