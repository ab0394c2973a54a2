#include "common.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

namespace cm3p {

static thread_local char g_err[1024] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
const char* last_error() { return g_err; }

static int g_num_sms = -1;
static int g_cc = -1;

static void query_device() {
  if (g_num_sms >= 0) return;
  int dev = 0, sms = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    g_num_sms = 0;
    g_cc = 0;
    return;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  g_num_sms = sms;
  g_cc = major * 10 + minor;
}

int num_sms() {
  query_device();
  return g_num_sms;
}

int check_arch() {
  query_device();
  if (g_cc != 100 && g_cc != 103)
    return set_error(kUnsupportedArch,
                     "cm3p_b200 kernels are built for sm_100a only; current device reports sm_%d (no fallback path)",
                     g_cc);
  return kOk;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  CM3P_REQUIRE(fn != nullptr, kDriverError, "cuTensorMapEncodeTiled entry point not available");
  CM3P_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, kBadAlignment, "TMA base %p not 16-byte aligned", base);
  CM3P_REQUIRE((pitch_bytes & 15) == 0, kBadAlignment, "TMA row pitch %llu B not a multiple of 16",
               (unsigned long long)pitch_bytes);
  CM3P_REQUIRE(box_inner * 2 <= 128 && box_outer <= 256, kBadShape, "TMA box %ux%u too large", box_inner, box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CM3P_REQUIRE(r == CUDA_SUCCESS, kDriverError,
               "cuTensorMapEncodeTiled(2d) failed with %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)", (int)r,
               (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch_bytes, box_inner,
               box_outer);
  return kOk;
}

int encode_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t mid, uint64_t outer,
                        uint64_t pitch_mid_bytes, uint64_t pitch_outer_bytes, uint32_t box_inner, uint32_t box_mid,
                        uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  CM3P_REQUIRE(fn != nullptr, kDriverError, "cuTensorMapEncodeTiled entry point not available");
  CM3P_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, kBadAlignment, "TMA base %p not 16-byte aligned", base);
  CM3P_REQUIRE((pitch_mid_bytes & 15) == 0 && (pitch_outer_bytes & 15) == 0, kBadAlignment,
               "TMA pitches %llu/%llu B not multiples of 16", (unsigned long long)pitch_mid_bytes,
               (unsigned long long)pitch_outer_bytes);
  cuuint64_t dims[3] = {inner, mid, outer};
  cuuint64_t strides[2] = {pitch_mid_bytes, pitch_outer_bytes};
  cuuint32_t box[3] = {box_inner, box_mid, box_outer};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CM3P_REQUIRE(r == CUDA_SUCCESS, kDriverError, "cuTensorMapEncodeTiled(3d) failed with %d", (int)r);
  return kOk;
}

}  // namespace cm3p
