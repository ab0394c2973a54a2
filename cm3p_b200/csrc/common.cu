#include "common.h"

#include <cudaTypedefs.h>

#include <atomic>
#include <mutex>
#include <unordered_map>
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

// per device ordinal; written once with the same values by whichever thread gets there first (benign race)
constexpr int kMaxDevices = 128;
static int g_num_sms[kMaxDevices];
static int g_cc[kMaxDevices];
static std::atomic<int> g_known[kMaxDevices];

static int query_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  if (g_known[dev].load(std::memory_order_acquire)) return dev;
  int sms = 0, major = 0, minor = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  g_num_sms[dev] = sms;
  g_cc[dev] = major * 10 + minor;
  g_known[dev].store(1, std::memory_order_release);
  return dev;
}

int num_sms() {
  const int dev = query_device();
  return dev < 0 ? 0 : g_num_sms[dev];
}

int check_arch() {
  const int dev = query_device();
  const int cc = dev < 0 ? 0 : g_cc[dev];
  if (cc != 100 && cc != 103)
    return set_error(kUnsupportedArch,
                     "cm3p_b200 kernels are built for sm_100a only; current device reports sm_%d (no fallback path)",
                     cc);
  return kOk;
}

bool DeviceOnce::done() const {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return false;
  return (__atomic_load_n(&mask[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63)) & 1ull;
}
void DeviceOnce::mark() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return;
  __atomic_fetch_or(&mask[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
}

static std::atomic<int> g_options[kOptCount] = {{0}, {0}, {2}, {0}, {0}, {1}, {1}};

int get_option(int opt) { return (opt >= 0 && opt < kOptCount) ? g_options[opt].load(std::memory_order_relaxed) : 0; }
int set_option(int opt, int value) {
  if (opt < 0 || opt >= kOptCount) return set_error(kBadShape, "cm3p_set_option: unknown option %d", opt);
  g_options[opt].store(value, std::memory_order_relaxed);
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

// Encoded tensor maps are pure functions of (base, dims, pitch, box): cache them so that a steady-state step
// (same activation buffers from the caching allocator, same weights) does not pay a driver call per operand
// (3-6 per GEMM launch).  SURVEY.md 8b: "no global state except cached CUtensorMaps keyed by (ptr, shape)".
struct TmapKey {
  uint64_t base, inner, outer, pitch;
  uint32_t box_inner, box_outer;
  bool operator==(const TmapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && pitch == o.pitch && box_inner == o.box_inner &&
           box_outer == o.box_outer;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = k.base * 0x9E3779B97F4A7C15ull;
    h ^= (k.inner + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.outer + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= (k.pitch + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    h ^= ((static_cast<uint64_t>(k.box_inner) << 32 | k.box_outer) + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2));
    return static_cast<size_t>(h);
  }
};
static std::mutex g_tmap_mutex;
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmap_cache;
constexpr size_t kTmapCacheMax = 8192;

int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer) {
  const bool use_cache = get_option(kOptTmapCache) != 0;
  const TmapKey key{reinterpret_cast<uint64_t>(base), inner, outer, pitch_bytes, box_inner, box_outer};
  if (use_cache) {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    auto it = g_tmap_cache.find(key);
    if (it != g_tmap_cache.end()) {
      *map = it->second;
      return kOk;
    }
  }
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
  if (use_cache) {
    std::lock_guard<std::mutex> lock(g_tmap_mutex);
    if (g_tmap_cache.size() >= kTmapCacheMax) g_tmap_cache.clear();
    g_tmap_cache.emplace(key, *map);
  }
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

int encode_tmap_4d_bf16(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t pitches_bytes[3],
                        const uint32_t box[4]) {
  EncodeTiledFn fn = get_encode_fn();
  CM3P_REQUIRE(fn != nullptr, kDriverError, "cuTensorMapEncodeTiled entry point not available");
  CM3P_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, kBadAlignment, "TMA base %p not 16-byte aligned", base);
  for (int i = 0; i < 3; ++i)
    CM3P_REQUIRE((pitches_bytes[i] & 15) == 0, kBadAlignment, "TMA pitch %llu B not a multiple of 16",
                 (unsigned long long)pitches_bytes[i]);
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t st[3] = {pitches_bytes[0], pitches_bytes[1], pitches_bytes[2]};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), d, st, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  CM3P_REQUIRE(r == CUDA_SUCCESS, kDriverError, "cuTensorMapEncodeTiled(4d) failed with %d", (int)r);
  return kOk;
}

}  // namespace cm3p
