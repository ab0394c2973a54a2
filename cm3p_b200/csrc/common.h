// Host-side helpers shared by the kernels' launchers: error reporting across the C ABI and
// CUtensorMap encoding through the driver entry point (no link-time dependency on libcuda).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

enum Status : int {
  kOk = 0,
  kBadShape = -1,
  kBadAlignment = -2,
  kUnsupportedArch = -3,
  kCudaError = -4,
  kDriverError = -5,
};

int set_error(int code, const char* fmt, ...);
const char* last_error();

// device properties of the CURRENT device (cached per device ordinal)
int num_sms();
int check_arch();  // kOk on sm_100, error otherwise (no fallback)

// One-time-per-device marker for per-device state such as cudaFuncSetAttribute opt-ins (a `static bool` would
// configure only the first GPU a process touches).
struct DeviceOnce {
  unsigned long long mask[2] = {0ull, 0ull};  // device ordinals 0..127
  bool done() const;
  void mark();
};
#define CM3P_ENSURE_DYN_SMEM(kernel, bytes)                                                                   \
  do {                                                                                                        \
    static ::cm3p::DeviceOnce _once;                                                                          \
    if (!_once.done()) {                                                                                      \
      CM3P_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));        \
      _once.mark();                                                                                           \
    }                                                                                                         \
  } while (0)

// Run-time tuning knobs (cm3p_set_option in the C ABI; values are part of the ABI, see include/cm3p_b200.h).
// They exist for the tests (sweeps of the streaming depth) and for A/B measurements; the defaults are the product.
enum Option : int {
  kOptFwdBlocksPerCta = 0,   // attention forward: 256-query blocks streamed per CTA (0 = heuristic)
  kOptBwdOuterPerCta = 1,    // attention backward: outer tiles streamed per CTA (0 = heuristic)
  kOptGemmCluster = 2,       // 1 = no clusters, 2 = CTA pairs with one cta_group::2 MMA (default), 3 = pairs with B multicast
  kOptAttnForceTileKernels = 3,  // 1 = one-tile-per-CTA attention kernels for every sequence length
  kOptWgradDeterministic = 4,    // 1 = split-K partial sums added in split order through per-tile turnstiles (bit-
                                 // reproducible weight gradients, ~+70 % on those GEMMs), 0 = fp32 atomics (default)
  kOptTmapCache = 5,             // 1 = cache encoded CUtensorMaps by (pointer, shape, pitch, box) (default)
  kOptAttnWindowWalk = 6,        // 1 = fused band-walk backward for sliding-window layers (default), 0 = two-kernel v3
  kOptCount = 7,
};
int get_option(int opt);
int set_option(int opt, int value);

// 2-D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch in BYTES,
// 128-byte swizzle, zero fill out of bounds.  box_inner * 2 bytes must be <= 128.
int encode_tmap_2d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                        uint32_t box_inner, uint32_t box_outer);
// 3-D variant (inner, mid, outer) with byte pitches for mid and outer.
int encode_tmap_3d_bf16(CUtensorMap* map, const void* base, uint64_t inner, uint64_t mid, uint64_t outer,
                        uint64_t pitch_mid_bytes, uint64_t pitch_outer_bytes, uint32_t box_inner, uint32_t box_mid,
                        uint32_t box_outer);

// 4-D variant (d0 contiguous; byte pitches for d1..d3); boxes of {box0, box1, box2, 1}.
int encode_tmap_4d_bf16(CUtensorMap* map, const void* base, const uint64_t dims[4], const uint64_t pitches_bytes[3],
                        const uint32_t box[4]);

#define CM3P_CUDA_TRY(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return ::cm3p::set_error(::cm3p::kCudaError, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                                \
  } while (0)

#define CM3P_REQUIRE(cond, code, ...)                        \
  do {                                                       \
    if (!(cond)) return ::cm3p::set_error(code, __VA_ARGS__); \
  } while (0)

}  // namespace cm3p
