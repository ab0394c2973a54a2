// Helpers shared by the attention backward kernels (attn_bwd_sm100.cu, attn_bwd_win_sm100.cu): bf16 packing, rows of
// 128-byte-swizzled operand tiles written from registers, and the gradient epilogue with the inverse RoPE rotation.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace cm3p {
namespace bwd_detail {

__device__ __forceinline__ uint4 pack8f(const float* v) {
  return make_uint4(ptx::pack_bf16x2(v[0], v[1]), ptx::pack_bf16x2(v[2], v[3]), ptx::pack_bf16x2(v[4], v[5]),
                    ptx::pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ void unpack8f(const uint4& u, float* f) {
  float2 t;
  t = ptx::unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = ptx::unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = ptx::unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = ptx::unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}

// Row `t` of a [rows][64] bf16 K-major tile with the 128-byte swizzle the MMA descriptors expect:
// 16-byte unit u of row t lives at unit (u ^ (t & 7)).
__device__ __forceinline__ void store_row_units(uint8_t* tile, int t, int first_unit, const uint32_t (&packed)[16]) {
  uint8_t* row = tile + t * 128;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int unit = (first_unit + u) ^ (t & 7);
    *reinterpret_cast<uint4*>(row + unit * 16) =
        make_uint4(packed[u * 4], packed[u * 4 + 1], packed[u * 4 + 2], packed[u * 4 + 3]);
  }
}

__device__ __forceinline__ void store_zero_units(uint8_t* tile, int t, int first_unit) {
  uint8_t* row = tile + t * 128;
#pragma unroll
  for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(row + (((first_unit + u) ^ (t & 7)) << 4)) = make_uint4(0, 0, 0, 0);
}

// TMEM accumulator row (64 fp32 columns) -> optional inverse RoPE -> bf16 -> global.
//   forward: y1 = x1 c - x2 s, y2 = x2 c + x1 s   =>   dx1 = dy1 c + dy2 s, dx2 = dy2 c - dy1 s
__device__ __forceinline__ void store_grad_row(uint32_t taddr, __nv_bfloat16* dst, const float2* cs, bool valid) {
  uint32_t r1[32], r2[32];
  ptx::tmem_ld_32x32b_x32(taddr, r1);
  ptx::tmem_ld_32x32b_x32(taddr + 32, r2);
  ptx::tmem_ld_wait();
  if (!valid) return;
  float o1[32], o2[32];
  if (cs) {
    const float4* tab = reinterpret_cast<const float4*>(cs);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 f = __ldg(tab + k);  // (cos, sin) of frequencies 2k, 2k+1
      const float a0 = __uint_as_float(r1[2 * k]), b0 = __uint_as_float(r2[2 * k]);
      const float a1 = __uint_as_float(r1[2 * k + 1]), b1 = __uint_as_float(r2[2 * k + 1]);
      o1[2 * k] = a0 * f.x + b0 * f.y;
      o2[2 * k] = b0 * f.x - a0 * f.y;
      o1[2 * k + 1] = a1 * f.z + b1 * f.w;
      o2[2 * k + 1] = b1 * f.z - a1 * f.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      o1[k] = __uint_as_float(r1[k]);
      o2[k] = __uint_as_float(r2[k]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    *reinterpret_cast<uint4*>(dst + i * 8) = pack8f(o1 + i * 8);
    *reinterpret_cast<uint4*>(dst + 32 + i * 8) = pack8f(o2 + i * 8);
  }
}

}  // namespace bwd_detail
}  // namespace cm3p
