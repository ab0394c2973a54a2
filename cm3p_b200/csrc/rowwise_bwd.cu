// HBM-bound backward kernels of the CM3P train step: LayerNorm (+ residual-gradient add),
// embedding gather/LayerNorm (scatter-add into the embedding-table gradient, audio rows routed
// back to the audio encoder), GeGLU, GELU, bias (column) sums, pooling, L2 normalisation, the
// CLIP-style loss and the conv2 col2im.  Same style as rowwise_fwd.cu: 16-byte accesses, one warp
// per row with shuffle reductions, fp32 statistics, per-CTA partial sums for parameter gradients
// followed by a handful of fp32 atomics.
//
// These are the gradients autograd derives for the reference's torch ops (nn.LayerNorm, F.gelu,
// GeGLU in the third-party ModernBertMLP, cm3p/modeling_cm3p.py:27-62 loss + norm, :624-642
// pooling, :591-605 embedding gather / audio scatter, :501-502 conv + GELU).
#include <cuda_bf16.h>
#include <math.h>

#include "common.h"
#include "ptx.cuh"
#include "rowwise.h"

namespace cm3p {
namespace {

constexpr int MAXV = 4;  // up to 4 x (32 lanes x 8 elements) = hidden size 1024

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  float2 t;
  t = ptx::unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
  t = ptx::unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
  t = ptx::unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
  t = ptx::unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(ptx::pack_bf16x2(f[0], f[1]), ptx::pack_bf16x2(f[2], f[3]), ptx::pack_bf16x2(f[4], f[5]),
                    ptx::pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// d/dx gelu_erf(x) = Phi(x) + x * phi(x)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float y, dy;
  ptx::gelu_erf_and_grad(x, y, dy);
  return dy;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward (weight only, no bias):  y = (x - mean) * rstd * gamma
//   g = dy * gamma;  dx = rstd * (g - mean(g) - xhat * mean(g * xhat));  dgamma += sum_rows dy * xhat
// Statistics are recomputed from x (nothing but x was kept from the forward pass).  `dres`, when
// given, is the gradient that reached the same x through the residual connection and is added.
// GATHER variant = embedding layer: the row is re-gathered exactly as in embed_gather_ln_kernel and
// dx is scattered: audio rows -> d_audio[slot] (unique), token rows -> atomic add into d_tok[id].
template <bool GATHER, int NV>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                     const float* __restrict__ gamma, const __nv_bfloat16* __restrict__ dres,
                     __nv_bfloat16* __restrict__ dx, float* __restrict__ dgamma, int64_t rows, int H, float eps,
                     // GATHER only:
                     const int64_t* __restrict__ ids, const int32_t* __restrict__ src_index,
                     const int32_t* __restrict__ audio_slot, const __nv_bfloat16* __restrict__ tok_emb,
                     const __nv_bfloat16* __restrict__ audio_embeds, float* __restrict__ d_tok,
                     __nv_bfloat16* __restrict__ d_audio, int vocab) {
  // NV = 16-byte vectors per lane (H <= NV * 256).  The kernel is latency-bound unless enough rows are in
  // flight (ncu: 3 TB/s with 16 warps/SM and two dependent DRAM round trips per row), so: every load of a
  // row (x, dy, residual gradient) is issued up front and kept PACKED (bf16) in registers, statistics use
  // two shuffle rounds (sum & sum-of-squares, then the two dx moments), and the register budget allows
  // three CTAs (24 warps) per SM.  Only the dgamma partial sums live across rows.
  // dgamma partial sums: one private shared-memory row per warp (keeps 8*NV registers free for occupancy)
  __shared__ __align__(16) float red[8][NV * 256];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nvec = H >> 3;
  float* my_dg = red[warp];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    *reinterpret_cast<float4*>(my_dg + (lane + i * 32) * 8) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(my_dg + (lane + i * 32) * 8 + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invH = 1.f / H;
  // Software pipeline: the loads of the warp's NEXT row are issued before the current row is reduced, so a warp
  // always has a row of loads in flight (one row at a time left the memory system idle during the two shuffle
  // rounds and the dgamma update: 4.0 TB/s).
  const int64_t row_step = static_cast<int64_t>(gridDim.x) * 8;
  uint4 nx[NV], ng[NV], nr[NV];
  int nslot = -1;
  int64_t nid = 0;
  auto load_row = [&](int64_t row) {
    const __nv_bfloat16* src;
    nslot = -1;
    nid = 0;
    if constexpr (GATHER) {
      const int64_t flat = src_index ? src_index[row] : row;
      nslot = audio_slot ? audio_slot[row] : -1;
      if (nslot >= 0 && audio_embeds) {
        src = audio_embeds + static_cast<int64_t>(nslot) * H;
      } else {
        nslot = -1;
        nid = ids[flat];
        nid = nid < 0 ? 0 : (nid >= vocab ? vocab - 1 : nid);
        src = tok_emb + nid * H;
      }
    } else {
      src = x + row * H;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      nx[i] = ng[i] = nr[i] = make_uint4(0, 0, 0, 0);
      if (vi < nvec) {
        nx[i] = *reinterpret_cast<const uint4*>(src + vi * 8);
        ng[i] = *reinterpret_cast<const uint4*>(dy + row * H + vi * 8);
        if constexpr (!GATHER) {
          if (dres) nr[i] = *reinterpret_cast<const uint4*>(dres + row * H + vi * 8);
        }
      }
    }
  };
  const int64_t first_row = static_cast<int64_t>(blockIdx.x) * 8 + warp;
  if (first_row < rows) load_row(first_row);
  for (int64_t row = first_row; row < rows; row += row_step) {
    uint4 px[NV], pg[NV], pr[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      px[i] = nx[i];
      pg[i] = ng[i];
      pr[i] = nr[i];
    }
    const int slot = nslot;
    const int64_t id = nid;
    if (row + row_step < rows) load_row(row + row_step);
    float s = 0.f, q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float v[8];
      unpack8(px[i], v);  // zeros beyond nvec contribute nothing
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        s += v[k];
        q += v[k] * v[k];
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    const float mean = s * invH;
    const float rstd = rsqrtf(fmaxf(q * invH - mean * mean, 0.f) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float v[8], g[8], gam[8];
        unpack8(px[i], v);
        unpack8(pg[i], g);
        *reinterpret_cast<float4*>(gam) = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        *reinterpret_cast<float4*>(gam + 4) = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
        float d[8];
        *reinterpret_cast<float4*>(d) = *reinterpret_cast<const float4*>(my_dg + vi * 8);
        *reinterpret_cast<float4*>(d + 4) = *reinterpret_cast<const float4*>(my_dg + vi * 8 + 4);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float xh = (v[k] - mean) * rstd;
          d[k] += g[k] * xh;
          const float gg = g[k] * gam[k];
          s1 += gg;
          s2 += gg * xh;
        }
        *reinterpret_cast<float4*>(my_dg + vi * 8) = *reinterpret_cast<const float4*>(d);
        *reinterpret_cast<float4*>(my_dg + vi * 8 + 4) = *reinterpret_cast<const float4*>(d + 4);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
      s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    s1 *= invH;
    s2 *= invH;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        float v[8], g[8], gam[8], o[8];
        unpack8(px[i], v);
        unpack8(pg[i], g);
        *reinterpret_cast<float4*>(gam) = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
        *reinterpret_cast<float4*>(gam + 4) = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = rstd * (g[k] * gam[k] - s1 - (v[k] - mean) * rstd * s2);
        if constexpr (GATHER) {
          if (slot >= 0) {
            if (d_audio) *reinterpret_cast<uint4*>(d_audio + static_cast<int64_t>(slot) * H + vi * 8) = pack8(o);
          } else if (d_tok) {
            float* dst = d_tok + id * H + vi * 8;
            red_add_f32x4(dst, o[0], o[1], o[2], o[3]);
            red_add_f32x4(dst + 4, o[4], o[5], o[6], o[7]);
          }
        } else {
          if (dres) {
            float r[8];
            unpack8(pr[i], r);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[k] += r[k];
          }
          *reinterpret_cast<uint4*>(dx + row * H + vi * 8) = pack8(o);
        }
      }
    }
  }
  if (dgamma == nullptr) return;
  // per-CTA reduction of the gamma gradient, then one atomic per column
  __syncthreads();
  for (int c = threadIdx.x; c < H; c += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][c];
    atomicAdd(dgamma + c, t);
  }
}

// ------------------------------------------------------------------------------------------------
// GeGLU backward on the interleaved pre-activation kept by the forward GEMM (EPI_GEGLU_SAVE):
// ug[t, 32*G + 0..15] = u, ug[t, 32*G + 16..31] = gate of channels 16*G..16*G+15.
//   h = gelu(u) * gate (recomputed: it is the A-side operand of the Wo weight gradient)
//   du = dh * gate * gelu'(u);  dgate = dh * gelu(u)
__global__ void __launch_bounds__(256)
geglu_bwd_kernel(const __nv_bfloat16* __restrict__ ug, const __nv_bfloat16* __restrict__ dh,
                 __nv_bfloat16* __restrict__ dug, __nv_bfloat16* __restrict__ h, int64_t rows, int I) {
  const int vec_per_row = I >> 3;
  const int64_t total = rows * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vec_per_row;
    const int c = static_cast<int>(i % vec_per_row) * 8;  // first of 8 channels
    const int64_t ucol = static_cast<int64_t>(c >> 4) * 32 + (c & 15);
    float u[8], g[8], d[8], du[8], dgt[8], hh[8];
    unpack8(*reinterpret_cast<const uint4*>(ug + r * 2 * I + ucol), u);
    unpack8(*reinterpret_cast<const uint4*>(ug + r * 2 * I + ucol + 16), g);
    unpack8(*reinterpret_cast<const uint4*>(dh + r * I + c), d);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float ge, gd;
      ptx::gelu_erf_and_grad(u[k], ge, gd);
      hh[k] = ge * g[k];
      du[k] = d[k] * g[k] * gd;
      dgt[k] = d[k] * ge;
    }
    *reinterpret_cast<uint4*>(dug + r * 2 * I + ucol) = pack8(du);
    *reinterpret_cast<uint4*>(dug + r * 2 * I + ucol + 16) = pack8(dgt);
    if (h) *reinterpret_cast<uint4*>(h + r * I + c) = pack8(hh);
  }
}

// y = gelu(z) / dz = dy * gelu'(z): the training path keeps the pre-activation z of the conv and
// projector GELUs and applies the activation in a separate pass.
__global__ void __launch_bounds__(256)
gelu_fwd_kernel(const __nv_bfloat16* __restrict__ z, __nv_bfloat16* __restrict__ y, int64_t nvec) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float f[8];
    unpack8(reinterpret_cast<const uint4*>(z)[i], f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = ptx::gelu_erf(f[k]);
    reinterpret_cast<uint4*>(y)[i] = pack8(f);
  }
}
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dy,
                __nv_bfloat16* __restrict__ dz, int64_t nvec) {
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float f[8], d[8];
    unpack8(reinterpret_cast<const uint4*>(z)[i], f);
    unpack8(reinterpret_cast<const uint4*>(dy)[i], d);
#pragma unroll
    for (int k = 0; k < 8; ++k) d[k] *= gelu_erf_grad(f[k]);
    reinterpret_cast<uint4*>(dz)[i] = pack8(d);
  }
}

// out[c] += sum_r dy[r, c]  (bias gradients).  One CTA = 64 columns x a slab of rows.
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ dy, float* __restrict__ out, int64_t rows, int N, int rows_per_cta) {
  __shared__ float red[32][64 + 1];
  const int c0 = blockIdx.x * 64;
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_cta;
  const int64_t r1 = min(rows, r0 + rows_per_cta);
  const int v = threadIdx.x & 7, rl = threadIdx.x >> 3;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 + v * 8 < N) {
    for (int64_t r = r0 + rl; r < r1; r += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(dy + r * N + c0 + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][v * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < N) {
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    atomicAdd(out + c0 + threadIdx.x, s);
  }
}

// ------------------------------------------------------------------------------------------------
// Pooling backward.  mode 0 (first token): dh[cu[b]] (+)= dp[b], every other row 0 (or untouched when
// accumulating).  mode 1 (masked mean): dh[t] (+)= dp[b] / max(len_b, 1e-9).  One CTA per sequence.
__global__ void __launch_bounds__(256)
pool_bwd_kernel(const __nv_bfloat16* __restrict__ dp, const int32_t* __restrict__ cu, __nv_bfloat16* __restrict__ dh,
                int mode, int accumulate, int H) {
  const int b = blockIdx.x;
  const int start = cu[b], len = cu[b + 1] - start;
  const int nvec = H >> 3;
  const float w = mode == 0 ? 1.f : 1.f / fmaxf(static_cast<float>(len), 1e-9f);
  const int rows = (mode == 0 && accumulate) ? min(len, 1) : len;
  for (int i = threadIdx.x; i < rows * nvec; i += blockDim.x) {
    const int t = i / nvec, v = i % nvec;
    float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (mode == 1 || t == 0) {
      unpack8(*reinterpret_cast<const uint4*>(dp + static_cast<int64_t>(b) * H + v * 8), o);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] *= w;
    }
    __nv_bfloat16* dst = dh + static_cast<int64_t>(start + t) * H + v * 8;
    if (accumulate) {
      float old[8];
      unpack8(*reinterpret_cast<const uint4*>(dst), old);
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] += old[k];
    }
    *reinterpret_cast<uint4*>(dst) = pack8(o);
  }
}

// emb = e * inv, inv = 1/|e|  ->  de = inv * (demb - emb * <demb, emb>)   (no epsilon, quirk Q3)
__global__ void __launch_bounds__(256)
l2norm_bwd_kernel(const float* __restrict__ e, const float* __restrict__ inv_norm, const float* __restrict__ demb,
                  __nv_bfloat16* __restrict__ de, int rows, int P) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float inv = inv_norm[row];
  const float* er = e + static_cast<int64_t>(row) * P;
  const float* dr = demb + static_cast<int64_t>(row) * P;
  float dot = 0.f;
  for (int i = lane; i < P; i += 32) dot += dr[i] * er[i] * inv;
  dot = warp_sum(dot);
  for (int i = lane; i < P; i += 32)
    de[static_cast<int64_t>(row) * P + i] = __float2bfloat16(inv * (dr[i] - er[i] * inv * dot));
}

// ------------------------------------------------------------------------------------------------
// CLIP loss backward (cm3p/modeling_cm3p.py:33-51) on S [R = Bm*V, Bb] fp32:
//   dS[r, j] = g/2 * ( (softmax_col_j(S)[r] - [r == j*V + t_j]) / Bb
//                    + [r == (r/V)*V + t_{r/V}] * (softmax_row_r(S)[j] - [j == r/V]) / Bm )
// and d logit_scale = sum dS * S (S = exp(logit_scale) * cos).  g = upstream gradient (device scalar).
__global__ void __launch_bounds__(256)
clip_loss_bwd_kernel(const float* __restrict__ S, const int32_t* __restrict__ true_idx,
                     const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                     const float* __restrict__ grad_out, __nv_bfloat16* __restrict__ dS, int64_t ld_ds,
                     float* __restrict__ dlogit_scale, int Bm, int V, int Bb) {
  __shared__ float red[8];
  const int64_t total = static_cast<int64_t>(Bm) * V * Bb;
  const float g = 0.5f * (grad_out ? *grad_out : 1.f);
  float acc = 0.f;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % Bb);
    const int r = static_cast<int>(i / Bb);
    const int m = r / V;
    const float s = S[i];
    float d = (__expf(s - col_lse[j]) - ((r == j * V + true_idx[j]) ? 1.f : 0.f)) / Bb;
    if (r == m * V + true_idx[m]) d += (__expf(s - row_lse[m]) - ((j == m) ? 1.f : 0.f)) / Bm;
    d *= g;
    dS[static_cast<int64_t>(r) * ld_ds + j] = __float2bfloat16(d);
    acc += d * s;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0 && dlogit_scale) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    atomicAdd(dlogit_scale, t);
  }
}

// ------------------------------------------------------------------------------------------------
// conv2 (k=3, pad 1, stride 2) input gradient: col2im of dA2 [(b, t), j*C + c] onto the channels-last
// conv1 output grid, fused with conv1's GELU backward:
//   dy1[b, ts, c] = sum_{j, t : 2t + j - 1 = ts} dA2[(b, t), j*C + c];   dz1 = dy1 * gelu'(z1)
__global__ void __launch_bounds__(256)
conv2_col2im_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ da2, const __nv_bfloat16* __restrict__ z1,
                             __nv_bfloat16* __restrict__ dz1, int B, int F, int C) {
  const int Fo = F / 2;
  const int vec_per_row = C >> 3;
  const int64_t total = static_cast<int64_t>(B) * F * vec_per_row;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % vec_per_row) * 8;
    const int64_t r = i / vec_per_row;  // b * F + ts
    const int ts = static_cast<int>(r % F);
    const int b = static_cast<int>(r / F);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int num = ts + 1 - j;
      if (num < 0 || (num & 1)) continue;
      const int t = num >> 1;
      if (t >= Fo) continue;
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(da2 + (static_cast<int64_t>(b) * Fo + t) * (3 * C) + j * C + c), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
    float z[8];
    unpack8(*reinterpret_cast<const uint4*>(z1 + r * C + c), z);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= gelu_erf_grad(z[k]);
    *reinterpret_cast<uint4*>(dz1 + r * C + c) = pack8(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Cross-entropy over a vocabulary (MLM head, cm3p/modeling_cm3p.py:994-996 / :1365-1367 ->
// transformers ForMaskedLMLoss: F.cross_entropy(logits.float(), labels, ignore_index=-100)) and the
// 2..n-way classifier loss (:1196-1212).  One warp per row, online max / sum-exp in fp32.
//   target(row) = labels[src_index ? src_index[row] : row]; rows whose target == ignore_index are skipped.
//   fwd: row_lse[row]; loss_sum += lse - x[target]; count += 1
//   bwd: logits <- (softmax(x) - onehot(target)) * scale  in place (0 for ignored rows and pad columns)
__global__ void __launch_bounds__(256)
vocab_ce_fwd_kernel(const __nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                    const int32_t* __restrict__ src_index, int ignore_index, float* __restrict__ row_lse,
                    float* __restrict__ loss_sum, float* __restrict__ count, int64_t rows, int V) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t tgt = labels[src_index ? src_index[row] : row];
  if (tgt == ignore_index || tgt < 0 || tgt >= V) {
    if (lane == 0) row_lse[row] = 0.f;
    return;
  }
  const __nv_bfloat16* x = logits + row * ld;
  float m = -INFINITY, s = 0.f;
  for (int c = lane; c < V; c += 32) {
    const float v = __bfloat162float(x[c]);
    const float mn = fmaxf(m, v);
    s = s * __expf(m - mn) + __expf(v - mn);
    m = mn;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    s = ((m == -INFINITY) ? 0.f : s * __expf(m - mn)) + ((m2 == -INFINITY) ? 0.f : s2 * __expf(m2 - mn));
    m = mn;
  }
  if (lane == 0) {
    const float lse = m + logf(s);
    row_lse[row] = lse;
    atomicAdd(loss_sum, lse - __bfloat162float(x[tgt]));
    atomicAdd(count, 1.f);
  }
}

__global__ void __launch_bounds__(256)
vocab_ce_bwd_kernel(__nv_bfloat16* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                    const int32_t* __restrict__ src_index, int ignore_index, const float* __restrict__ row_lse,
                    const float* __restrict__ scale, int64_t rows, int V) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t tgt = labels[src_index ? src_index[row] : row];
  __nv_bfloat16* x = logits + row * ld;
  const bool skip = (tgt == ignore_index || tgt < 0 || tgt >= V);
  const float lse = row_lse[row], sc = *scale;
  for (int c = lane; c < ld; c += 32) {
    float g = 0.f;
    if (!skip && c < V) g = (__expf(__bfloat162float(x[c]) - lse) - ((c == tgt) ? 1.f : 0.f)) * sc;
    x[c] = __float2bfloat16(g);
  }
}

// out[r] = x[index[r]] and x[index[r]] += dx[r] (unique indices): sparse MLM prediction (:1349-1357)
__global__ void __launch_bounds__(256)
gather_rows_i32_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ index,
                       __nv_bfloat16* __restrict__ out, int64_t rows, int H) {
  const int nvec = H >> 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / nvec;
    const int v = static_cast<int>(i % nvec);
    *reinterpret_cast<uint4*>(out + r * H + v * 8) =
        *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(index[r]) * H + v * 8);
  }
}
__global__ void __launch_bounds__(256)
scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ dx_rows, const int32_t* __restrict__ index,
                        __nv_bfloat16* __restrict__ dx, int64_t rows, int H) {
  const int nvec = H >> 3;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < rows * nvec;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / nvec;
    const int v = static_cast<int>(i % nvec);
    __nv_bfloat16* dst = dx + static_cast<int64_t>(index[r]) * H + v * 8;
    float a[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(dst), a);
    unpack8(*reinterpret_cast<const uint4*>(dx_rows + r * H + v * 8), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] += b[k];
    *reinterpret_cast<uint4*>(dst) = pack8(a);
  }
}

inline int grid_for(int64_t work_items, int threads) {
  int64_t g = (work_items + threads - 1) / threads;
  const int64_t cap = static_cast<int64_t>(num_sms() > 0 ? num_sms() : 148) * 16;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace

int layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* dres, void* dx, float* dgamma,
                  int64_t rows, int H, float eps, cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0 && H <= MAXV * 256, kBadShape, "layernorm_bwd: hidden size %d must be a multiple of 8, <= %d", H,
               MAXV * 256);
  if (rows == 0) return kOk;
  const int64_t want = (rows + 7) / 8;
  const int grid = static_cast<int>(want < 4LL * num_sms() ? want : 4LL * num_sms());  // 2 waves of 2 CTAs per SM
#define CM3P_LN_BWD(NV)                                                                                             \
  layernorm_bwd_kernel<false, NV><<<grid, 256, 0, stream>>>(                                                        \
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy), gamma,                 \
      reinterpret_cast<const __nv_bfloat16*>(dres), reinterpret_cast<__nv_bfloat16*>(dx), dgamma, rows, H, eps,     \
      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0)
  if (H <= 256) CM3P_LN_BWD(1);
  else if (H <= 512) CM3P_LN_BWD(2);
  else if (H <= 768) CM3P_LN_BWD(3);
  else CM3P_LN_BWD(4);
#undef CM3P_LN_BWD
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int embed_gather_ln_bwd(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot, const void* tok_emb,
                        const void* audio_embeds, const float* gamma, const void* dy, float* d_tok_emb,
                        void* d_audio_embeds, float* dgamma, int64_t rows, int H, int vocab, float eps,
                        cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0 && H <= MAXV * 256, kBadShape, "embed_bwd: hidden size %d must be a multiple of 8, <= %d", H,
               MAXV * 256);
  if (rows == 0) return kOk;
  const int64_t want = (rows + 7) / 8;
  const int grid = static_cast<int>(want < 4LL * num_sms() ? want : 4LL * num_sms());  // 2 waves of 2 CTAs per SM
#define CM3P_LN_BWD(NV)                                                                                            \
  layernorm_bwd_kernel<true, NV><<<grid, 256, 0, stream>>>(                                                        \
      nullptr, reinterpret_cast<const __nv_bfloat16*>(dy), gamma, nullptr, nullptr, dgamma, rows, H, eps, ids,     \
      src_index, audio_slot, reinterpret_cast<const __nv_bfloat16*>(tok_emb),                                      \
      reinterpret_cast<const __nv_bfloat16*>(audio_embeds), d_tok_emb,                                             \
      reinterpret_cast<__nv_bfloat16*>(d_audio_embeds), vocab)
  if (H <= 256) CM3P_LN_BWD(1);
  else if (H <= 512) CM3P_LN_BWD(2);
  else if (H <= 768) CM3P_LN_BWD(3);
  else CM3P_LN_BWD(4);
#undef CM3P_LN_BWD
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int geglu_bwd(const void* ug, const void* dh, void* dug, void* h, int64_t rows, int I, cudaStream_t stream) {
  CM3P_REQUIRE(I % 16 == 0, kBadShape, "geglu_bwd: intermediate size %d must be a multiple of 16", I);
  if (rows == 0) return kOk;
  geglu_bwd_kernel<<<grid_for(rows * (I / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(ug), reinterpret_cast<const __nv_bfloat16*>(dh),
      reinterpret_cast<__nv_bfloat16*>(dug), reinterpret_cast<__nv_bfloat16*>(h), rows, I);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int gelu_fwd(const void* z, void* y, int64_t n, cudaStream_t stream) {
  CM3P_REQUIRE(n % 8 == 0, kBadShape, "gelu_fwd: element count must be a multiple of 8");
  if (n == 0) return kOk;
  gelu_fwd_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(z),
                                                            reinterpret_cast<__nv_bfloat16*>(y), n / 8);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int gelu_bwd(const void* z, const void* dy, void* dz, int64_t n, cudaStream_t stream) {
  CM3P_REQUIRE(n % 8 == 0, kBadShape, "gelu_bwd: element count must be a multiple of 8");
  if (n == 0) return kOk;
  gelu_bwd_kernel<<<grid_for(n / 8, 256), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(z),
                                                            reinterpret_cast<const __nv_bfloat16*>(dy),
                                                            reinterpret_cast<__nv_bfloat16*>(dz), n / 8);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int colsum_f32(const void* dy, float* out, int64_t rows, int N, cudaStream_t stream) {
  CM3P_REQUIRE(N % 8 == 0, kBadShape, "colsum: N %% 8 required");
  if (rows == 0) return kOk;
  const int col_blocks = (N + 63) / 64;
  int64_t row_blocks = (4LL * num_sms() + col_blocks - 1) / col_blocks;
  if (row_blocks > (rows + 255) / 256) row_blocks = (rows + 255) / 256;
  if (row_blocks < 1) row_blocks = 1;
  const int rows_per_cta = static_cast<int>((rows + row_blocks - 1) / row_blocks);
  dim3 grid(col_blocks, static_cast<unsigned>((rows + rows_per_cta - 1) / rows_per_cta));
  colsum_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dy), out, rows, N, rows_per_cta);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int pool_bwd(const void* dpooled, const int32_t* cu_seqlens, void* dhidden, int mode, int accumulate, int batch, int H,
             cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0, kBadShape, "pool_bwd: H %% 8 required");
  if (batch == 0) return kOk;
  pool_bwd_kernel<<<batch, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(dpooled), cu_seqlens,
                                             reinterpret_cast<__nv_bfloat16*>(dhidden), mode, accumulate, H);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int l2norm_bwd(const float* e, const float* inv_norm, const float* dembeds, void* de_bf16, int rows, int P,
               cudaStream_t stream) {
  if (rows == 0) return kOk;
  l2norm_bwd_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(e, inv_norm, dembeds, reinterpret_cast<__nv_bfloat16*>(de_bf16),
                                                        rows, P);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int clip_loss_bwd(const float* S, const int32_t* true_idx, const float* row_lse, const float* col_lse,
                  const float* grad_out, void* dS, int64_t ld_ds, float* dlogit_scale, int Bm, int V, int Bb,
                  cudaStream_t stream) {
  CM3P_REQUIRE(Bm == Bb && Bm > 0 && V > 0, kBadShape, "clip_loss_bwd: need Bm == Bb > 0 (got %d, %d)", Bm, Bb);
  CM3P_REQUIRE(ld_ds >= Bb, kBadShape, "clip_loss_bwd: ld_ds %lld < Bb", (long long)ld_ds);
  const int64_t total = static_cast<int64_t>(Bm) * V * Bb;
  clip_loss_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(S, true_idx, row_lse, col_lse, grad_out,
                                                                 reinterpret_cast<__nv_bfloat16*>(dS), ld_ds,
                                                                 dlogit_scale, Bm, V, Bb);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int conv2_col2im_gelu_bwd(const void* da2, const void* z1, void* dz1, int B, int F, int C, cudaStream_t stream) {
  CM3P_REQUIRE(C % 8 == 0 && F % 2 == 0, kBadShape, "col2im: C %% 8 and F %% 2 required (C=%d F=%d)", C, F);
  const int64_t total = static_cast<int64_t>(B) * F * (C / 8);
  if (total == 0) return kOk;
  conv2_col2im_gelu_bwd_kernel<<<grid_for(total, 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(da2), reinterpret_cast<const __nv_bfloat16*>(z1),
      reinterpret_cast<__nv_bfloat16*>(dz1), B, F, C);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int vocab_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index, int ignore_index,
                 float* row_lse, float* loss_sum, float* count, int64_t rows, int V, cudaStream_t stream) {
  CM3P_REQUIRE(ld >= V && V > 0, kBadShape, "vocab_ce: ld %lld < V %d", (long long)ld, V);
  if (rows == 0) return kOk;
  vocab_ce_fwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(logits), ld, labels, src_index, ignore_index, row_lse, loss_sum, count, rows,
      V);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int vocab_ce_bwd(void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index, int ignore_index,
                 const float* row_lse, const float* scale, int64_t rows, int V, cudaStream_t stream) {
  CM3P_REQUIRE(ld >= V && V > 0 && scale != nullptr, kBadShape, "vocab_ce_bwd: bad arguments");
  if (rows == 0) return kOk;
  vocab_ce_bwd_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      reinterpret_cast<__nv_bfloat16*>(logits), ld, labels, src_index, ignore_index, row_lse, scale, rows, V);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int gather_rows_i32(const void* x, const int32_t* index, void* out, int64_t rows, int H, cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0, kBadShape, "gather_rows: H %% 8 required");
  if (rows == 0) return kOk;
  gather_rows_i32_kernel<<<grid_for(rows * (H / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), index, reinterpret_cast<__nv_bfloat16*>(out), rows, H);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int scatter_add_rows(const void* dx_rows, const int32_t* index, void* dx, int64_t rows, int H, cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0, kBadShape, "scatter_add_rows: H %% 8 required");
  if (rows == 0) return kOk;
  scatter_add_rows_kernel<<<grid_for(rows * (H / 8), 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(dx_rows), index, reinterpret_cast<__nv_bfloat16*>(dx), rows, H);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
