// HBM-bound forward kernels: LayerNorm, token-embedding gather (+ audio-slot select) + LayerNorm,
// the conv input transpose, pooling, L2 normalisation and the CLIP-style loss.  All are
// vectorised (16-byte accesses), coalesced, and use warp-shuffle reductions; rows are bf16, all
// statistics are fp32.
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
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
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

// One warp normalises one row held in registers: y = (x - mean) * rstd * gamma  (no beta).
__device__ __forceinline__ void ln_row(const __nv_bfloat16* __restrict__ src, const float* __restrict__ gamma,
                                       __nv_bfloat16* __restrict__ dst, float2* stat_out, int H, float eps, int lane) {
  const int nvec = H >> 3;
  float v[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 u = *reinterpret_cast<const uint4*>(src + vi * 8);
      unpack8(u, v[i]);
#pragma unroll
      for (int k = 0; k < 8; ++k) s += v[i][k];
    }
  }
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float d = v[i][k] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
  if (stat_out && lane == 0) *stat_out = make_float2(mean, rstd);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
      const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
      float o[8];
      o[0] = (v[i][0] - mean) * rstd * g0.x; o[1] = (v[i][1] - mean) * rstd * g0.y;
      o[2] = (v[i][2] - mean) * rstd * g0.z; o[3] = (v[i][3] - mean) * rstd * g0.w;
      o[4] = (v[i][4] - mean) * rstd * g1.x; o[5] = (v[i][5] - mean) * rstd * g1.y;
      o[6] = (v[i][6] - mean) * rstd * g1.z; o[7] = (v[i][7] - mean) * rstd * g1.w;
      *reinterpret_cast<uint4*>(dst + vi * 8) = pack8(o);
    }
  }
}

// NR rows per warp with every load issued before the first use: at 6.4 TB/s the memory system needs ~80 KB in
// flight per SM, which one 1.5 KB row per warp (plus its dependent shuffle reductions) does not sustain
// (measured 4.6 TB/s with one row per warp).
#ifndef CM3P_LN_NR
#define CM3P_LN_NR 2    // rows per warp
#define CM3P_LN_MINB 3  // CTAs per SM the register budget is sized for (measured best: 2 rows, 3 CTAs -> 5.8 TB/s)
#endif
template <int NR>
__global__ void __launch_bounds__(256, CM3P_LN_MINB) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            __nv_bfloat16* __restrict__ y, float2* __restrict__ stats,
                                                            int64_t rows, int H, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = (static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5)) * NR;
  if (row0 >= rows) return;
  const int nvec = H >> 3;
  uint4 raw[NR][MAXV];
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      raw[r][i] = make_uint4(0, 0, 0, 0);
      if (vi < nvec && row0 + r < rows) raw[r][i] = *reinterpret_cast<const uint4*>(x + (row0 + r) * H + vi * 8);
    }
  float s[NR], q[NR];
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    s[r] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      float f[8];
      unpack8(raw[r][i], f);  // zero beyond the row
#pragma unroll
      for (int k = 0; k < 8; ++k) s[r] += f[k];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < NR; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    const float mean = s[r] / H;
    s[r] = mean;
    q[r] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      if (lane + i * 32 < nvec) {
        float f[8];
        unpack8(raw[r][i], f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float d = f[k] - mean;
          q[r] += d * d;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int r = 0; r < NR; ++r) q[r] += __shfl_xor_sync(0xffffffffu, q[r], o);
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    if (row0 + r >= rows) break;
    const float mean = s[r];
    const float rstd = rsqrtf(q[r] / H + eps);
    if (stats && lane == 0) stats[row0 + r] = make_float2(mean, rstd);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      const int vi = lane + i * 32;
      if (vi < nvec) {
        const float4 g0 = *reinterpret_cast<const float4*>(gamma + vi * 8);
        const float4 g1 = *reinterpret_cast<const float4*>(gamma + vi * 8 + 4);
        float f[8], o[8];
        unpack8(raw[r][i], f);
        o[0] = (f[0] - mean) * rstd * g0.x; o[1] = (f[1] - mean) * rstd * g0.y;
        o[2] = (f[2] - mean) * rstd * g0.z; o[3] = (f[3] - mean) * rstd * g0.w;
        o[4] = (f[4] - mean) * rstd * g1.x; o[5] = (f[5] - mean) * rstd * g1.y;
        o[6] = (f[6] - mean) * rstd * g1.z; o[7] = (f[7] - mean) * rstd * g1.w;
        *reinterpret_cast<uint4*>(y + (row0 + r) * H + vi * 8) = pack8(o);
      }
    }
  }
}

// x0[t] = LN( is_audio(t) ? audio_embeds[audio_slot[t]] : tok_emb[ids[src_index[t]]] )
// Reference: modeling_cm3p.py:591-592 (gather), :603-605 (boolean-mask scatter of audio embeddings,
// row-major order == running count of [AUDIO] tokens), ModernBertEmbeddings.norm.
__global__ void __launch_bounds__(256)
embed_gather_ln_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ src_index,
                       const int32_t* __restrict__ audio_slot, const __nv_bfloat16* __restrict__ tok_emb,
                       const __nv_bfloat16* __restrict__ audio_embeds, const float* __restrict__ gamma,
                       __nv_bfloat16* __restrict__ y, float2* __restrict__ stats, int64_t rows, int H, int vocab,
                       float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t flat = src_index ? src_index[row] : row;
  const int slot = audio_slot ? audio_slot[row] : -1;
  const __nv_bfloat16* src;
  if (slot >= 0 && audio_embeds) {
    src = audio_embeds + static_cast<int64_t>(slot) * H;
  } else {
    int64_t id = ids[flat];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    src = tok_emb + id * H;
  }
  ln_row(src, gamma, y + row * H, stats ? stats + row : nullptr, H, eps, lane);
}

// Log-mel input [B, C, F] fp32 (channels-first, as the processor produces it) -> [B, F, C] bf16 channels-last, the
// layout the implicit-GEMM conv reads through its tensor map.  32 x 32 tiles through shared memory: both the read
// (along F) and the write (along C) are coalesced.
__global__ void __launch_bounds__(256)
transpose_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int C, int F) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int f0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const float* xb = x + static_cast<int64_t>(b) * C * F;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = c0 + ty + k, f = f0 + tx;
    tile[ty + k][tx] = (c < C && f < F) ? xb[static_cast<int64_t>(c) * F + f] : 0.f;
  }
  __syncthreads();
  __nv_bfloat16* ob = out + static_cast<int64_t>(b) * F * C;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int f = f0 + ty + k, c = c0 + tx;
    if (f < F && c < C) ob[static_cast<int64_t>(f) * C + c] = __float2bfloat16(tile[tx][ty + k]);
  }
}

// Row gather: out[r] = x[index[r]]  (first-token pooling: index = cu_seqlens[:-1]).
__global__ void __launch_bounds__(256)
gather_rows_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ index,
                   __nv_bfloat16* __restrict__ out, int rows, int H) {
  const int nvec = H >> 3;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < rows * nvec; i += gridDim.x * blockDim.x) {
    const int r = i / nvec, v = i % nvec;
    *reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * H + v * 8) =
        *reinterpret_cast<const uint4*>(x + static_cast<int64_t>(index[r]) * H + v * 8);
  }
}

// Masked-mean pooling over the real tokens of each sequence (fp32 accumulate, quirk Q5):
// out[b] = sum_t x[cu[b] + t] / max(len_b, 1e-9).  One CTA per (sequence, 64-column slab).
__global__ void __launch_bounds__(256)
mean_pool_kernel(const __nv_bfloat16* __restrict__ x, const int32_t* __restrict__ cu, __nv_bfloat16* __restrict__ out,
                 int H) {
  __shared__ float red[32][64 + 1];
  const int b = blockIdx.x;  // sequences in grid.x: B*V can exceed the 65535 limit of grid.y
  const int c0 = blockIdx.y * 64;
  const int start = cu[b], len = cu[b + 1] - start;
  const int v = threadIdx.x & 7;   // 8 vectors of 8 columns
  const int rl = threadIdx.x >> 3;  // 32 row lanes
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 + v * 8 < H) {
    for (int t = rl; t < len; t += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const uint4*>(x + static_cast<int64_t>(start + t) * H + c0 + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][v * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < H) {
    float s = 0.f;
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    out[static_cast<int64_t>(b) * H + c0 + threadIdx.x] = __float2bfloat16(s / fmaxf(static_cast<float>(len), 1e-9f));
  }
}

// e / sqrt(sum e^2), no epsilon (quirk Q3, modeling_cm3p.py:54-62).  fp32 in, fp32 + bf16 out.
__global__ void __launch_bounds__(256)
l2norm_rows_kernel(const float* __restrict__ e, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                   float* __restrict__ inv_norm, int rows, int P) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* src = e + static_cast<int64_t>(row) * P;
  float s = 0.f;
  for (int i = lane; i < P; i += 32) s += src[i] * src[i];
  const float inv = 1.f / sqrtf(warp_sum(s));
  if (inv_norm && lane == 0) inv_norm[row] = inv;
  for (int i = lane; i < P; i += 32) {
    const float v = src[i] * inv;
    if (out_f32) out_f32[static_cast<int64_t>(row) * P + i] = v;
    if (out_bf16) out_bf16[static_cast<int64_t>(row) * P + i] = __float2bfloat16(v);
  }
}

// ---------------------------------------------------------------------------------------------
// CLIP-style symmetric loss (modeling_cm3p.py:33-51) on S = logits_per_metadata viewed as
// [R = Bm*V rows, Bb columns] fp32:
//   metadata term: CE over row (i*V + t_i) with target column i
//   beatmap  term: CE over column j (all R rows are negatives, quirk Q2) with target row j*V + t_j
// Pass 1: blocks [0, col_blocks) reduce 32 columns each; the remaining blocks take 8 rows each.
__global__ void __launch_bounds__(256)
clip_lse_kernel(const float* __restrict__ S, const int32_t* __restrict__ true_idx, float* __restrict__ row_lse,
                float* __restrict__ col_lse, int Bm, int V, int Bb, int col_blocks) {
  __shared__ float sm[8][33], ss[8][33];
  const int R = Bm * V;
  if (static_cast<int>(blockIdx.x) < col_blocks) {
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int j = blockIdx.x * 32 + x;
    float m = -INFINITY, s = 0.f;
    if (j < Bb) {
      for (int r = y; r < R; r += 8) {
        const float v = S[static_cast<int64_t>(r) * Bb + j];
        const float mn = fmaxf(m, v);
        s = s * __expf(m - mn) + __expf(v - mn);
        m = mn;
      }
    }
    sm[y][x] = m;
    ss[y][x] = s;
    __syncthreads();
    if (y == 0 && j < Bb) {
      float M = -INFINITY;
      for (int k = 0; k < 8; ++k) M = fmaxf(M, sm[k][x]);
      float Ssum = 0.f;
      for (int k = 0; k < 8; ++k) Ssum += (sm[k][x] == -INFINITY) ? 0.f : ss[k][x] * __expf(sm[k][x] - M);
      col_lse[j] = M + logf(Ssum);
    }
  } else {
    const int lane = threadIdx.x & 31;
    const int i = (blockIdx.x - col_blocks) * 8 + (threadIdx.x >> 5);
    if (i >= Bm) return;
    const float* row = S + static_cast<int64_t>(i * V + true_idx[i]) * Bb;
    float m = -INFINITY;
    for (int c = lane; c < Bb; c += 32) m = fmaxf(m, row[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < Bb; c += 32) s += __expf(row[c] - m);
    s = warp_sum(s);
    if (lane == 0) row_lse[i] = m + logf(s);
  }
}

__global__ void __launch_bounds__(256)
clip_loss_finalize_kernel(const float* __restrict__ S, const int32_t* __restrict__ true_idx,
                          const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                          float* __restrict__ loss, int Bm, int V, int Bb) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int i = threadIdx.x; i < Bm; i += blockDim.x) {
    const float diag = S[static_cast<int64_t>(i * V + true_idx[i]) * Bb + i];
    acc += (row_lse[i] - diag) / Bm + (col_lse[i] - diag) / Bb;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int k = 0; k < 8; ++k) t += red[k];
    *loss = 0.5f * t;
  }
}

// Embedding extraction (extract_beatmap_embeddings.py:243-262): window embeddings are summed per beatmap
// (sums[slot[i]] += e[i], counts[slot[i]] += 1) and finalised as mean / |mean| (left as the mean when its norm is 0).
__global__ void __launch_bounds__(256)
segment_accumulate_kernel(const float* __restrict__ e, const int32_t* __restrict__ slot, float* __restrict__ sums,
                          float* __restrict__ counts, int rows, int P) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int s = slot[row];
  if (s < 0) return;
  for (int i = lane; i < P; i += 32) atomicAdd(sums + static_cast<int64_t>(s) * P + i, e[static_cast<int64_t>(row) * P + i]);
  if (lane == 0) atomicAdd(counts + s, 1.f);
}
__global__ void __launch_bounds__(256)
mean_renormalize_kernel(const float* __restrict__ sums, const float* __restrict__ counts, float* __restrict__ out,
                        int rows, int P) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float inv_n = 1.f / fmaxf(counts[row], 1.f);
  float ss = 0.f;
  for (int i = lane; i < P; i += 32) {
    const float m = sums[static_cast<int64_t>(row) * P + i] * inv_n;
    ss += m * m;
  }
  ss = warp_sum(ss);
  const float norm = sqrtf(ss);
  const float sc = norm > 0.f ? inv_n / norm : inv_n;
  for (int i = lane; i < P; i += 32) out[static_cast<int64_t>(row) * P + i] = sums[static_cast<int64_t>(row) * P + i] * sc;
}

}  // namespace

int segment_accumulate(const float* e, const int32_t* slot, float* sums, float* counts, int rows, int P,
                       cudaStream_t stream) {
  if (rows == 0) return kOk;
  segment_accumulate_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(e, slot, sums, counts, rows, P);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}
int mean_renormalize(const float* sums, const float* counts, float* out, int rows, int P, cudaStream_t stream) {
  if (rows == 0) return kOk;
  mean_renormalize_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(sums, counts, out, rows, P);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int layernorm_fwd(const void* x, const float* gamma, void* y, float* stats, int64_t rows, int H, float eps,
                  cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0 && H <= MAXV * 256, kBadShape, "layernorm: hidden size %d must be a multiple of 8 and <= %d",
               H, MAXV * 256);
  if (rows == 0) return kOk;
  constexpr int NR = CM3P_LN_NR;
  layernorm_fwd_kernel<NR><<<static_cast<unsigned>((rows + 8 * NR - 1) / (8 * NR)), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), gamma, reinterpret_cast<__nv_bfloat16*>(y),
      reinterpret_cast<float2*>(stats), rows, H, eps);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int embed_gather_ln(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot, const void* tok_emb,
                    const void* audio_embeds, const float* gamma, void* y, float* stats, int64_t rows, int H,
                    int vocab, float eps, cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0 && H <= MAXV * 256, kBadShape, "embed: hidden size %d must be a multiple of 8 and <= %d", H,
               MAXV * 256);
  if (rows == 0) return kOk;
  embed_gather_ln_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      ids, src_index, audio_slot, reinterpret_cast<const __nv_bfloat16*>(tok_emb),
      reinterpret_cast<const __nv_bfloat16*>(audio_embeds), gamma, reinterpret_cast<__nv_bfloat16*>(y),
      reinterpret_cast<float2*>(stats), rows, H, vocab, eps);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int transpose_cast(const float* x, void* out, int B, int C, int F, cudaStream_t stream) {
  CM3P_REQUIRE(B > 0 && B <= 65535 && C > 0 && F > 0, kBadShape, "transpose_cast: B=%d C=%d F=%d", B, C, F);
  dim3 grid((F + 31) / 32, (C + 31) / 32, B);
  transpose_cast_kernel<<<grid, 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), C, F);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int gather_rows(const void* x, const int32_t* index, void* out, int rows, int H, cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0, kBadShape, "gather_rows: H %% 8 required");
  if (rows == 0) return kOk;
  const int total = rows * (H / 8);
  gather_rows_kernel<<<(total + 255) / 256, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), index,
                                                              reinterpret_cast<__nv_bfloat16*>(out), rows, H);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int mean_pool(const void* x, const int32_t* cu_seqlens, void* out, int batch, int H, cudaStream_t stream) {
  CM3P_REQUIRE(H % 8 == 0, kBadShape, "mean_pool: H %% 8 required");
  if (batch == 0) return kOk;
  dim3 grid(batch, (H + 63) / 64);
  mean_pool_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), cu_seqlens,
                                             reinterpret_cast<__nv_bfloat16*>(out), H);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int l2norm_rows(const float* e, float* out_f32, void* out_bf16, float* inv_norm, int rows, int P,
                cudaStream_t stream) {
  if (rows == 0) return kOk;
  l2norm_rows_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(e, out_f32, reinterpret_cast<__nv_bfloat16*>(out_bf16),
                                                         inv_norm, rows, P);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int clip_loss_fwd(const float* S, const int32_t* true_idx, float* row_lse, float* col_lse, float* loss, int Bm, int V,
                  int Bb, cudaStream_t stream) {
  CM3P_REQUIRE(Bm == Bb, kBadShape, "clip_loss: metadata batch %d != beatmap batch %d", Bm, Bb);
  CM3P_REQUIRE(Bm > 0 && V > 0, kBadShape, "clip_loss: empty batch");
  const int col_blocks = (Bb + 31) / 32;
  const int row_blocks = (Bm + 7) / 8;
  clip_lse_kernel<<<col_blocks + row_blocks, 256, 0, stream>>>(S, true_idx, row_lse, col_lse, Bm, V, Bb, col_blocks);
  CM3P_CUDA_TRY(cudaGetLastError());
  clip_loss_finalize_kernel<<<1, 256, 0, stream>>>(S, true_idx, row_lse, col_lse, loss, Bm, V, Bb);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
