// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> 128B-swizzled smem ring -> tcgen05.mma -> fp32
// accumulators in TMEM (2 stages x 256 columns, so the epilogue of tile i overlaps the MMAs of tile i+1) -> fused
// epilogue.  bf16 outputs are staged per 128x64 slab in 128B-swizzled shared memory and written with TMA stores
// (the residual slab is TMA-loaded into the same buffer one slab ahead), so the LSU never sees row-strided global
// accesses; fp32 / unaligned outputs use the direct register->global path.
//
//   C[M,N] = epilogue( A[M,K] . B[N,K]^T )
//
// Default form (MMA2): CTA pairs (2-CTA clusters, the two SMs of a TPC) on adjacent M tiles run ONE
// tcgen05.mma.cta_group::2 of M = 256, N = 256, K = 16 per step: each CTA holds its 128 rows of A, HALF of the B tile
// and its 128 accumulator rows; the rank-0 CTA issues, both CTAs' TMA loads report to its barrier, its commits free
// the stage / publish the accumulator in both CTAs.  Per CTA and k-block 32 KB instead of 48 KB enter shared memory
// and 8 KB instead of 12 KB are read per MMA step, and the same 192 KB hold 6 stages instead of 4 (measured against
// the previous form - pairs sharing B by TMA multicast, each CTA its own M = 128 MMA: +4 .. +15 % on the train-step
// shapes, profiles/r2_kernel_microbench_mma2.jsonl).  Single CTAs (cluster 1) and the implicit-GEMM convolutions use
// cta_group::1, M = 128.
//
// Replaces every nn.Linear on the hot path (reference: ModernBERT Wqkv/Wo/Wi/Wo, the audio projector
// modeling_cm3p.py:470-481, the projections :959/:971, the logits matmul :976-977, the MLM head
// :1229-1238) and, as an implicit GEMM, the two Conv1d of the audio front-end (:488-489).
//
// Warp roles (256 threads, 1 CTA / SM):
//   warp 0 lane 0 : TMA producer            warp 1 lane 0 : MMA issuer
//   warp 2        : TMEM alloc / dealloc    warps 4..7    : epilogue (TMEM lane quadrant = warp - 4)
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.h"
#include "gemm.h"
#include "ptx.cuh"

namespace cm3p {
namespace {

constexpr int BM = 128;
constexpr int BN = 256;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle row
constexpr int STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
// CTA-pair MMA (cta_group::2, M = 256 over two CTAs): a CTA holds only its half of every B tile, so the same 192 KB would
// give 6 stages of (16 KB A + 16 KB B-half); 5 measure the same (profiles/r2_kernel_microbench_mma2.jsonl), and the
// 32 KB they leave hold two more output slabs: with two, every slab of the residual epilogue waited for the previous
// slab's TMA store to drain before it could prefetch the next residual into that buffer
constexpr int STAGES_2CTA = 5;
constexpr int MAX_STAGES = 6;
constexpr int SLAB_BUFS = 2, SLAB_BUFS_2CTA = 4;
constexpr int MAX_SLAB_BUFS = 4;
constexpr int ACC_STAGES = 2;
constexpr int TMEM_COLS = ACC_STAGES * BN;  // 512
constexpr int SLAB_BYTES = BM * 128;  // 128 rows x 64 bf16
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + SLAB_BUFS * SLAB_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
static_assert(STAGES_2CTA * (A_STAGE_BYTES + B_STAGE_BYTES / 2) + SLAB_BUFS_2CTA * SLAB_BYTES ==
                  STAGES * STAGE_BYTES + SLAB_BUFS * SLAB_BYTES, "both forms use the same shared-memory footprint");
static_assert((2 * MAX_STAGES + 2 * ACC_STAGES + MAX_SLAB_BUFS) * 8 + 4 <= 256, "barrier block");
constexpr int THREADS = 256;
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

struct Params {
  int64_t M, N, K;
  void* c;
  int64_t ldc;
  const void* aux;  // residual [M, N] bf16 (ld_aux) or bias [N] fp32
  int64_t ld_aux;
  void* c2;  // second output (raw pre-activation for GEGLU_SAVE)
  int64_t ldc2;
  float scale;
  const int32_t* positions;  // [M] token position inside its sequence (ROPE)
  const float2* rope_table;  // [max_pos][32] (cos, sin)
  int64_t rope_cols;         // columns [0, rope_cols) are rotated per 64-wide head
  int trans_a, trans_b;
  int accumulate;  // F32 epilogue: C += acc (atomic, so K may be split across CTAs)
  int vec_c;       // C (and C2 / aux) rows allow 16-byte accesses
  int splits;        // split-K factor (>1 only with the atomic F32 epilogue: weight gradients, K = tokens)
  int kb_per_split;  // k-blocks per split
  int cluster;       // 1, or 2: CTA pairs on adjacent M tiles share every B tile through TMA multicast
  int mma2;          // cluster == 2 only: ONE tcgen05.mma.cta_group::2 of M = 256 per pair instead (each CTA loads half of B)
  float* stats_out;        // RESIDUAL: [ceil(N/256)][M] (sum, sum^2) partials of the written rows, one per N tile, or null
  const float* row_stats;  // ROPE / GEGLU(_SAVE): [ceil(K/256)][M] partials of the A rows -> LayerNorm folded in, or null
  const float* col_corr;   // [N] column sums of B (= W . diag(gamma))
  float ln_eps;
  int32_t* tile_sem;  // split-K turnstile, one counter per output tile (zero between launches), or null = fp32 atomics
  int64_t mn_tiles;   // output tiles (per split) as the kernel counts them
  int64_t group_m;    // > 0: grouped GEMM, see GemmArgs::group_m
  // Implicit-GEMM conv1d (k = 3, pad 1): the A (forward) / B (weight gradient) operand is never materialised; its
  // 64-channel x 128-frame boxes are fetched straight from the channels-last input [B, F, C] through a 4-D tensor
  // map (c, frame parity, frame / stride, window), one shifted box per tap; out-of-range frames (the zero padding)
  // are the tensor map's out-of-bounds fill.
  int conv_mode;      // 0 plain GEMM, 1 conv forward, 2 conv weight gradient
  int conv_stride;    // 1 or 2
  int conv_per_win;   // forward: M tiles per window; weight gradient: 64-row K blocks per window
  int conv_kb_per_tap;  // forward: 64-channel K blocks per tap
  int conv_cpad;      // weight gradient: padded input channels per tap (N = 3 * conv_cpad)
};

// input frame (parity r, index u) read by output frame t' for tap j: stride * t' + j - 1
__device__ __forceinline__ void conv_tap_coord(int stride, int tap, int t0, int& r, int& u0) {
  if (stride == 1) {
    r = 0;
    u0 = t0 + tap - 1;
  } else {  // stride 2: frame 2 t' + tap - 1 = 2 (t' - 1) + 1 | 2 t' | 2 t' + 1
    r = (tap == 1) ? 0 : 1;
    u0 = (tap == 0) ? t0 - 1 : t0;
  }
}

__device__ __forceinline__ void red_add_f32x4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float (&v)[32], int64_t col, int64_t ncols,
                                              bool vec = true) {
  // dst points at (row, col); 16-byte stores, guarded per 8 columns
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (vec && col + i * 8 + 8 <= ncols) {
      uint4 u;
      u.x = ptx::pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
      u.y = ptx::pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
      u.z = ptx::pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
      u.w = ptx::pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
      *reinterpret_cast<uint4*>(dst + i * 8) = u;
    } else {
      for (int j = 0; j < 8; ++j)
        if (col + i * 8 + j < ncols) dst[i * 8 + j] = __float2bfloat16(v[i * 8 + j]);
    }
  }
}

template <int EPI>
__device__ __forceinline__ void epilogue_tile(const Params& p, uint32_t tmem_acc, int quad, int64_t m0, int64_t n0) {
  const int lane = ptx::lane_id();
  const int64_t row = m0 + quad * 32 + lane;
  const bool row_ok = row < p.M;
  const uint32_t taddr_row = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);

  if constexpr (EPI == EPI_ROPE) {
    // one 64-wide head per iteration: x1 = cols [0,32), x2 = cols [32,64)
    // (cos, sin) of this row's position: loaded once per tile, reused by the tile's 4 heads
    float2 cs[32];
    if (n0 < p.rope_cols) {
      const int pos = row_ok ? p.positions[row] : 0;
      const float4* tab = reinterpret_cast<const float4*>(p.rope_table + static_cast<int64_t>(pos) * 32);
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 f = __ldg(tab + k);
        cs[2 * k] = make_float2(f.x, f.y);
        cs[2 * k + 1] = make_float2(f.z, f.w);
      }
    }
#pragma unroll 1
    for (int hcol = 0; hcol < BN; hcol += 64) {
      const int64_t col = n0 + hcol;
      if (col >= p.N) break;
      uint32_t r1[32], r2[32];
      ptx::tmem_ld_32x32b_x32(taddr_row + hcol, r1);
      ptx::tmem_ld_32x32b_x32(taddr_row + hcol + 32, r2);
      ptx::tmem_ld_wait();
      float o1[32], o2[32];
      if (col < p.rope_cols) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const float x1 = __uint_as_float(r1[k]), x2 = __uint_as_float(r2[k]);
          o1[k] = x1 * cs[k].x - x2 * cs[k].y;
          o2[k] = x2 * cs[k].x + x1 * cs[k].y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          o1[k] = __uint_as_float(r1[k]);
          o2[k] = __uint_as_float(r2[k]);
        }
      }
      if (row_ok) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col;
        store_bf16x32(dst, o1, col, p.N, p.vec_c);
        store_bf16x32(dst + 32, o2, col + 32, p.N, p.vec_c);
      }
    }
    return;
  }

  float sc = p.scale;
  if constexpr (EPI == EPI_SCALE_F32) {
    // optional device-side factor exp(*aux): the logits' `* logit_scale.exp()` without a host read of the parameter
    if (p.aux) sc *= expf(__ldg(reinterpret_cast<const float*>(p.aux)));
  }
#pragma unroll 1
  for (int c = 0; c < BN; c += 32) {
    const int64_t col = n0 + c;
    if (col >= p.N) break;  // warp-uniform
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(taddr_row + c, r);
    ptx::tmem_ld_wait();
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);

    if constexpr (EPI == EPI_STORE) {
      if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, v, col, p.N, p.vec_c);
    } else if constexpr (EPI == EPI_RESIDUAL) {
      if (row_ok) {
        const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(p.aux) + row * p.ld_aux + col;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (p.vec_c && col + i * 8 + 8 <= p.N) {
            const uint4 u = *reinterpret_cast<const uint4*>(res + i * 8);
            float2 f;
            f = ptx::unpack_bf16x2(u.x); v[i * 8 + 0] += f.x; v[i * 8 + 1] += f.y;
            f = ptx::unpack_bf16x2(u.y); v[i * 8 + 2] += f.x; v[i * 8 + 3] += f.y;
            f = ptx::unpack_bf16x2(u.z); v[i * 8 + 4] += f.x; v[i * 8 + 5] += f.y;
            f = ptx::unpack_bf16x2(u.w); v[i * 8 + 6] += f.x; v[i * 8 + 7] += f.y;
          } else {
            for (int j = 0; j < 8; ++j)
              if (col + i * 8 + j < p.N) v[i * 8 + j] += __bfloat162float(res[i * 8 + j]);
          }
        }
        store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, v, col, p.N, p.vec_c);
      }
    } else if constexpr (EPI == EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = ptx::gelu_erf(v[i]);
      if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, v, col, p.N, p.vec_c);
    } else if constexpr (EPI == EPI_BIAS_GELU || EPI == EPI_BIAS) {
      const float* bias = reinterpret_cast<const float*>(p.aux);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float b = (col + i < p.N) ? __ldg(bias + col + i) : 0.f;
        v[i] = (EPI == EPI_BIAS_GELU) ? ptx::gelu_erf(v[i] + b) : v[i] + b;
      }
      if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + col, v, col, p.N, p.vec_c);
    } else if constexpr (EPI == EPI_GEGLU || EPI == EPI_GEGLU_SAVE) {
      // Wi rows are interleaved in groups of 16 (u0..u15, g0..g15, u16.., g16..): see host prep.
      if constexpr (EPI == EPI_GEGLU_SAVE) {
        if (row_ok) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.c2) + row * p.ldc2 + col, v, col, p.N, p.vec_c);
      }
      float o[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = ptx::gelu_erf(v[i]) * v[16 + i];
      if (row_ok) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + (col >> 1);
        uint4 u0, u1;
        u0.x = ptx::pack_bf16x2(o[0], o[1]);   u0.y = ptx::pack_bf16x2(o[2], o[3]);
        u0.z = ptx::pack_bf16x2(o[4], o[5]);   u0.w = ptx::pack_bf16x2(o[6], o[7]);
        u1.x = ptx::pack_bf16x2(o[8], o[9]);   u1.y = ptx::pack_bf16x2(o[10], o[11]);
        u1.z = ptx::pack_bf16x2(o[12], o[13]); u1.w = ptx::pack_bf16x2(o[14], o[15]);
        *reinterpret_cast<uint4*>(dst) = u0;
        *reinterpret_cast<uint4*>(dst + 8) = u1;
      }
    } else if constexpr (EPI == EPI_SCALE_F32) {
      if (row_ok) {
        float* dst = reinterpret_cast<float*>(p.c) + row * p.ldc + col;
        if (p.accumulate == 2 && p.vec_c && col + 32 <= p.N) {
          // ordered accumulation (this CTA owns the tile until it passes the turnstile on): all eight 16-byte loads
          // of the row chunk are in flight together, ONE memory round trip per chunk instead of eight dependent ones
          float4 cur[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) cur[i] = __ldcg(reinterpret_cast<const float4*>(dst + i * 4));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            cur[i].x += v[i * 4] * sc; cur[i].y += v[i * 4 + 1] * sc;
            cur[i].z += v[i * 4 + 2] * sc; cur[i].w += v[i * 4 + 3] * sc;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) __stcg(reinterpret_cast<float4*>(dst + i * 4), cur[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (p.vec_c && col + i * 4 + 4 <= p.N) {
              if (p.accumulate == 2) {
                float4 cur = __ldcg(reinterpret_cast<const float4*>(dst + i * 4));
                cur.x += v[i * 4] * sc; cur.y += v[i * 4 + 1] * sc;
                cur.z += v[i * 4 + 2] * sc; cur.w += v[i * 4 + 3] * sc;
                __stcg(reinterpret_cast<float4*>(dst + i * 4), cur);
              } else if (p.accumulate) {
                red_add_f32x4(dst + i * 4, v[i * 4] * sc, v[i * 4 + 1] * sc, v[i * 4 + 2] * sc, v[i * 4 + 3] * sc);
              } else {
                *reinterpret_cast<float4*>(dst + i * 4) = make_float4(v[i * 4] * sc, v[i * 4 + 1] * sc,
                                                                      v[i * 4 + 2] * sc, v[i * 4 + 3] * sc);
              }
            } else {
              for (int j = 0; j < 4; ++j)
                if (col + i * 4 + j < p.N) {
                  if (p.accumulate == 2)
                    __stcg(dst + i * 4 + j, __ldcg(dst + i * 4 + j) + v[i * 4 + j] * sc);
                  else if (p.accumulate)
                    atomicAdd(dst + i * 4 + j, v[i * 4 + j] * sc);
                  else
                    dst[i * 4 + j] = v[i * 4 + j] * sc;
                }
            }
          }
        }
      }
    }
  }
}

// Split-K turnstile: the splits of one output tile add their partial sums into C in split order, so weight
// gradients are bit-reproducible from run to run (fp32 atomics are not: the order in which CTAs arrive changes the
// rounding).  Work units are numbered split-major and every CTA takes its units in increasing order, so split s of
// a tile starts after split s-1 has started: the wait is short and cannot deadlock (the lowest unfinished unit
// never waits for an unfinished one).
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Staged epilogue (bf16 outputs with 16-byte aligned rows): per 128x64 output slab
//   TMEM -> registers -> (op) -> 128B-swizzled smem slab -> TMA store.
// For EPI_RESIDUAL the residual slab is TMA-loaded into the same smem buffer one slab ahead and the
// sum is written back in place.  Two slab buffers alternate; `leader` = first epilogue thread.
struct EpiState {
  int g = 0;                  // processed-slab counter (selects the buffer)
  uint32_t res_parity = 0;    // bit b = parity of the next residual load into buffer b
  int pair = 0;               // GeGLU+save: product slabs stored so far (selects the product buffer)
};

template <int EPI>
__device__ __forceinline__ constexpr int slab_acc_cols() {
  return (EPI == EPI_GEGLU) ? 128 : 64;  // accumulator columns consumed per 64-column output slab
}

__device__ __forceinline__ void slab_write_row(uint8_t* slab, int r, int chunk, const uint4& v) {
  *reinterpret_cast<uint4*>(slab + r * 128 + ((chunk ^ (r & 7)) << 4)) = v;
}
__device__ __forceinline__ uint4 slab_read_row(const uint8_t* slab, int r, int chunk) {
  return *reinterpret_cast<const uint4*>(slab + r * 128 + ((chunk ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint4 pack8f(const float* v) {
  return make_uint4(ptx::pack_bf16x2(v[0], v[1]), ptx::pack_bf16x2(v[2], v[3]), ptx::pack_bf16x2(v[4], v[5]),
                    ptx::pack_bf16x2(v[6], v[7]));
}

template <int EPI, int NBUF>
__device__ __forceinline__ void epilogue_tile_staged(const Params& p, const CUtensorMap* tma_c,
                                                     const CUtensorMap* tma_c2, const CUtensorMap* tma_aux,
                                                     uint8_t* slabs, uint64_t* res_full, EpiState& st,
                                                     uint32_t tmem_acc, int quad, int64_t m0, int64_t n0,
                                                     int64_t next_m0, int64_t next_n0, bool has_next_tile,
                                                     const float2 (&cs)[32]) {
  constexpr int ACC = slab_acc_cols<EPI>();
  constexpr int NSLAB = BN / ACC;
  const int lane = ptx::lane_id();
  const int r = quad * 32 + lane;  // row inside the tile == TMEM lane
  const int64_t row = m0 + r;
  const bool leader = (quad == 0 && lane == 0);
  const uint32_t taddr_row = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16);

  // LayerNorm folded into this GEMM (see GemmArgs): y = rstd * (x.W'^T - mean * colsum(W'))
  float ln_mean = 0.f, ln_rstd = 1.f;
  bool ln = false;
  if constexpr (EPI == EPI_ROPE || EPI == EPI_GEGLU || EPI == EPI_GEGLU_SAVE) {
    ln = p.row_stats != nullptr;
    if (ln && row < p.M) {
      // the producer wrote one (sum, sum^2) partial per 256-column tile of this row; summed in tile order, so the
      // statistics (and everything downstream) do not depend on which CTA finished first
      float2 sq = make_float2(0.f, 0.f);
      for (int64_t part = 0; part < (p.K + BN - 1) / BN; ++part) {
        const float2 t = *reinterpret_cast<const float2*>(p.row_stats + (part * p.M + row) * 2);
        sq.x += t.x;
        sq.y += t.y;
      }
      const float inv_w = 1.f / static_cast<float>(p.K);
      ln_mean = sq.x * inv_w;
      ln_rstd = rsqrtf(fmaxf(sq.y * inv_w - ln_mean * ln_mean, 0.f) + p.ln_eps);
    }
  }
  auto ln_fix = [&](auto& v, int64_t first_col) {  // all columns < N (N % 64 == 0 is required)
    constexpr int COUNT = sizeof(v) / sizeof(float);
#pragma unroll
    for (int i = 0; i < COUNT; i += 4) {
      const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.col_corr + first_col + i));
      v[i] = ln_rstd * (v[i] - ln_mean * c4.x);
      v[i + 1] = ln_rstd * (v[i + 1] - ln_mean * c4.y);
      v[i + 2] = ln_rstd * (v[i + 2] - ln_mean * c4.z);
      v[i + 3] = ln_rstd * (v[i + 3] - ln_mean * c4.w);
    }
  };
  float st_sum = 0.f, st_sq = 0.f;  // RESIDUAL: statistics of the rows written by this thread

#pragma unroll 1
  for (int sl = 0; sl < NSLAB; ++sl) {
    const int64_t col = n0 + sl * ACC;  // first accumulator column of the slab
    if (col >= p.N) break;              // CTA-uniform
    // NBUF slab buffers in rotation: slab g is assembled in buffer g % NBUF, whose previous TMA store (slab g - NBUF)
    // must have left shared memory: at most NBUF - 1 younger stores may still be reading.  The residual of slab g + 1
    // is prefetched one slab ahead into ITS buffer, last stored by slab g + 1 - NBUF: at most NBUF - 2 pending.
    // GeGLU+save with 4 buffers: buffers 0, 1 rotate for the raw slabs; buffers 2, 3 collect the PRODUCT of two
    // consecutive raw slabs (2 x 32 columns = one 128-byte row per token) and go out with one TMA store in the bulk
    // group of the second slab, so a bulk group still is "one slab" for the wait counts.  (Writing the product with
    // one 64-byte row segment per lane cost 32 L1 lines per store instruction.)
    constexpr bool PSTAGE = (EPI == EPI_GEGLU_SAVE) && NBUF == 4;
    constexpr int RBUF = PSTAGE ? 2 : NBUF;  // buffers in rotation for the staged slab itself
    const int buf = st.g % RBUF, nbuf = (st.g + 1) % RBUF;
    uint8_t* slab = slabs + buf * SLAB_BYTES;
    uint8_t* pslab = slabs + (2 + ((st.pair + (sl >> 1)) & 1)) * SLAB_BYTES;  // PSTAGE only
    const int64_t out_col = (EPI == EPI_GEGLU) ? (col >> 1) : col;

    // ---- free the buffer(s); prefetch the next residual slab
    // With 4 buffers in rotation the leader waits one store EARLIER than this slab needs (all but the latest 2 have left
    // shared memory: that frees the buffer of the NEXT slab), and everybody learns it at the barrier before this
    // slab's TMA store - so nobody has to wait for the leader at the top of a slab (ncu: 10 % of the residual GEMM's
    // stall samples sat at that barrier).  With 2 buffers the barrier stays.
    constexpr bool EARLY_FREE = RBUF >= 4;
    if (leader) {
      if constexpr (EPI == EPI_RESIDUAL) {
        ptx::tma_store_wait_read<RBUF - 2>();
        int64_t ncol = col + ACC, nm0 = m0;
        bool have = (sl + 1 < NSLAB) && (ncol < p.N);
        if (!have && has_next_tile) { have = true; ncol = next_n0; nm0 = next_m0; }
        if (have) {
          ptx::mbar_arrive_expect_tx(&res_full[nbuf], SLAB_BYTES);
          ptx::tma_load_2d(slabs + nbuf * SLAB_BYTES, tma_aux, &res_full[nbuf], static_cast<int32_t>(ncol),
                           static_cast<int32_t>(nm0));
        }
      } else {
        ptx::tma_store_wait_read<EARLY_FREE ? RBUF - 2 : RBUF - 1>();
      }
    }
    if constexpr (!EARLY_FREE) ptx::named_bar_sync(1, 128);

    if constexpr (EPI == EPI_GEGLU) {
      // 128 accumulator columns (interleaved 16 u / 16 g) -> 64 output columns
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t ra[32], rb[32];
        ptx::tmem_ld_32x32b_x32(taddr_row + sl * ACC + h * 64, ra);
        ptx::tmem_ld_32x32b_x32(taddr_row + sl * ACC + h * 64 + 32, rb);
        ptx::tmem_ld_wait();
        float fa[32], fb[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          fa[i] = __uint_as_float(ra[i]);
          fb[i] = __uint_as_float(rb[i]);
        }
        if (ln) {
          ln_fix(fa, col + h * 64);
          ln_fix(fb, col + h * 64 + 32);
        }
        float o[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          o[i] = ptx::gelu_erf(fa[i]) * fa[16 + i];
          o[16 + i] = ptx::gelu_erf(fb[i]) * fb[16 + i];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) slab_write_row(slab, r, h * 4 + c, pack8f(o + c * 8));
      }
    } else {
      uint32_t r1[32], r2[32];
      ptx::tmem_ld_32x32b_x32(taddr_row + sl * ACC, r1);
      ptx::tmem_ld_32x32b_x32(taddr_row + sl * ACC + 32, r2);
      ptx::tmem_ld_wait();
      float v[64];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = __uint_as_float(r1[i]);
        v[32 + i] = __uint_as_float(r2[i]);
      }
      if constexpr (EPI == EPI_ROPE || EPI == EPI_GEGLU_SAVE) {
        if (ln) ln_fix(v, col);
      }
      if constexpr (EPI == EPI_RESIDUAL) {
        ptx::mbar_wait(&res_full[buf], (st.res_parity >> buf) & 1u);
        st.res_parity ^= 1u << buf;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 u = slab_read_row(slab, r, c);
          float2 f;
          f = ptx::unpack_bf16x2(u.x); v[c * 8 + 0] += f.x; v[c * 8 + 1] += f.y;
          f = ptx::unpack_bf16x2(u.y); v[c * 8 + 2] += f.x; v[c * 8 + 3] += f.y;
          f = ptx::unpack_bf16x2(u.z); v[c * 8 + 4] += f.x; v[c * 8 + 5] += f.y;
          f = ptx::unpack_bf16x2(u.w); v[c * 8 + 6] += f.x; v[c * 8 + 7] += f.y;
        }
        if (p.stats_out) {
          // statistics of what is actually stored (bf16-rounded), for the LayerNorm folded into the next GEMM
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            if (col + i < p.N) {
              const float rv = __bfloat162float(__float2bfloat16(v[i]));
              st_sum += rv;
              st_sq += rv * rv;
            }
          }
        }
      } else if constexpr (EPI == EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 64; ++i) v[i] = ptx::gelu_erf(v[i]);
      } else if constexpr (EPI == EPI_BIAS_GELU || EPI == EPI_BIAS) {
        const float* bias = reinterpret_cast<const float*>(p.aux);
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float b = (col + i < p.N) ? __ldg(bias + col + i) : 0.f;
          v[i] = (EPI == EPI_BIAS_GELU) ? ptx::gelu_erf(v[i] + b) : v[i] + b;
        }
      } else if constexpr (EPI == EPI_ROPE) {
        if (col < p.rope_cols) {
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float x1 = v[k], x2 = v[32 + k];
            v[k] = x1 * cs[k].x - x2 * cs[k].y;
            v[32 + k] = x2 * cs[k].x + x1 * cs[k].y;
          }
        }
      } else if constexpr (EPI == EPI_GEGLU_SAVE) {
        float o[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          o[i] = ptx::gelu_erf(v[i]) * v[16 + i];
          o[16 + i] = ptx::gelu_erf(v[32 + i]) * v[48 + i];
        }
        if constexpr (PSTAGE) {
          // half of a product row: 16-byte units 0..3 (even slab of the pair) or 4..7 (odd slab)
#pragma unroll
          for (int c = 0; c < 4; ++c) slab_write_row(pslab, r, (sl & 1) * 4 + c, pack8f(o + c * 8));
        } else if (row < p.M) {
          // product written directly (1/3 of the bytes); the raw slab goes through the staged path
          __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.c) + row * p.ldc + (col >> 1);
#pragma unroll
          for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(dst + c * 8) = pack8f(o + c * 8);
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) slab_write_row(slab, r, c, pack8f(v + c * 8));
    }

    ptx::fence_proxy_async_smem();
    ptx::named_bar_sync(2, 128);
    if (leader) {
      if (p.conv_mode == 1) {  // output [B, frames / stride, C_out] through a 3-D map: rows past the window are clipped
        const int32_t tile_idx = static_cast<int32_t>(m0 / BM);
        ptx::tma_store_3d(tma_c, slab, static_cast<int32_t>(out_col), (tile_idx % p.conv_per_win) * BM,
                          tile_idx / p.conv_per_win);
      } else {
        ptx::tma_store_2d((EPI == EPI_GEGLU_SAVE) ? tma_c2 : tma_c, slab, static_cast<int32_t>(out_col),
                          static_cast<int32_t>(m0));
        if constexpr (PSTAGE) {
          // the pair's product slab is complete after its odd slab (or at the end of the tile / of N: the unwritten
          // half then lies beyond column N / 2 and is clipped by the tensor map)
          const bool last = (sl + 1 == NSLAB) || (col + ACC >= p.N);
          if ((sl & 1) || last)
            ptx::tma_store_2d(tma_c, pslab, static_cast<int32_t>((col - (sl & 1) * ACC) >> 1), static_cast<int32_t>(m0));
        }
      }
      ptx::tma_store_commit();
    }
    ++st.g;
  }
  if constexpr (EPI == EPI_GEGLU_SAVE) {
    const int64_t cols = (p.N - n0 < BN) ? p.N - n0 : BN;
    st.pair += static_cast<int>((cols + 2 * ACC - 1) / (2 * ACC));  // product slabs of this tile
  }
  if constexpr (EPI == EPI_RESIDUAL) {
    if (p.stats_out && row < p.M)  // this tile's partial: slot [n0 / BN][row]
      *reinterpret_cast<float2*>(p.stats_out + ((n0 / BN) * p.M + row) * 2) = make_float2(st_sum, st_sq);
  }
}

template <int EPI, bool STAGED, bool MMA2>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf16_sm100_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                       const __grid_constant__ CUtensorMap tma_c, const __grid_constant__ CUtensorMap tma_c2,
                       const __grid_constant__ CUtensorMap tma_aux, const __grid_constant__ CUtensorMap tma_bh,
                       const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  // MMA2 is a template parameter: a kernel that contains cta_group::2 instructions can only be launched as CTA pairs
  constexpr bool mma2 = MMA2;
  constexpr int nstages = mma2 ? STAGES_2CTA : STAGES;
  constexpr int b_stage_bytes = mma2 ? B_STAGE_BYTES / 2 : B_STAGE_BYTES;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + nstages * A_STAGE_BYTES;
  constexpr int nslab_bufs = mma2 ? SLAB_BUFS_2CTA : SLAB_BUFS;
  uint8_t* slabs = smem + nstages * (A_STAGE_BYTES + b_stage_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(slabs + nslab_bufs * SLAB_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tmem_full = bars + 2 * MAX_STAGES;
  uint64_t* tmem_empty = bars + 2 * MAX_STAGES + ACC_STAGES;
  uint64_t* res_full = bars + 2 * MAX_STAGES + 2 * ACC_STAGES;  // [nslab_bufs]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 * ACC_STAGES + MAX_SLAB_BUFS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tma_a);
    ptx::prefetch_tmap(&tma_b);
  }
  // cluster mode: the pair's two CTAs each load half of every B tile and multicast it to both, so a
  // stage may only be refilled once BOTH tensor pipes have drained it (empty barriers count 2 arrivals)
  const uint32_t crank = p.cluster > 1 ? ptx::cluster_ctarank() : 0u;
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < nstages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], mma2 ? 1 : p.cluster);  // pair MMA: one commit (the leader's) frees the stage in both CTAs
    }
    for (int s = 0; s < ACC_STAGES; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      // one arrive per epilogue warp; pair MMA: the leader's issuer waits for the epilogue warps of BOTH CTAs
      ptx::mbar_init(&tmem_empty[s], mma2 ? 8 : 4);
    }
    for (int s = 0; s < nslab_bufs; ++s) ptx::mbar_init(&res_full[s], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (mma2) {
      ptx::tmem_alloc_2cta(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish_2cta();
    } else {
      ptx::tmem_alloc(tmem_slot, TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if (p.cluster > 1) ptx::cluster_sync_all();  // barrier inits visible to the peer before any remote arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t tiles_m = (p.M + BM - 1) / BM;
  const int64_t tiles_n = (p.N + BN - 1) / BN;
  // work units: (group of `cluster` adjacent M tiles, N tile, K split); CTA `crank` of the cluster takes M tile
  // group*cluster + crank (possibly past the end of M: it then only serves its half of the B loads)
  const int64_t mn_tiles = ((tiles_m + p.cluster - 1) / p.cluster) * tiles_n;  // counted in cluster units
  const int64_t num_tiles = mn_tiles * p.splits;
  const int64_t unit0 = blockIdx.x / p.cluster, unit_step = gridDim.x / p.cluster;
  const int num_kb_total = static_cast<int>((p.K + BK - 1) / BK);
  auto tile_m0 = [&](int64_t t) { return (((t % mn_tiles) / tiles_n) * p.cluster + crank) * BM; };
  auto tile_n0 = [&](int64_t t) { return ((t % mn_tiles) % tiles_n) * BN; };

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------------ TMA producer
    int s = 0;
    uint32_t ph = 0;
    // (An L2 prefetch cursor running 8 k-blocks ahead of the loads was tried here and measured 40% SLOWER:
    //  cp.async.bulk.prefetch.tensor competes with the real loads for the same TMA issue slot.)
    for (int64_t t = unit0; t < num_tiles; t += unit_step) {
      const int32_t m0 = static_cast<int32_t>(tile_m0(t));
      const int32_t n0 = static_cast<int32_t>(tile_n0(t));
      const int kb0 = static_cast<int>(t / mn_tiles) * p.kb_per_split;
      const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
      // grouped GEMM: operands of group g = m0 / group_m are stacked along their outer dimension
      int32_t a_m = m0, a_koff = 0, b_noff = 0, b_koff = 0;
      if (p.group_m > 0) {
        const int32_t grp = static_cast<int32_t>(m0 / p.group_m);
        if (p.trans_a) { a_m = m0 - grp * static_cast<int32_t>(p.group_m); a_koff = grp * static_cast<int32_t>(p.K); }
        if (p.trans_b) b_koff = grp * static_cast<int32_t>(p.K);
        else b_noff = grp * static_cast<int32_t>(p.N);
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* sa = smem_a + s * A_STAGE_BYTES;
        uint8_t* sb = smem_b + s * b_stage_bytes;
        if constexpr (mma2) {
          // this CTA's 128 rows of A and its half of the B tile; the bytes of both CTAs are counted by the leader's barrier
          if (crank == 0) ptx::mbar_arrive_expect_tx(&full[s], 2 * (A_STAGE_BYTES + B_STAGE_BYTES / 2));
          if (!p.trans_a) {
            ptx::tma_load_2d_2cta(sa, &tma_a, &full[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i)
              ptx::tma_load_2d_2cta(sa + i * (BK * 128), &tma_a, &full[s], a_m + i * 64, kb * BK + a_koff);
          }
          if (!p.trans_b) {
            ptx::tma_load_2d_2cta(sb, &tma_bh, &full[s], kb * BK, n0 + b_noff + static_cast<int32_t>(crank) * (BN / 2));
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i)
              ptx::tma_load_2d_2cta(sb + i * (BK * 128), &tma_b, &full[s],
                                    n0 + (static_cast<int>(crank) * (BN / 128) + i) * 64, kb * BK + b_koff);
          }
          if (++s == nstages) { s = 0; ph ^= 1; }
          continue;
        }
        ptx::mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        if (p.conv_mode == 1) {
          // A rows = output frames t0 .. t0+127 of window b; K block = 64 input channels of one tap
          const int tile_idx = m0 / BM, b = tile_idx / p.conv_per_win, t0 = (tile_idx % p.conv_per_win) * BM;
          const int tap = kb / p.conv_kb_per_tap, c0 = (kb % p.conv_kb_per_tap) * BK;
          int r, u0;
          conv_tap_coord(p.conv_stride, tap, t0, r, u0);
          ptx::tma_load_4d(sa, &tma_a, &full[s], c0, r, u0, b);  // box {64 c, 1, 128 frames, 1}
        } else if (p.conv_mode == 2) {
          // K block = 64 output frames of window b; A = dz^T (3-D map: rows past the window are zero-filled)
          const int b = kb / p.conv_per_win, kq = (kb % p.conv_per_win) * BK;
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)
            ptx::tma_load_3d(sa + i * (BK * 128), &tma_a, &full[s], m0 + i * 64, kq, b);
#pragma unroll
          for (int i = 0; i < BN / 64; ++i) {  // B = the input shifted by the tap this 64-column chunk belongs to
            const int col = n0 + i * 64;
            const int tap = min(col / p.conv_cpad, 2), c0 = col - tap * p.conv_cpad;  // c0 >= C: zero fill
            int r, u0;
            conv_tap_coord(p.conv_stride, tap, kq, r, u0);
            ptx::tma_load_4d(sb + i * (BK * 128), &tma_b, &full[s], c0, r, u0, b);  // box {64 c, 1, 64 frames, 1}
          }
        } else if (!p.trans_a) {
          ptx::tma_load_2d(sa, &tma_a, &full[s], kb * BK, m0);  // box {64 k, 128 m}
        } else {
#pragma unroll
          for (int i = 0; i < BM / 64; ++i)  // box {64 m, 64 k} per 64-wide M chunk
            ptx::tma_load_2d(sa + i * (BK * 128), &tma_a, &full[s], a_m + i * 64, kb * BK + a_koff);
        }
        if (p.conv_mode == 2) {
          // B already loaded above
        } else if (p.cluster == 1) {
          if (!p.trans_b) {
            ptx::tma_load_2d(sb, &tma_b, &full[s], kb * BK, n0 + b_noff);  // box {64 k, 256 n}
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i)  // box {64 n, 64 k}
              ptx::tma_load_2d(sb + i * (BK * 128), &tma_b, &full[s], n0 + i * 64, kb * BK + b_koff);
          }
        } else {
          // this CTA's half of the B tile, delivered to both CTAs of the pair
          if (!p.trans_b) {
            ptx::tma_load_2d_multicast(sb + crank * (B_STAGE_BYTES / 2), &tma_bh, &full[s], kb * BK,
                                       n0 + b_noff + static_cast<int32_t>(crank) * (BN / 2), 0x3);  // box {64 k, 128 n}
          } else {
#pragma unroll
            for (int i = 0; i < BN / 128; ++i) {
              const int j = static_cast<int>(crank) * (BN / 128) + i;
              ptx::tma_load_2d_multicast(sb + j * (BK * 128), &tma_b, &full[s], n0 + j * 64, kb * BK + b_koff, 0x3);
            }
          }
        }
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (warp-uniform control flow keeps the descriptor arithmetic in the uniform
    // datapath); one elected lane issues the MMAs and commits.
    // Pair MMA: only the rank-0 CTA issues (M = 256: its own 128 rows and the peer's); its commits arrive in both CTAs.
    const bool leader = ptx::elect_one() && !(mma2 && crank != 0);
    const uint32_t idesc = ptx::umma_idesc_bf16(mma2 ? 2 * BM : BM, BN, p.trans_a ? 1u : 0u, p.trans_b ? 1u : 0u);
    // K-major: 8-row groups 1024 B apart; one UMMA_K (16 elem) step = +32 B inside the swizzle row.
    // MN-major: 8-k groups 1024 B apart (SBO), 64-wide MN chunks BK*128 B apart (LBO); UMMA_K step = +2048 B.
    const uint32_t a_lbo = p.trans_a ? BK * 128 : 16, b_lbo = p.trans_b ? BK * 128 : 16;
    const uint32_t a_kstep = p.trans_a ? 2048 : 32, b_kstep = p.trans_b ? 2048 : 32;
    int s = 0;
    uint32_t ph = 0;
    int as = 0;
    uint32_t aph = 0;
    const bool issues = !(mma2 && crank != 0);  // the peer CTA of a pair MMA has nothing to issue
    for (int64_t t = unit0; issues && t < num_tiles; t += unit_step) {
      ptx::mbar_wait(&tmem_empty[as], aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      const int kb0 = static_cast<int>(t / mn_tiles) * p.kb_per_split;
      const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
      for (int kb = kb0; kb < kb1; ++kb) {
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(smem_a + s * A_STAGE_BYTES);
        const uint32_t b_addr = ptx::smem_u32(smem_b + s * b_stage_bytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + k * a_kstep, a_lbo, 1024);
          const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + k * b_kstep, b_lbo, 1024);
          if (leader) {
            if constexpr (mma2) ptx::umma_bf16_2cta(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            else ptx::umma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
        }
        // frees the smem stage once these MMAs have read it (in both CTAs of a pair: either may refill it)
        if (leader) {
          if constexpr (mma2) {
            ptx::umma_commit_2cta(&empty[s], 0x3);
            if (kb == kb1 - 1) ptx::umma_commit_2cta(&tmem_full[as], 0x3);
          } else {
            if (p.cluster > 1) ptx::umma_commit_multicast(&empty[s], 0x3);
            else ptx::umma_commit(&empty[s]);
            if (kb == kb1 - 1) ptx::umma_commit(&tmem_full[as]);
          }
        }
        if (++s == nstages) { s = 0; ph ^= 1; }
      }
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ epilogue
    const int quad = warp - 4;
    int as = 0;
    uint32_t aph = 0;
    EpiState st;
    if constexpr (STAGED && EPI == EPI_RESIDUAL) {
      // residual slab of the very first output slab
      if (quad == 0 && lane == 0 && unit0 < num_tiles) {
        ptx::mbar_arrive_expect_tx(&res_full[0], SLAB_BYTES);
        ptx::tma_load_2d(slabs, &tma_aux, &res_full[0], static_cast<int32_t>(tile_n0(unit0)),
                         static_cast<int32_t>(tile_m0(unit0)));
      }
    }
    for (int64_t t = unit0; t < num_tiles; t += unit_step) {
      const int64_t m0 = tile_m0(t);
      const int64_t n0 = tile_n0(t);
      [[maybe_unused]] float2 cs[32];
      if constexpr (STAGED && EPI == EPI_ROPE) {
        // (cos, sin) of this row's position for the tile's heads: two dependent global loads (position -> table row),
        // issued BEFORE the wait for the accumulator so that they travel under the tile's MMAs, not in front of the
        // epilogue
        if (n0 < p.rope_cols) {
          const int64_t row = m0 + quad * 32 + lane;
          const int pos = (row < p.M) ? p.positions[row] : 0;
          const float4* tab = reinterpret_cast<const float4*>(p.rope_table + static_cast<int64_t>(pos) * 32);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float4 f = __ldg(tab + k);
            cs[2 * k] = make_float2(f.x, f.y);
            cs[2 * k + 1] = make_float2(f.z, f.w);
          }
        }
      }
      ptx::mbar_wait(&tmem_full[as], aph);
      ptx::tc_fence_after();
      if constexpr (STAGED) {
        const int64_t tn = t + unit_step;
        epilogue_tile_staged<EPI, nslab_bufs>(p, &tma_c, &tma_c2, &tma_aux, slabs, res_full, st, tmem_base + as * BN, quad, m0,
                                  n0, tile_m0(tn), tile_n0(tn), tn < num_tiles, cs);
      } else {
        if constexpr (EPI == EPI_SCALE_F32) {
          if (p.accumulate == 2) {
            // wait until the previous split of this CTA's output tile has added its partial sums
            int32_t* sem = p.tile_sem + (t % mn_tiles) * p.cluster + crank;
            const int32_t split = static_cast<int32_t>(t / mn_tiles);
            if (quad == 0 && lane == 0) {
              const long long t_wait = clock64();
              while (ld_acquire_gpu(sem) != split) {
                __nanosleep(64);
                if (clock64() - t_wait > CM3P_MBAR_TIMEOUT_CYCLES) {  // stale counters: never hang the GPU
                  printf("cm3p_b200: split-K turnstile timed out (tile %lld split %d)\n", (long long)(t % mn_tiles), split);
                  __trap();
                }
              }
            }
            ptx::named_bar_sync(1, 128);
            epilogue_tile<EPI>(p, tmem_base + as * BN, quad, m0, n0);
            __threadfence();
            ptx::named_bar_sync(2, 128);
            if (quad == 0 && lane == 0) st_release_gpu(sem, split + 1 == p.splits ? 0 : split + 1);
          } else {
            epilogue_tile<EPI>(p, tmem_base + as * BN, quad, m0, n0);
          }
        } else {
          epilogue_tile<EPI>(p, tmem_base + as * BN, quad, m0, n0);
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (mma2) ptx::mbar_arrive_leader(&tmem_empty[as]);
        else ptx::mbar_arrive(&tmem_empty[as]);
      }
      if (++as == ACC_STAGES) { as = 0; aph ^= 1; }
    }
    if constexpr (STAGED) {
      if (quad == 0 && lane == 0) ptx::tma_store_wait_all<0>();
    }
  }

  ptx::tc_fence_before();
  if (p.cluster > 1) ptx::cluster_sync_all();  // the peer may still multicast into / signal this CTA
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if constexpr (mma2) ptx::tmem_dealloc_2cta(tmem_base, TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int EPI, bool STAGED, bool MMA2>
int launch_kernel(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tc2,
           const CUtensorMap& taux, const CUtensorMap& tbh, const Params& p, cudaStream_t stream) {
  CM3P_ENSURE_DYN_SMEM((gemm_bf16_sm100_kernel<EPI, STAGED, MMA2>), SMEM_BYTES);
  const int64_t tiles_m = (p.M + BM - 1) / BM, tiles_n = (p.N + BN - 1) / BN;
  const int64_t units = ((tiles_m + p.cluster - 1) / p.cluster) * tiles_n * p.splits;
  const int64_t max_clusters = num_sms() / p.cluster;
  const int grid = static_cast<int>((units < max_clusters ? units : max_clusters) * p.cluster);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = p.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CM3P_CUDA_TRY(cudaLaunchKernelEx(&cfg, gemm_bf16_sm100_kernel<EPI, STAGED, MMA2>, ta, tb, tc, tc2, taux, tbh, p));
  return kOk;
}

template <int EPI, bool STAGED>
int launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const CUtensorMap& tc2,
           const CUtensorMap& taux, const CUtensorMap& tbh, const Params& p, cudaStream_t stream) {
  return p.mma2 ? launch_kernel<EPI, STAGED, true>(ta, tb, tc, tc2, taux, tbh, p, stream)
                : launch_kernel<EPI, STAGED, false>(ta, tb, tc, tc2, taux, tbh, p, stream);
}

template <int EPI>
int launch_any(bool staged, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc,
               const CUtensorMap& tc2, const CUtensorMap& taux, const CUtensorMap& tbh, const Params& p,
               cudaStream_t stream) {
  if constexpr (EPI == EPI_SCALE_F32) {
    return launch<EPI, false>(ta, tb, tc, tc2, taux, tbh, p, stream);
  } else {
    return staged ? launch<EPI, true>(ta, tb, tc, tc2, taux, tbh, p, stream)
                  : launch<EPI, false>(ta, tb, tc, tc2, taux, tbh, p, stream);
  }
}

}  // namespace

int gemm_bf16(const GemmArgs& g_in, cudaStream_t stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  GemmArgs g = g_in;
  int conv_rows = 0, conv_per_win = 0;  // output frames per window; M tiles (fwd) / K blocks (wgrad) per window
  if (g.conv_mode != 0) {
    CM3P_REQUIRE(g.conv_mode == 1 || g.conv_mode == 2, kBadShape, "conv: unknown mode %d", g.conv_mode);
    CM3P_REQUIRE((g.conv_stride == 1 || g.conv_stride == 2) && g.conv_batch > 0 && g.conv_frames > 0 &&
                     g.conv_frames % g.conv_stride == 0 && g.conv_cin > 0 && g.conv_cin % 8 == 0 &&
                     g.conv_cpad % BK == 0 && g.conv_cpad >= g.conv_cin,
                 kBadShape, "conv: stride %d batch %d frames %d c_in %d c_pad %d unsupported", g.conv_stride,
                 g.conv_batch, g.conv_frames, g.conv_cin, g.conv_cpad);
    CM3P_REQUIRE(g.group_m == 0 && !g.stats_out && !g.row_stats, kBadShape, "conv: not combinable with grouping / LN folding");
    conv_rows = g.conv_frames / g.conv_stride;
    if (g.conv_mode == 1) {
      CM3P_REQUIRE(g.epilogue == EPI_BIAS || g.epilogue == EPI_BIAS_GELU, kBadShape, "conv forward: epilogue must be BIAS(_GELU)");
      conv_per_win = (conv_rows + BM - 1) / BM;
      g.M = static_cast<int64_t>(g.conv_batch) * conv_per_win * BM;  // virtual rows: tiles never straddle windows
      g.K = 3LL * g.conv_cpad;
      g.trans_a = 0; g.trans_b = 0;
      g.ldb = 3LL * g.conv_cpad;
      g.ldc = g.N;
    } else {
      CM3P_REQUIRE(g.epilogue == EPI_SCALE_F32 && g.accumulate, kBadShape, "conv wgrad: needs EPI_SCALE_F32 with accumulate");
      conv_per_win = (conv_rows + BK - 1) / BK;
      g.N = 3LL * g.conv_cpad;
      g.K = static_cast<int64_t>(g.conv_batch) * conv_per_win * BK;  // virtual K: blocks never straddle windows
      g.trans_a = 1; g.trans_b = 1;
      if (g.ldc == 0) g.ldc = g.N;
    }
  }
  CM3P_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, kBadShape, "gemm: empty problem M=%lld N=%lld K=%lld", (long long)g.M,
               (long long)g.N, (long long)g.K);
  CM3P_REQUIRE(g.epilogue >= 0 && g.epilogue < EPI_COUNT, kBadShape, "gemm: unknown epilogue %d", g.epilogue);
  CM3P_REQUIRE(g.a && g.b && g.c, kBadShape, "gemm: null operand");
  if (g.epilogue == EPI_GEGLU || g.epilogue == EPI_GEGLU_SAVE)
    CM3P_REQUIRE(g.N % 32 == 0 && g.ldc % 8 == 0, kBadShape, "gemm(geglu): N=%lld must be a multiple of 32",
                 (long long)g.N);
  if (g.epilogue == EPI_ROPE)
    CM3P_REQUIRE(g.N % 64 == 0 && g.rope_cols % 64 == 0 && g.positions && g.rope_table, kBadShape,
                 "gemm(rope): N and rope_cols must be multiples of 64 and positions/table given");
  if (g.epilogue == EPI_RESIDUAL || g.epilogue == EPI_BIAS || g.epilogue == EPI_BIAS_GELU)
    CM3P_REQUIRE(g.aux != nullptr, kBadShape, "gemm: epilogue %d needs aux", g.epilogue);
  const int out_elt = (g.epilogue == EPI_SCALE_F32) ? 4 : 2;
  auto aligned16 = [](const void* ptr, int64_t ld, int elt) {
    return (reinterpret_cast<uintptr_t>(ptr) % 16) == 0 && (ld * elt) % 16 == 0;
  };
  bool vec = aligned16(g.c, g.ldc, out_elt);
  if (g.epilogue == EPI_RESIDUAL) vec = vec && aligned16(g.aux, g.ld_aux, 2);
  if (g.epilogue == EPI_GEGLU_SAVE) vec = vec && aligned16(g.c2, g.ldc2, 2);
  if (g.epilogue == EPI_GEGLU || g.epilogue == EPI_GEGLU_SAVE)
    CM3P_REQUIRE(vec, kBadAlignment, "gemm(geglu): outputs must be 16-byte aligned with ld %% 8 == 0");

  // grouped GEMM: `groups` independent problems of group_m x N x K whose operands are stacked along their outer
  // dimension (A: M rows, or K rows when transposed; B: N rows, or K rows when transposed; C/aux: M rows)
  int64_t groups = 1;
  if (g.group_m > 0) {
    CM3P_REQUIRE(g.M % g.group_m == 0 && g.group_m % (2 * BM) == 0 && g.K % BK == 0, kBadShape,
                 "gemm(grouped): M=%lld must be a multiple of group_m=%lld, group_m of %d and K=%lld of %d",
                 (long long)g.M, (long long)g.group_m, 2 * BM, (long long)g.K, BK);
    CM3P_REQUIRE(!(g.epilogue == EPI_SCALE_F32 && g.accumulate) && !g.stats_out && !g.row_stats, kBadShape,
                 "gemm(grouped): split-K accumulation and LayerNorm folding are not supported");
    groups = g.M / g.group_m;
  }
  CUtensorMap ta, tb;
  if (g.conv_mode != 0) {
    // x [B, F, C] seen as (c, frame parity, frame / stride, window)
    const void* x = g.conv_mode == 1 ? g.a : g.b;
    const uint64_t C = g.conv_cin, S = g.conv_stride, F = g.conv_frames;
    const uint64_t dims[4] = {C, S, F / S, static_cast<uint64_t>(g.conv_batch)};
    const uint64_t pitches[3] = {C * 2, S * C * 2, F * C * 2};
    const uint32_t box[4] = {BK, 1, static_cast<uint32_t>(g.conv_mode == 1 ? BM : BK), 1};
    rc = encode_tmap_4d_bf16(g.conv_mode == 1 ? &ta : &tb, x, dims, pitches, box);
    if (rc != kOk) return rc;
    if (g.conv_mode == 1) {
      rc = encode_tmap_2d_bf16(&tb, g.b, g.K, g.N, g.ldb * 2, BK, BN);
    } else {  // dz [B, rows, M] as (m, frame, window): frames past the window are zero-filled
      rc = encode_tmap_3d_bf16(&ta, g.a, g.M, conv_rows, g.conv_batch, static_cast<uint64_t>(g.M) * 2,
                               static_cast<uint64_t>(conv_rows) * g.M * 2, 64, BK, 1);
    }
    if (rc != kOk) return rc;
  } else {
  if (!g.trans_a)
    rc = encode_tmap_2d_bf16(&ta, g.a, g.K, g.M, g.lda * 2, BK, BM);
  else
    rc = encode_tmap_2d_bf16(&ta, g.a, g.group_m > 0 ? g.group_m : g.M, g.K * groups, g.lda * 2, 64, BK);
  if (rc != kOk) return rc;
  if (!g.trans_b)
    rc = encode_tmap_2d_bf16(&tb, g.b, g.K, g.N * groups, g.ldb * 2, BK, BN);
  else
    rc = encode_tmap_2d_bf16(&tb, g.b, g.N, g.K * groups, g.ldb * 2, 64, BK);
  if (rc != kOk) return rc;
  }

  // staged epilogue: output (and residual / raw) rows must allow TMA (16-byte aligned base and pitch)
  const bool staged = vec && g.epilogue != EPI_SCALE_F32;
  CUtensorMap tc = ta, tc2 = ta, taux = ta;  // placeholders when unused
  if (staged) {
    const int64_t n_out = (g.epilogue == EPI_GEGLU) ? g.N / 2 : g.N;
    if (g.epilogue == EPI_GEGLU_SAVE) {
      rc = encode_tmap_2d_bf16(&tc2, g.c2, g.N, g.M, g.ldc2 * 2, 64, BM);
      if (rc == kOk) rc = encode_tmap_2d_bf16(&tc, g.c, g.N / 2, g.M, g.ldc * 2, 64, BM);  // product, via staging
    } else if (g.conv_mode == 1) {
      rc = encode_tmap_3d_bf16(&tc, g.c, g.N, conv_rows, g.conv_batch, static_cast<uint64_t>(g.N) * 2,
                               static_cast<uint64_t>(conv_rows) * g.N * 2, 64, BM, 1);
    } else {
      rc = encode_tmap_2d_bf16(&tc, g.c, n_out, g.M, g.ldc * 2, 64, BM);
    }
    if (rc != kOk) return rc;
    if (g.epilogue == EPI_RESIDUAL) {
      rc = encode_tmap_2d_bf16(&taux, g.aux, g.N, g.M, g.ld_aux * 2, 64, BM);
      if (rc != kOk) return rc;
    }
  }

  Params p;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.c = g.c; p.ldc = g.ldc;
  p.aux = g.aux; p.ld_aux = g.ld_aux;
  p.c2 = g.c2; p.ldc2 = g.ldc2;
  p.scale = g.scale;
  p.positions = g.positions;
  p.rope_table = reinterpret_cast<const float2*>(g.rope_table);
  p.rope_cols = g.rope_cols;
  p.trans_a = g.trans_a; p.trans_b = g.trans_b;
  p.accumulate = g.accumulate;
  p.vec_c = vec ? 1 : 0;
  p.stats_out = g.stats_out;
  p.row_stats = g.row_stats;
  p.col_corr = g.col_corr;
  p.ln_eps = g.ln_eps;
  p.tile_sem = nullptr;
  p.mn_tiles = 0;
  p.group_m = g.group_m;
  p.conv_mode = g.conv_mode;
  p.conv_stride = g.conv_stride;
  p.conv_per_win = conv_per_win;
  p.conv_kb_per_tap = g.conv_cpad / BK;
  p.conv_cpad = g.conv_cpad;
  if (g.conv_mode == 1)
    CM3P_REQUIRE(vec && g.N % 8 == 0, kBadAlignment, "conv forward: output must be 16-byte aligned with C_out %% 8 == 0");
  if (g.stats_out)
    CM3P_REQUIRE(g.epilogue == EPI_RESIDUAL && vec, kBadShape,
                 "gemm: stats_out needs the staged EPI_RESIDUAL epilogue (16-byte aligned bf16 rows)");
  if (g.row_stats)
    CM3P_REQUIRE((g.epilogue == EPI_ROPE || g.epilogue == EPI_GEGLU || g.epilogue == EPI_GEGLU_SAVE) && vec &&
                     g.col_corr != nullptr && g.N % 64 == 0,
                 kBadShape, "gemm: row_stats needs a staged ROPE/GEGLU epilogue, col_corr and N %% 64 == 0");
  p.splits = 1;
  const int num_kb = static_cast<int>((g.K + BK - 1) / BK);
  p.kb_per_split = num_kb;
  if (g.epilogue == EPI_SCALE_F32 && g.accumulate) {
    // weight gradients: few output tiles, K = number of tokens.  Split K so that ~4 work units per SM
    // exist; partial sums meet in C through fp32 atomics.
    const int64_t tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    // Pick the split count that minimises (waves of work units) x (K per unit): e.g. 54 output tiles on 148
    // SMs -> 8 or 19 splits fill whole waves, where a plain "4 units per SM" rule leaves the last wave 65% full.
    const int sms = num_sms() > 0 ? num_sms() : 148;
    int64_t max_splits = num_kb / 8;  // at least 8 k-blocks (512 K elements) per unit
    if (max_splits > (8LL * sms) / tiles + 1) max_splits = (8LL * sms) / tiles + 1;
    if (max_splits < 1) max_splits = 1;
    int64_t best = 1;
    double best_cost = 1e30;
    for (int64_t sp = 1; sp <= max_splits; ++sp) {
      const int64_t kb_per = (num_kb + sp - 1) / sp;
      const int64_t units = tiles * ((num_kb + kb_per - 1) / kb_per);
      const int64_t waves = (units + sms - 1) / sms;
      const double cost = static_cast<double>(waves) * (static_cast<double>(kb_per) + 6.0);  // +6: per-unit prologue/epilogue
      if (cost < best_cost * 0.995) {
        best_cost = cost;
        best = sp;
      }
    }
    p.kb_per_split = static_cast<int>((num_kb + best - 1) / best);
    p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  }

  // CTA pairs (2-CTA clusters) on adjacent M tiles share each B tile through TMA multicast: a third less
  // L2 -> SM traffic per CTA (ncu: the MMA thread waited ~30% of the time for operands without it).
  const int cluster_mode = get_option(kOptGemmCluster) == 1 ? 1 : 2;
  const int64_t tiles_m_total = (g.M + BM - 1) / BM;
  p.cluster = (cluster_mode == 2 && tiles_m_total >= 2 && g.N > BN / 2 && g.conv_mode != 2) ? 2 : 1;
  // pair MMA for plain / grouped / split-K GEMMs (the implicit-GEMM convolutions keep the multicast form)
  p.mma2 = (p.cluster == 2 && g.conv_mode == 0 && get_option(kOptGemmCluster) != 3) ? 1 : 0;
  CUtensorMap tbh = tb;
  if (p.cluster == 2 && !g.trans_b) {
    rc = encode_tmap_2d_bf16(&tbh, g.b, g.K, g.N * groups, g.ldb * 2, BK, BN / 2);
    if (rc != kOk) return rc;
  }
  if (g.epilogue == EPI_SCALE_F32 && g.accumulate && g.tile_sem && get_option(kOptWgradDeterministic)) {
    // ordered split-K accumulation through one turnstile counter per output tile
    const int64_t sems = ((tiles_m_total + p.cluster - 1) / p.cluster) * ((g.N + BN - 1) / BN) * p.cluster;
    CM3P_REQUIRE(sems <= g.tile_sem_count, kBadShape, "gemm: %lld output tiles exceed the %lld turnstile counters",
                 (long long)sems, (long long)g.tile_sem_count);
    p.tile_sem = g.tile_sem;
    p.accumulate = 2;
  }

  switch (g.epilogue) {
    case EPI_STORE: return launch_any<EPI_STORE>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_RESIDUAL: return launch_any<EPI_RESIDUAL>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_GELU: return launch_any<EPI_GELU>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_BIAS_GELU: return launch_any<EPI_BIAS_GELU>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_BIAS: return launch_any<EPI_BIAS>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_GEGLU: return launch_any<EPI_GEGLU>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_GEGLU_SAVE: return launch_any<EPI_GEGLU_SAVE>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_ROPE: return launch_any<EPI_ROPE>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
    case EPI_SCALE_F32: return launch_any<EPI_SCALE_F32>(staged, ta, tb, tc, tc2, taux, tbh, p, stream);
  }
  return set_error(kBadShape, "gemm: unreachable epilogue %d", g.epilogue);
}

}  // namespace cm3p
