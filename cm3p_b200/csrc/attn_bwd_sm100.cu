// Unpadded (varlen) attention backward for sm_100a, head_dim 64, global and +-window layers.
//
// Gradient of what attn_fwd_sm100.cu computes (reference: autograd through the third-party
// ModernBertAttention called from /root/reference/cm3p/modeling_cm3p.py:359-369,509-514,607-619):
//   P = softmax(Q K^T / 8 | mask),  O = P V
//   dV = P^T dO,  dP = dO V^T,  dZ = P o (dP - delta),  delta_i = <dO_i, O_i>,
//   dQ = dZ K / 8,  dK = dZ^T Q / 8
// followed by the inverse RoPE rotation on dQ / dK (the forward rotates q, k in the Wqkv GEMM
// epilogue), so the result is the gradient w.r.t. the *un-rotated* Wqkv output.
//
// Two kernels, no atomics, deterministic:
//   attn_bwd_dq_kernel   one CTA = (sequence, head, 128 queries), streams 64-row K/V tiles:
//                        S = Q K^T, dP = dO V^T (tcgen05, TMEM), dZ -> bf16 smem, dQ += dZ K (TMEM).
//                        Also produces delta[head][token] for the second kernel.
//   attn_bwd_dkv_kernel  one CTA = (sequence, head, 128 keys), streams 64-row Q/dO tiles, working on
//                        the transposed problem so that TMEM lanes are key rows:
//                        S^T = K Q^T, dP^T = V dO^T, P^T and dZ^T -> bf16 smem, dV += P^T dO, dK += dZ^T Q.
// P is recomputed from the forward's log2-domain logsumexp.  Each CTA uses 256 TMEM columns and
// < 100 KB of shared memory, so two CTAs share an SM and overlap each other's MMA and exp phases.
//
// Warps: 0..3 = element-wise stage / epilogue (thread t <-> TMEM lane t), 4 = TMA producer (+ TMEM
// alloc, + lse/delta staging), 5 = MMA issuer.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "attn.h"
#include "attn_bwd_common.cuh"
#include "common.h"
#include "ptx.cuh"

namespace cm3p {
namespace {
using namespace bwd_detail;

constexpr int BT = 128;  // outer tile (rows owned by the CTA == TMEM lanes)
constexpr int BI = 64;   // inner (streamed) tile
constexpr int D = 64;
constexpr int OUTER_BYTES = BT * D * 2;  // 16 KB
constexpr int INNER_BYTES = BI * D * 2;  // 8 KB
constexpr int STAGES = 2;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;

struct BwdParams {
  const int32_t* cu_seqlens;
  const __nv_bfloat16* out;   // [T, H]   forward output
  const __nv_bfloat16* dout;  // [T, H]
  const float* lse;           // [heads, T] log2-domain logsumexp from the forward
  float* delta;               // [heads, T] workspace: written by the dQ kernel, read by the dKV kernel
  __nv_bfloat16* dqkv;        // [T, 3, heads, 64]
  const int32_t* positions;   // [T] or nullptr (no inverse RoPE)
  const float2* rope_table;   // [max_pos][32] (cos, sin) or nullptr
  int64_t total_tokens;
  int heads;
  int hidden;
  int window;
  float scale_log2;  // (1/8) * log2(e)
  float scale;       // 1/8
  int ctas_per_seq;  // grid.x = batch * ctas_per_seq (grid.z stops at 65535 sequences)
};

// Allowed inner-tile columns [a,b) of one row and, per 32-column chunk, the state of its warp (32 rows):
// 0 = no row of the warp needs the chunk (skip the exp work, write zeros), 1 = per-element masking,
// 2 = every column allowed for every row (no predicates).
struct Band {
  int a, b;
  int state[2];
};
__device__ __forceinline__ Band band_of(int row_pos, int warp_pos, int t0, int len, int window, bool row_valid) {
  Band r;
  int a = 0, b = min(BI, len - t0);
  int wa = 0, wb = b, ia = 0, ib = b;
  if (window >= 0) {
    a = max(a, row_pos - window - t0);
    b = min(b, row_pos + window + 1 - t0);
    wa = max(wa, warp_pos - window - t0);
    wb = min(wb, warp_pos + 31 + window + 1 - t0);
    ia = max(ia, warp_pos + 31 - window - t0);
    ib = min(ib, warp_pos + window + 1 - t0);
  }
  if (warp_pos + 31 >= len) { ia = BI; ib = 0; }
  if (warp_pos >= len) { wa = BI; wb = 0; }
  if (!row_valid) b = 0;
  r.a = a;
  r.b = b;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const int c0 = q * 32, c1 = q * 32 + 32;
    r.state[q] = (c1 <= wa || c0 >= wb) ? 0 : ((c0 >= ia && c1 <= ib) ? 2 : 1);
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// dQ kernel.  smem: Q 16K | dO 16K | K 2x8K | V 2x8K | dZ 16K | barriers.   TMEM: S [0,64) dP [64,128) dQ [128,192)
constexpr int DQ_SMEM_TILES = 2 * OUTER_BYTES + STAGES * 2 * INNER_BYTES + OUTER_BYTES;  // 80 KB
constexpr int DQ_SMEM_BYTES = DQ_SMEM_TILES + 256;

__global__ void __launch_bounds__(THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tma_qkv128, const __grid_constant__ CUtensorMap tma_qkv64,
                   const __grid_constant__ CUtensorMap tma_do128, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.ctas_per_seq, head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int q0 = (blockIdx.x % p.ctas_per_seq) * BT;
  if (q0 >= len) return;

  uint8_t* smem_q = smem;
  uint8_t* smem_do = smem + OUTER_BYTES;
  uint8_t* smem_k = smem + 2 * OUTER_BYTES;
  uint8_t* smem_v = smem_k + STAGES * INNER_BYTES;
  uint8_t* smem_dz = smem_v + STAGES * INNER_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_SMEM_TILES);
  uint64_t* q_full = bars;        // Q + dO landed
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // S and dP of the current tile are in TMEM
  uint64_t* dz_full = bars + 6;   // dZ in smem, S/dP consumed (128 arrivals)
  uint64_t* acc_full = bars + 7;  // all MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();

  int kv_lo = 0, kv_hi = len - 1;
  if (p.window >= 0) {
    kv_lo = max(0, q0 - p.window);
    kv_hi = min(len - 1, q0 + BT - 1 + p.window);
  }
  const int tile_lo = kv_lo / BI;
  const int n_tiles = kv_hi / BI - tile_lo + 1;

  if (warp == 5 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(dz_full, 128);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_qkv128);
      ptx::prefetch_tmap(&tma_qkv64);
      ptx::prefetch_tmap(&tma_do128);
    }
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_S = 0, TM_DP = 64, TM_DQ = 128;

  if (warp == 4) {
    if (lane == 0) {
      const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
      ptx::mbar_arrive_expect_tx(q_full, 2 * OUTER_BYTES);
      ptx::tma_load_2d(smem_q, &tma_qkv128, q_full, col_q, seq_start + q0);
      ptx::tma_load_2d(smem_do, &tma_do128, q_full, head * D, seq_start + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        ptx::mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * INNER_BYTES);
        const int row = seq_start + (tile_lo + j) * BI;
        ptx::tma_load_2d(smem_k + s * INNER_BYTES, &tma_qkv64, &kv_full[s], col_k, row);
        ptx::tma_load_2d(smem_v + s * INNER_BYTES, &tma_qkv64, &kv_full[s], col_v, row);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BT, BI, 0, 0);   // [128 q] x [64 kv], both K-major (d)
      const uint32_t idesc_dq = ptx::umma_idesc_bf16(BT, D, 0, 1);   // A = dZ K-major (kv), B = K MN-major (d)
      const uint32_t q_addr = ptx::smem_u32(smem_q), do_addr = ptx::smem_u32(smem_do);
      const uint32_t dz_addr = ptx::smem_u32(smem_dz);
      auto issue_s_dp = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait(&kv_full[s], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem_k + s * INNER_BYTES);
        const uint32_t v_addr = ptx::smem_u32(smem_v + s * INNER_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_S, ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_DP, ptx::umma_smem_desc_sw128(do_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(v_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      };
      ptx::mbar_wait(q_full, 0);
      issue_s_dp(0);
      for (int j = 0; j < n_tiles; ++j) {
        ptx::mbar_wait(dz_full, j & 1);
        ptx::tc_fence_after();
        const int s = j & 1;
        const uint32_t k_addr = ptx::smem_u32(smem_k + s * INNER_BYTES);
#pragma unroll
        for (int k = 0; k < BI / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_DQ, ptx::umma_smem_desc_sw128(dz_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_dq, (j | k) != 0 ? 1u : 0u);
        ptx::umma_commit(&kv_empty[s]);
        // S/dP of tile j+1 are issued behind dQ_j: their commit (s_full) therefore also tells the
        // element-wise warps that the MMAs reading the dZ buffer have retired.
        if (j + 1 < n_tiles) issue_s_dp(j + 1);
      }
      ptx::umma_commit(acc_full);
    }
  } else {
    const int t = threadIdx.x;  // query row inside the tile == TMEM lane
    const int qi = q0 + t;
    const bool valid = qi < len;
    const int64_t row = static_cast<int64_t>(seq_start) + qi;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    // delta_i = <dO_i, O_i>
    float delta = 0.f, lse = 0.f;
    if (valid) {
      const uint4* po = reinterpret_cast<const uint4*>(p.out + row * p.hidden + head * D);
      const uint4* pd = reinterpret_cast<const uint4*>(p.dout + row * p.hidden + head * D);
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        float a[8], b[8];
        unpack8f(__ldg(po + i), a);
        unpack8f(__ldg(pd + i), b);
#pragma unroll
        for (int k = 0; k < 8; ++k) delta += a[k] * b[k];
      }
      lse = p.lse[static_cast<int64_t>(head) * p.total_tokens + row];
      p.delta[static_cast<int64_t>(head) * p.total_tokens + row] = delta;
    }
    const float c = p.scale_log2;
    const float dsc = delta * p.scale;  // dZ = P * (dP*scale - delta*scale)
    const int warp_pos = q0 + warp * 32;
    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = (tile_lo + j) * BI;
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after();
      const Band bd = band_of(qi, warp_pos, kv0, len, p.window, valid);
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (bd.state[q] == 0) {
          store_zero_units(smem_dz, t, q * 4);
          continue;
        }
        uint32_t rs[32], rp[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + TM_S + lane_off + q * 32, rs);
        ptx::tmem_ld_32x32b_x32(tmem_base + TM_DP + lane_off + q * 32, rp);
        ptx::tmem_ld_wait();
        uint32_t packed[16];
        if (bd.state[q] == 2) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float p0 = ptx::ex2_approx(__uint_as_float(rs[i]) * c - lse);
            const float p1 = ptx::ex2_approx(__uint_as_float(rs[i + 1]) * c - lse);
            packed[i >> 1] = ptx::pack_bf16x2(p0 * (__uint_as_float(rp[i]) * p.scale - dsc),
                                              p1 * (__uint_as_float(rp[i + 1]) * p.scale - dsc));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const int kj = q * 32 + i;
            const float p0 = (kj >= bd.a && kj < bd.b) ? ptx::ex2_approx(__uint_as_float(rs[i]) * c - lse) : 0.f;
            const float p1 =
                (kj + 1 >= bd.a && kj + 1 < bd.b) ? ptx::ex2_approx(__uint_as_float(rs[i + 1]) * c - lse) : 0.f;
            packed[i >> 1] = ptx::pack_bf16x2(p0 * (__uint_as_float(rp[i]) * p.scale - dsc),
                                              p1 * (__uint_as_float(rp[i + 1]) * p.scale - dsc));
          }
        }
        store_row_units(smem_dz, t, q * 4, packed);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(dz_full);
    }
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const float2* cs = nullptr;
    if (p.rope_table && p.positions && valid) cs = p.rope_table + static_cast<int64_t>(p.positions[row]) * 32;
    store_grad_row(tmem_base + TM_DQ + lane_off, p.dqkv + row * 3 * p.hidden + head * D, cs, valid);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// dKV kernel.  smem: K 16K | V 16K | Q 2x8K | dO 2x8K | P^T 16K | dZ^T 16K | lse/delta 2x512 B | barriers
// TMEM: S^T [0,64)  dP^T [64,128)  dK [128,192)  dV [192,256)
constexpr int DKV_SMEM_TILES = 2 * OUTER_BYTES + STAGES * 2 * INNER_BYTES + 2 * OUTER_BYTES;  // 96 KB
constexpr int DKV_VEC_BYTES = STAGES * 2 * BI * 4;                                            // 1 KB
constexpr int DKV_SMEM_BYTES = DKV_SMEM_TILES + DKV_VEC_BYTES + 256;

__global__ void __launch_bounds__(THREADS, 2)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tma_qkv128, const __grid_constant__ CUtensorMap tma_qkv64,
                    const __grid_constant__ CUtensorMap tma_do64, const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.ctas_per_seq, head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int k0 = (blockIdx.x % p.ctas_per_seq) * BT;
  if (k0 >= len) return;

  uint8_t* smem_k = smem;
  uint8_t* smem_v = smem + OUTER_BYTES;
  uint8_t* smem_q = smem + 2 * OUTER_BYTES;
  uint8_t* smem_do = smem_q + STAGES * INNER_BYTES;
  uint8_t* smem_pt = smem_do + STAGES * INNER_BYTES;
  uint8_t* smem_dzt = smem_pt + OUTER_BYTES;
  float* smem_vec = reinterpret_cast<float*>(smem + DKV_SMEM_TILES);  // [stage][lse 64 | delta 64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_SMEM_TILES + DKV_VEC_BYTES);
  uint64_t* kv_full = bars;
  uint64_t* qdo_full = bars + 1;   // [2]  TMA bytes + 32 staging-lane arrivals
  uint64_t* qdo_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* pz_full = bars + 6;    // P^T and dZ^T in smem (128 arrivals)
  uint64_t* acc_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();

  int q_lo = 0, q_hi = len - 1;
  if (p.window >= 0) {
    q_lo = max(0, k0 - p.window);
    q_hi = min(len - 1, k0 + BT - 1 + p.window);
  }
  const int tile_lo = q_lo / BI;
  const int n_tiles = q_hi / BI - tile_lo + 1;

  if (warp == 5 && lane == 0) {
    ptx::mbar_init(kv_full, 1);
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&qdo_full[s], 33);
      ptx::mbar_init(&qdo_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(pz_full, 128);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_qkv128);
      ptx::prefetch_tmap(&tma_qkv64);
      ptx::prefetch_tmap(&tma_do64);
    }
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_ST = 0, TM_DPT = 64, TM_DK = 128, TM_DV = 192;

  if (warp == 4) {
    // ------------------------------------------------------------------ producer (whole warp)
    const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(kv_full, 2 * OUTER_BYTES);
      ptx::tma_load_2d(smem_k, &tma_qkv128, kv_full, col_k, seq_start + k0);
      ptx::tma_load_2d(smem_v, &tma_qkv128, kv_full, col_v, seq_start + k0);
    }
    const float* lse_h = p.lse + static_cast<int64_t>(head) * p.total_tokens;
    const float* delta_h = p.delta + static_cast<int64_t>(head) * p.total_tokens;
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i & 1;
      ptx::mbar_wait(&qdo_empty[s], ((i >> 1) & 1) ^ 1);
      const int64_t row = static_cast<int64_t>(seq_start) + (tile_lo + i) * BI;
      if (lane == 0) {
        ptx::mbar_arrive_expect_tx(&qdo_full[s], 2 * INNER_BYTES);
        ptx::tma_load_2d(smem_q + s * INNER_BYTES, &tma_qkv64, &qdo_full[s], col_q, static_cast<int32_t>(row));
        ptx::tma_load_2d(smem_do + s * INNER_BYTES, &tma_do64, &qdo_full[s], col_q, static_cast<int32_t>(row));
      }
      float* vec = smem_vec + s * 2 * BI;
#pragma unroll
      for (int h = 0; h < BI; h += 32) {
        const int64_t r = row + h + lane;
        const bool ok = r < p.total_tokens;
        vec[h + lane] = ok ? lse_h[r] : 0.f;
        vec[BI + h + lane] = ok ? delta_h[r] * p.scale : 0.f;  // pre-scaled
      }
      ptx::mbar_arrive(&qdo_full[s]);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BT, BI, 0, 0);   // [128 kv] x [64 q], both K-major (d)
      const uint32_t idesc_acc = ptx::umma_idesc_bf16(BT, D, 0, 1);  // A = P^T / dZ^T K-major (q), B MN-major (d)
      const uint32_t k_addr = ptx::smem_u32(smem_k), v_addr = ptx::smem_u32(smem_v);
      const uint32_t pt_addr = ptx::smem_u32(smem_pt), dzt_addr = ptx::smem_u32(smem_dzt);
      auto issue_s_dp = [&](int i) {
        const int s = i & 1;
        ptx::mbar_wait(&qdo_full[s], (i >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t q_addr = ptx::smem_u32(smem_q + s * INNER_BYTES);
        const uint32_t do_addr = ptx::smem_u32(smem_do + s * INNER_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_ST, ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_DPT, ptx::umma_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      };
      ptx::mbar_wait(kv_full, 0);
      issue_s_dp(0);
      for (int i = 0; i < n_tiles; ++i) {
        ptx::mbar_wait(pz_full, i & 1);
        ptx::tc_fence_after();
        const int s = i & 1;
        const uint32_t q_addr = ptx::smem_u32(smem_q + s * INNER_BYTES);
        const uint32_t do_addr = ptx::smem_u32(smem_do + s * INNER_BYTES);
#pragma unroll
        for (int k = 0; k < BI / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_DV, ptx::umma_smem_desc_sw128(pt_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(do_addr + k * 2048, 8192, 1024), idesc_acc, (i | k) != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < BI / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_DK, ptx::umma_smem_desc_sw128(dzt_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_acc, (i | k) != 0 ? 1u : 0u);
        ptx::umma_commit(&qdo_empty[s]);
        if (i + 1 < n_tiles) issue_s_dp(i + 1);
      }
      ptx::umma_commit(acc_full);
    }
  } else {
    const int t = threadIdx.x;  // key row inside the tile == TMEM lane
    const int kj = k0 + t;
    const bool valid = kj < len;
    const int64_t row = static_cast<int64_t>(seq_start) + kj;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    const int warp_pos = k0 + warp * 32;
    for (int i = 0; i < n_tiles; ++i) {
      const int s = i & 1;
      const int q0i = (tile_lo + i) * BI;
      ptx::mbar_wait(s_full, i & 1);
      ptx::mbar_wait(&qdo_full[s], (i >> 1) & 1);  // already complete: orders the lse/delta staging writes
      ptx::tc_fence_after();
      const Band bd = band_of(kj, warp_pos, q0i, len, p.window, valid);
      const float4* lse4 = reinterpret_cast<const float4*>(smem_vec + s * 2 * BI);
      const float4* del4 = lse4 + BI / 4;  // delta * scale
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (bd.state[q] == 0) {
          store_zero_units(smem_pt, t, q * 4);
          store_zero_units(smem_dzt, t, q * 4);
          continue;
        }
        uint32_t rs[32], rp[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + TM_ST + lane_off + q * 32, rs);
        ptx::tmem_ld_32x32b_x32(tmem_base + TM_DPT + lane_off + q * 32, rp);
        ptx::tmem_ld_wait();
        uint32_t pp[16], pz[16];
        if (bd.state[q] == 2) {
#pragma unroll
          for (int i4 = 0; i4 < 32; i4 += 4) {
            const float4 l4 = lse4[(q * 32 + i4) >> 2];
            const float4 d4 = del4[(q * 32 + i4) >> 2];
            const float p0 = ptx::ex2_approx(__uint_as_float(rs[i4]) * c - l4.x);
            const float p1 = ptx::ex2_approx(__uint_as_float(rs[i4 + 1]) * c - l4.y);
            const float p2 = ptx::ex2_approx(__uint_as_float(rs[i4 + 2]) * c - l4.z);
            const float p3 = ptx::ex2_approx(__uint_as_float(rs[i4 + 3]) * c - l4.w);
            pp[i4 >> 1] = ptx::pack_bf16x2(p0, p1);
            pp[(i4 >> 1) + 1] = ptx::pack_bf16x2(p2, p3);
            pz[i4 >> 1] = ptx::pack_bf16x2(p0 * (__uint_as_float(rp[i4]) * p.scale - d4.x),
                                           p1 * (__uint_as_float(rp[i4 + 1]) * p.scale - d4.y));
            pz[(i4 >> 1) + 1] = ptx::pack_bf16x2(p2 * (__uint_as_float(rp[i4 + 2]) * p.scale - d4.z),
                                                 p3 * (__uint_as_float(rp[i4 + 3]) * p.scale - d4.w));
          }
        } else {
#pragma unroll
          for (int i4 = 0; i4 < 32; i4 += 4) {
            const float4 l4 = lse4[(q * 32 + i4) >> 2];
            const float4 d4 = del4[(q * 32 + i4) >> 2];
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
            const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
            float pv[4], zv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int qj = q * 32 + i4 + e;
              const float ex = ptx::ex2_approx(__uint_as_float(rs[i4 + e]) * c - ls[e]);
              pv[e] = (qj >= bd.a && qj < bd.b) ? ex : 0.f;
              zv[e] = pv[e] * (__uint_as_float(rp[i4 + e]) * p.scale - dl[e]);
            }
            pp[i4 >> 1] = ptx::pack_bf16x2(pv[0], pv[1]);
            pp[(i4 >> 1) + 1] = ptx::pack_bf16x2(pv[2], pv[3]);
            pz[i4 >> 1] = ptx::pack_bf16x2(zv[0], zv[1]);
            pz[(i4 >> 1) + 1] = ptx::pack_bf16x2(zv[2], zv[3]);
          }
        }
        store_row_units(smem_pt, t, q * 4, pp);
        store_row_units(smem_dzt, t, q * 4, pz);
      }
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(pz_full);
    }
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const float2* cs = nullptr;
    if (p.rope_table && p.positions && valid) cs = p.rope_table + static_cast<int64_t>(p.positions[row]) * 32;
    __nv_bfloat16* base = p.dqkv + row * 3 * p.hidden + head * D;
    store_grad_row(tmem_base + TM_DK + lane_off, base + p.hidden, cs, valid);
    store_grad_row(tmem_base + TM_DV + lane_off, base + 2 * p.hidden, nullptr, valid);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Packed short sequences (metadata tower): ONE kernel per (group of sequences with <= 128 tokens, head) computes
// dQ, dK and dV with 5 GEMMs and one exponentiation, because every operand (Q, K, V, dO) is the same 128-row
// slab.  The score tile is formed transposed (TMEM lanes = key rows):
//   S^T = K Q^T, dP^T = V dO^T            (two M128 N128 K64 MMAs)
//   P^T = exp2(S^T c - lse_i) o mask,  dZ^T = P^T o (dP^T/8 - delta_i/8)     (mask: same sequence, block diagonal)
//   dV = P^T dO      A = P^T from TMEM (packed bf16, written in place over S^T)
//   dK = dZ^T Q      A = dZ^T from shared memory, K-major
//   dQ = dZ K        A = the same shared-memory tile read MN-major (no second copy, no transpose)
// delta_i = <dO_i, O_i> is computed in the kernel (thread t = query row t) and exchanged through shared memory.
// smem: Q | K | dO | V+16K (dZ^T overwrites V once dP^T has retired) = 80 KB; TMEM 256 columns: 2 CTAs / SM.
namespace packed {

constexpr int CHUNK_SEQS = 64;
constexpr int SB_INTS = CHUNK_SEQS + 4;
constexpr int P_SMEM_TILES = 5 * OUTER_BYTES;  // Q, K, dO, V, +16 KB
constexpr int P_VEC_BYTES = 2 * BT * 4;        // lse[128], delta/8 [128]
constexpr int P_SMEM_BYTES = P_SMEM_TILES + P_VEC_BYTES + SB_INTS * 4 + 128;

struct PackedBwdParams {
  const int32_t* cu_seqlens;
  const int2* groups;
  const int32_t* n_groups;
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  const float* lse;
  __nv_bfloat16* dqkv;
  const int32_t* positions;
  const float2* rope_table;
  int64_t total_tokens;
  int hidden;
  float scale_log2;
  float scale;
};

__device__ __forceinline__ void row_bounds(const int* sb, int nseq, int t, int& lo, int& hi) {
  int a = 0, b = nseq;  // sb[a] <= t < sb[b]
  while (b - a > 1) {
    const int m = (a + b) >> 1;
    if (sb[m] <= t) a = m; else b = m;
  }
  lo = sb[a];
  hi = sb[b];
}

__global__ void __launch_bounds__(THREADS, 2)
attn_bwd_packed_kernel(const __grid_constant__ CUtensorMap tma_qkv128, const __grid_constant__ CUtensorMap tma_do128,
                       const PackedBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x, head = blockIdx.y;
  if (g >= *p.n_groups) return;
  const int2 grp = p.groups[g];
  const int tok0 = p.cu_seqlens[grp.x];
  const int rows = p.cu_seqlens[grp.y] - tok0;
  const int nseq = grp.y - grp.x;
  if (rows <= 0) return;

  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + OUTER_BYTES;
  uint8_t* smem_do = smem + 2 * OUTER_BYTES;
  uint8_t* smem_v = smem + 3 * OUTER_BYTES;
  uint8_t* smem_dzt = smem_v;  // [2 blocks of 64 queries][128 key rows][128 B]: written after dP^T has read V
  float* s_lse = reinterpret_cast<float*>(smem + P_SMEM_TILES);
  float* s_del = s_lse + BT;
  int* sb = reinterpret_cast<int*>(smem + P_SMEM_TILES + P_VEC_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_SMEM_TILES + P_VEC_BYTES + SB_INTS * 4);
  uint64_t* ld_full = bars;
  uint64_t* s_full = bars + 1;
  uint64_t* pz_full = bars + 2;
  uint64_t* acc_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();
  for (int i = threadIdx.x; i <= nseq; i += THREADS) sb[i] = p.cu_seqlens[grp.x + i] - tok0;
  if (warp == 5 && lane == 0) {
    ptx::mbar_init(ld_full, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(pz_full, 128);
    ptx::mbar_init(acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_qkv128);
      ptx::prefetch_tmap(&tma_do128);
    }
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // S^T [0,128) and dP^T [128,256); afterwards P^T (bf16 pairs) [0,64), dV [64,128), dK [128,192), dQ [192,256)
  constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_PT = 0, TM_DV = 64, TM_DK = 128, TM_DQ = 192;
  const int ksteps = (rows + 15) >> 4;

  if (warp == 4) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(ld_full, 4 * OUTER_BYTES);
      ptx::tma_load_2d(smem_q, &tma_qkv128, ld_full, head * D, tok0);
      ptx::tma_load_2d(smem_k, &tma_qkv128, ld_full, p.hidden + head * D, tok0);
      ptx::tma_load_2d(smem_v, &tma_qkv128, ld_full, 2 * p.hidden + head * D, tok0);
      ptx::tma_load_2d(smem_do, &tma_do128, ld_full, head * D, tok0);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BT, BT, 0, 0);
      const uint32_t idesc_acc = ptx::umma_idesc_bf16(BT, D, 0, 1);    // A K-major (or TMEM), B MN-major
      const uint32_t idesc_dq = ptx::umma_idesc_bf16(BT, D, 1, 1);     // A = dZ read MN-major from the dZ^T tile
      const uint32_t q_addr = ptx::smem_u32(smem_q), k_addr = ptx::smem_u32(smem_k);
      const uint32_t v_addr = ptx::smem_u32(smem_v), do_addr = ptx::smem_u32(smem_do);
      const uint32_t dzt_addr = ptx::smem_u32(smem_dzt);
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      ptx::mbar_wait(ld_full, 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        ptx::umma_bf16(tmem_base + TM_ST, ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                       ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        ptx::umma_bf16(tmem_base + TM_DPT, ptx::umma_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                       ptx::umma_smem_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
      ptx::umma_commit(s_full);
      ptx::mbar_wait(pz_full, 0);
      ptx::tc_fence_after();
      const uint32_t do_lo = ptx::umma_desc_lo(do_addr, 8192);
      for (int k = 0; k < ksteps; ++k)  // dV = P^T dO: 16 queries per step = 8 TMEM columns of P^T
        ptx::umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_PT + k * 8, do_lo + ((k * 2048) >> 4), HI, idesc_acc,
                          k != 0 ? 1u : 0u);
      for (int k = 0; k < ksteps; ++k)  // dK = dZ^T Q
        ptx::umma_bf16(tmem_base + TM_DK,
                       ptx::umma_smem_desc_sw128(dzt_addr + (k >> 2) * (BT * 128) + (k & 3) * 32, 16, 1024),
                       ptx::umma_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_acc, k != 0 ? 1u : 0u);
      for (int k = 0; k < ksteps; ++k)  // dQ = dZ K: 16 keys per step = 16 rows of the dZ^T tile
        ptx::umma_bf16(tmem_base + TM_DQ, ptx::umma_smem_desc_sw128(dzt_addr + k * 2048, BT * 128, 1024),
                       ptx::umma_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_dq, k != 0 ? 1u : 0u);
      ptx::umma_commit(acc_full);
    }
  } else {
    const int t = threadIdx.x;  // key row (and, for delta / the dQ epilogue, query row) inside the group
    const bool valid = t < rows;
    const int64_t row = static_cast<int64_t>(tok0) + t;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    // delta_t = <dO_t, O_t>, lse_t -> shared memory (columns of the transposed tile are queries)
    float delta = 0.f, lse = 0.f;
    if (valid) {
      const uint4* po = reinterpret_cast<const uint4*>(p.out + row * p.hidden + head * D);
      const uint4* pd = reinterpret_cast<const uint4*>(p.dout + row * p.hidden + head * D);
#pragma unroll
      for (int i = 0; i < D / 8; ++i) {
        float a[8], b[8];
        unpack8f(__ldg(po + i), a);
        unpack8f(__ldg(pd + i), b);
#pragma unroll
        for (int k = 0; k < 8; ++k) delta += a[k] * b[k];
      }
      lse = p.lse[static_cast<int64_t>(head) * p.total_tokens + row];
    }
    s_lse[t] = lse;
    s_del[t] = delta * p.scale;
    int lo = 0, hi = 0;
    if (valid) row_bounds(sb, nseq, t, lo, hi);
    const int wa = __reduce_min_sync(0xffffffffu, valid ? lo : BT);
    const int wb = __reduce_max_sync(0xffffffffu, valid ? hi : 0);
    ptx::named_bar_sync(1, 128);  // lse / delta of all rows are in shared memory
    const float c = p.scale_log2;
    ptx::mbar_wait(s_full, 0);  // S^T and dP^T are in TMEM; V has been read (dZ^T may overwrite it)
    ptx::tc_fence_after();
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      const int c0 = ch * 32;
      uint8_t* dz_tile = smem_dzt + (ch >> 1) * (BT * 128);
      const int unit0 = (ch & 1) * 4;
      uint32_t pp[16], pz[16];
      if (c0 + 32 <= wa || c0 >= wb) {  // no row of this warp shares a sequence with these queries
#pragma unroll
        for (int i = 0; i < 16; ++i) pp[i] = 0u;
        ptx::tmem_st_32x32b_x16(tmem_base + TM_PT + lane_off + ch * 16, pp);
        store_zero_units(dz_tile, t, unit0);
        continue;
      }
      uint32_t rs[32], rp[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + TM_ST + lane_off + c0, rs);
      ptx::tmem_ld_32x32b_x32(tmem_base + TM_DPT + lane_off + c0, rp);
      ptx::tmem_ld_wait();
      const float4* lse4 = reinterpret_cast<const float4*>(s_lse + c0);
      const float4* del4 = reinterpret_cast<const float4*>(s_del + c0);
#pragma unroll
      for (int i4 = 0; i4 < 32; i4 += 4) {
        const float4 l4 = lse4[i4 >> 2];
        const float4 d4 = del4[i4 >> 2];
        const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
        const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
        float pv[4], zv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qj = c0 + i4 + e;
          const float ex = ptx::ex2_approx(__uint_as_float(rs[i4 + e]) * c - ls[e]);
          pv[e] = (qj >= lo && qj < hi) ? ex : 0.f;
          zv[e] = (qj >= lo && qj < hi) ? ex * (__uint_as_float(rp[i4 + e]) * p.scale - dl[e]) : 0.f;
        }
        pp[i4 >> 1] = ptx::pack_bf16x2(pv[0], pv[1]);
        pp[(i4 >> 1) + 1] = ptx::pack_bf16x2(pv[2], pv[3]);
        pz[i4 >> 1] = ptx::pack_bf16x2(zv[0], zv[1]);
        pz[(i4 >> 1) + 1] = ptx::pack_bf16x2(zv[2], zv[3]);
      }
      // P^T in place: columns [16 ch, 16 ch + 16) belong to score chunks this thread has already consumed
      ptx::tmem_st_32x32b_x16(tmem_base + TM_PT + lane_off + ch * 16, pp);
      store_row_units(dz_tile, t, unit0, pz);
    }
    ptx::tmem_st_wait();
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(pz_full);
    ptx::mbar_wait(acc_full, 0);
    ptx::tc_fence_after();
    const float2* cs = nullptr;
    if (p.rope_table && p.positions && valid) cs = p.rope_table + static_cast<int64_t>(p.positions[row]) * 32;
    __nv_bfloat16* base = p.dqkv + row * 3 * p.hidden + head * D;
    store_grad_row(tmem_base + TM_DQ + lane_off, base, cs, valid);
    store_grad_row(tmem_base + TM_DK + lane_off, base + p.hidden, cs, valid);
    store_grad_row(tmem_base + TM_DV + lane_off, base + 2 * p.hidden, nullptr, valid);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace packed

}  // namespace

int attn_varlen_bwd(const AttnBwdArgs& a, cudaStream_t stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  CM3P_REQUIRE(a.head_dim == 64, kBadShape, "attn_bwd: head_dim %d unsupported (kernel is specialised for 64)",
               a.head_dim);
  CM3P_REQUIRE(a.batch > 0 && a.heads > 0 && a.total_tokens > 0 && a.max_seqlen > 0, kBadShape,
               "attn_bwd: empty problem (batch=%d heads=%d tokens=%lld max_seqlen=%d)", a.batch, a.heads,
               (long long)a.total_tokens, a.max_seqlen);
  CM3P_REQUIRE(a.qkv && a.out && a.dout && a.lse && a.delta && a.dqkv && a.cu_seqlens, kBadShape,
               "attn_bwd: null pointer");
  CM3P_REQUIRE((a.positions == nullptr) == (a.rope_table == nullptr), kBadShape,
               "attn_bwd: positions and rope_table must be given together");
  CM3P_REQUIRE(a.heads <= 65535, kBadShape, "attn_bwd: heads=%d exceeds the grid limit", a.heads);
  const uint64_t H = static_cast<uint64_t>(a.heads) * 64;
  const uint64_t T = static_cast<uint64_t>(a.total_tokens);
  if (a.groups) {
    CM3P_REQUIRE(a.n_groups && a.max_groups > 0, kBadShape, "attn_bwd(packed): n_groups / max_groups missing");
    CM3P_REQUIRE(a.max_seqlen <= BT && a.window < 0, kBadShape,
                 "attn_bwd(packed): needs max_seqlen <= 128 (got %d) and a global layer (window %d)", a.max_seqlen,
                 a.window);
    CUtensorMap qkv128, do128;
    if ((rc = encode_tmap_2d_bf16(&qkv128, a.qkv, 3 * H, T, 3 * H * 2, 64, BT)) != kOk) return rc;
    if ((rc = encode_tmap_2d_bf16(&do128, a.dout, H, T, H * 2, 64, BT)) != kOk) return rc;
    CM3P_ENSURE_DYN_SMEM(packed::attn_bwd_packed_kernel, packed::P_SMEM_BYTES);
    packed::PackedBwdParams pp;
    pp.cu_seqlens = a.cu_seqlens;
    pp.groups = reinterpret_cast<const int2*>(a.groups);
    pp.n_groups = a.n_groups;
    pp.out = reinterpret_cast<const __nv_bfloat16*>(a.out);
    pp.dout = reinterpret_cast<const __nv_bfloat16*>(a.dout);
    pp.lse = a.lse;
    pp.dqkv = reinterpret_cast<__nv_bfloat16*>(a.dqkv);
    pp.positions = a.positions;
    pp.rope_table = reinterpret_cast<const float2*>(a.rope_table);
    pp.total_tokens = a.total_tokens;
    pp.hidden = static_cast<int>(H);
    pp.scale = 0.125f;
    pp.scale_log2 = 0.125f * 1.4426950408889634f;
    dim3 grid(a.max_groups, a.heads, 1);
    packed::attn_bwd_packed_kernel<<<grid, THREADS, packed::P_SMEM_BYTES, stream>>>(qkv128, do128, pp);
    CM3P_CUDA_TRY(cudaGetLastError());
    return kOk;
  }
  // sequences longer than one tile: the streaming v3 kernels (attn_bwd_v3_sm100.cu).  Short sequences stay on the
  // light 2-CTA/SM kernels below.
  if (!get_option(kOptAttnForceTileKernels) && a.max_seqlen > BT) {
    // sliding-window layers (|i - j| <= 64): the fused band walk; global layers: the streaming two-kernel backward
    if (a.window >= 0 && a.window <= 64 && get_option(kOptAttnWindowWalk)) return attn_varlen_bwd_window(a, stream);
    return attn_varlen_bwd_v3(a, stream);
  }
  CUtensorMap qkv128, qkv64, do128, do64;
  if ((rc = encode_tmap_2d_bf16(&qkv128, a.qkv, 3 * H, T, 3 * H * 2, 64, BT)) != kOk) return rc;
  if ((rc = encode_tmap_2d_bf16(&qkv64, a.qkv, 3 * H, T, 3 * H * 2, 64, BI)) != kOk) return rc;
  if ((rc = encode_tmap_2d_bf16(&do128, a.dout, H, T, H * 2, 64, BT)) != kOk) return rc;
  if ((rc = encode_tmap_2d_bf16(&do64, a.dout, H, T, H * 2, 64, BI)) != kOk) return rc;
  CM3P_ENSURE_DYN_SMEM(attn_bwd_dq_kernel, DQ_SMEM_BYTES);
  CM3P_ENSURE_DYN_SMEM(attn_bwd_dkv_kernel, DKV_SMEM_BYTES);
  BwdParams p;
  p.cu_seqlens = a.cu_seqlens;
  p.out = reinterpret_cast<const __nv_bfloat16*>(a.out);
  p.dout = reinterpret_cast<const __nv_bfloat16*>(a.dout);
  p.lse = a.lse;
  p.delta = a.delta;
  p.dqkv = reinterpret_cast<__nv_bfloat16*>(a.dqkv);
  p.positions = a.positions;
  p.rope_table = reinterpret_cast<const float2*>(a.rope_table);
  p.total_tokens = a.total_tokens;
  p.heads = a.heads;
  p.hidden = static_cast<int>(H);
  p.window = a.window;
  p.scale = 0.125f;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.ctas_per_seq = (a.max_seqlen + BT - 1) / BT;
  CM3P_REQUIRE(static_cast<int64_t>(p.ctas_per_seq) * a.batch <= 0x7fffffffLL && a.heads <= 65535, kBadShape,
               "attn_bwd: grid too large (batch=%d max_seqlen=%d heads=%d)", a.batch, a.max_seqlen, a.heads);
  dim3 grid(static_cast<unsigned>(p.ctas_per_seq) * a.batch, a.heads, 1);
  attn_bwd_dq_kernel<<<grid, THREADS, DQ_SMEM_BYTES, stream>>>(qkv128, qkv64, do128, p);
  CM3P_CUDA_TRY(cudaGetLastError());
  attn_bwd_dkv_kernel<<<grid, THREADS, DKV_SMEM_BYTES, stream>>>(qkv128, qkv64, do64, p);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
