// Attention backward, third generation (global layers over long sequences).  Same math and data layout
// as attn_bwd_sm100.cu (formulas and reference call sites are in its header); what changed is how the
// work is cut so that neither the tensor pipe nor the element-wise warps wait for each other:
//
//   * 128 x 128 tiles.  Every MMA is M128 x N128 (S, dP) or M128 x N64 over K = 128: with 64-wide inner
//     tiles the 128-row operand was re-read from shared memory for every 64 columns (192 B/clk, above
//     the 128 B/clk the SM delivers).
//   * 16 element-wise warps share ONE tile: warp w owns TMEM lanes 32 (w & 3) .. +31 and the 32-column
//     slice (w >> 2).  Four light warps per scheduler (about 100 registers each) hide the exp / TMEM /
//     mbarrier latencies that two heavy ones could not (measured: 2 warps per scheduler ran the
//     element-wise loop at 0.2 instructions per clock).  They copy their S and dP values into registers
//     and release the TMEM buffers at once (s_free), so the S / dP GEMMs of the next tile run while
//     exp / dZ of this tile are computed.
//   * dZ (and P^T) never touch shared memory: the warps write them as packed bf16 into TMEM (tcgen05.st,
//     thread = row = TMEM lane) and the accumulating GEMMs read their A operand from TMEM.  With both
//     operands in shared memory those N = 64 MMAs needed 192 B/clk and the dK/dV kernel moved 256 KB of
//     shared memory per tile (2048 clk at the SM's 128 B/clk).  They are handed over in the 32-column
//     slices, each with its own mbarrier pair: the GEMM on a slice is issued as soon as it is written,
//     and the slice is reusable by the time its warps get to it in the next tile.
//   * two issuer warps on different schedulers (S / dP of the next tile; the accumulating GEMMs on the
//     slices), each running its loop warp-uniformly with one elected lane issuing; all issuer / producer
//     waits are blocking mbarrier waits (a polling issuer steals issue slots from the element-wise warps
//     on its scheduler); packed f32x2 arithmetic in the element-wise loops.
//
// Two deterministic kernels as before (no atomics): dQ (outer = 128 queries, streams K/V) and dK/dV on
// the transposed problem (outer = 128 keys as TMEM lanes, streams Q/dO with their lse / delta rows).
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "attn.h"
#include "common.h"
#include "ptx.cuh"

namespace cm3p {
namespace {
namespace v3 {

constexpr int BT = 128;  // outer tile rows == TMEM lanes
constexpr int BI = 128;  // inner (streamed) tile rows
constexpr int HC = 32;   // columns per element-wise warp
constexpr int D = 64;
constexpr int TILE_BYTES = 128 * D * 2;  // 16 KB: one 128-row operand tile, or one 64-column block of dZ / P^T
constexpr int NS = 3;                    // ring stages of the streamed operands
constexpr int THREADS = 640;  // 16 element-wise warps, producer, issuer, 2 idle (warpgroup granularity)
constexpr int EW_WARPS = 16;
constexpr int REG_EW = 104, REG_AUX = 64;  // known: at 104 the dKV loop keeps t_s and the tile counter in local memory (two reloads per inner tile); 112 / 64 = all 65536 registers removes them but never starts (setmaxnreg.inc needs headroom)  // (112 / 64 = all 65536 registers removes the dKV loop's two spilled values but never starts: setmaxnreg.inc needs headroom)

struct BwdParams {
  const int32_t* cu_seqlens;
  const __nv_bfloat16* out;
  const __nv_bfloat16* dout;
  const float* lse;
  float* delta;
  __nv_bfloat16* dqkv;
  const int32_t* positions;
  const float2* rope_table;
  int64_t total_tokens;
  int heads;
  int hidden;
  int window;
  float scale_log2;
  float scale;
  int outer_per_cta;
  int ctas_per_seq;  // grid.x = batch * ctas_per_seq (grid.z stops at 65535 sequences)
};

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
// Allowed columns [a,b) of one row inside a 32-column slice whose column 0 sits at sequence position t0,
// and the state of the slice for the row's warp (0 skip, 1 mask per element, 2 no mask).
struct Band {
  int a, b;
  int state;
};
__device__ __forceinline__ Band band_of(int row_pos, int warp_pos, int t0, int len, int window, bool row_valid) {
  Band r;
  int a = 0, b = max(0, min(HC, len - t0));
  int wa = 0, wb = b, ia = 0, ib = b;
  if (window >= 0) {
    a = max(a, row_pos - window - t0);
    b = min(b, row_pos + window + 1 - t0);
    wa = max(wa, warp_pos - window - t0);
    wb = min(wb, warp_pos + 31 + window + 1 - t0);
    ia = max(ia, warp_pos + 31 - window - t0);
    ib = min(ib, warp_pos + window + 1 - t0);
  }
  if (warp_pos + 31 >= len) { ia = HC; ib = 0; }  // rows past the end of the sequence: never "fully allowed"
  if (warp_pos >= len) { wa = HC; wb = 0; }
  if (!row_valid) b = 0;
  r.a = a;
  r.b = b;
  r.state = (HC <= wa || 0 >= wb) ? 0 : ((0 >= ia && HC <= ib) ? 2 : 1);
  return r;
}

// Both kernels walk OUTER_PER_CTA consecutive outer tiles of one (sequence, head) in a single CTA and treat
// the (outer, inner) tile pairs as ONE stream: the producer runs ahead into the next outer tile (its 128-row
// operands are double-buffered), the issuer chains the S / dP GEMM of the next outer tile's first inner tile
// behind the current one, and the element-wise warps only swap their per-row constants.  A CTA per outer
// tile paid ~9000 clk of launch / TMEM allocation / first-load / drain latency with nothing to overlap it
// (1 CTA per SM): 35 % of a global-layer CTA and 70 % of a window-layer one (measured).
// The count per CTA is chosen by the launcher (BwdParams::outer_per_cta): enough CTAs for ~16 (global) / ~4
// (window) waves of uneven work, at most MAX_OUTER_PER_CTA tiles each.
constexpr int MAX_OUTER_PER_CTA = 16;

struct TileRange {
  int base;  // sequence position of inner tile 0
  int n;     // inner tiles
};
// inner tiles start AT the band (not on a 128 grid): a window layer streams 2 tiles per outer tile, not 3
__device__ __forceinline__ TileRange tile_range(int o0, int len, int window) {
  int lo = 0, hi = len - 1;
  if (window >= 0) {
    lo = max(0, o0 - window);
    hi = min(len - 1, o0 + BT - 1 + window);
  }
  TileRange r;
  r.base = lo;
  r.n = (hi - lo) / BI + 1;
  return r;
}

// delta[head, row] = <dO[row, head, :], O[row, head, :]>: 8 lanes per (row, head), one 16-byte load of each
// operand per lane (HBM-bound pre-pass; both backward kernels read the result).
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                      float* __restrict__ delta, int64_t total_tokens, int heads) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;  // 16-byte unit of the [T, heads*64] matrix
  const int64_t pair = idx >> 3;                                              // row * heads + head
  const bool ok = pair < total_tokens * heads;
  float acc = 0.f;
  if (ok) {
    float a[8], b[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(out) + idx), a);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(dout) + idx), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += a[k] * b[k];
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  if (ok && (threadIdx.x & 7) == 0) {
    const int64_t row = pair / heads;
    const int head = static_cast<int>(pair - row * heads);
    delta[static_cast<int64_t>(head) * total_tokens + row] = acc;
  }
}

// ================================================================================================
// dQ kernel.  smem: Q 2x16K | dO 2x16K | K 3x16K | V 3x16K | per-row -lse / -delta*scale 2x1 KB.
// TMEM: S = [0,128)  dP = [128,256)  dZ (bf16 pairs) = [256,320)  dQ[2] = [320,448) (double-buffered so the
// next outer tile accumulates while the previous one is written out).
constexpr int DQ_TILES = 4 * TILE_BYTES + NS * 2 * TILE_BYTES;  // 160 KB
constexpr int DQ_VEC = 2 * 2 * BT * 4;                          // 2 KB
constexpr int DQ_SMEM = DQ_TILES + DQ_VEC + 256;

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dq_v3_kernel(const __grid_constant__ CUtensorMap tma_qkv128, const __grid_constant__ CUtensorMap tma_do128,
                      const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.ctas_per_seq, head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int o_begin = (blockIdx.x % p.ctas_per_seq) * p.outer_per_cta;
  const int n_o = min(p.outer_per_cta, (len + BT - 1) / BT - o_begin);  // outer tiles of this CTA
  if (n_o <= 0) return;

  uint8_t* smem_q = smem;                       // [2][16K]
  uint8_t* smem_do = smem + 2 * TILE_BYTES;     // [2][16K]
  uint8_t* smem_k = smem + 4 * TILE_BYTES;      // [NS][16K]
  uint8_t* smem_v = smem_k + NS * TILE_BYTES;   // [NS][16K]
  float* smem_vec = reinterpret_cast<float*>(smem + DQ_TILES);  // [2][-lse 128 | -delta*scale 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DQ_TILES + DQ_VEC);
  uint64_t* q_full = bars;                 // [2] TMA bytes + 32 staging-lane arrivals
  uint64_t* q_empty = q_full + 2;          // [2] the S / dP MMAs of the outer tile have retired and its row constants are read
  uint64_t* kv_full = q_empty + 2;         // [NS]
  uint64_t* kv_empty = kv_full + NS;       // [NS]
  uint64_t* s_full = kv_empty + NS;        // S and dP of a tile are in TMEM
  uint64_t* s_free = s_full + 1;           // 16 arrivals (one per warp): S / dP are in registers
  uint64_t* dz_full = s_free + 1;          // [slice] 128 arrivals: a 32-column slice of dZ is in TMEM
  uint64_t* dz_free = dz_full + 4;         // [slice] the dQ MMAs that read the slice have retired
  uint64_t* acc_full = dz_free + 4;        // [2]
  uint64_t* acc_free = acc_full + 2;       // [2] 16 arrivals: the accumulator has been copied out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 2);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == EW_WARPS + 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 33);
      ptx::mbar_init(&q_empty[i], 1 + EW_WARPS);  // issuer commit (Q / dO) + every element-wise warp (row constants)
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_free[i], EW_WARPS);
    }
    for (int s = 0; s < NS; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 2);  // S / dP issuer + dQ issuer
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_free, EW_WARPS);
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&dz_full[i], 128);
      ptx::mbar_init(&dz_free[i], 1);
    }
    ptx::fence_barrier_init();
  }
  if (warp == EW_WARPS) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_qkv128);
      ptx::prefetch_tmap(&tma_do128);
    }
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_S = 0, TM_DP = 128, TM_DZ = 256, TM_DQ = 320;

  if (warp >= EW_WARPS) {
    ptx::setmaxnreg_dec<REG_AUX>();
    if (warp == EW_WARPS) {
      // ---------------------------------------------------------------- producer (whole warp)
      const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
      const int64_t vec_off = static_cast<int64_t>(head) * p.total_tokens;
      int it = 0;
      for (int oi = 0; oi < n_o; ++oi) {
        const int q0 = (o_begin + oi) * BT;
        const int qb = oi & 1;
        ptx::mbar_wait(&q_empty[qb], ((oi >> 1) & 1) ^ 1);
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(&q_full[qb], 2 * TILE_BYTES);
          ptx::tma_load_2d(smem_q + qb * TILE_BYTES, &tma_qkv128, &q_full[qb], col_q, seq_start + q0);
          ptx::tma_load_2d(smem_do + qb * TILE_BYTES, &tma_do128, &q_full[qb], head * D, seq_start + q0);
        }
        // per-row constants of the outer tile (delta = <dO, O> comes from attn_bwd_delta_kernel)
        float* vec = smem_vec + qb * 2 * BT;
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
          const int r = q0 + rr * 32 + lane;
          const int64_t row = static_cast<int64_t>(seq_start) + r;
          const bool ok = r < len;
          vec[rr * 32 + lane] = ok ? -p.lse[vec_off + row] : 0.f;
          vec[BT + rr * 32 + lane] = ok ? -p.delta[vec_off + row] * p.scale : 0.f;  // dZ = P * (dP*scale - delta*scale)
        }
        ptx::mbar_arrive(&q_full[qb]);
        const TileRange tr = tile_range(q0, len, p.window);
        for (int j = 0; j < tr.n; ++j, ++it) {
          const int s = it % NS;
          ptx::mbar_wait(&kv_empty[s], ((it / NS) & 1) ^ 1);
          if (lane == 0) {
            ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
            const int row = seq_start + tr.base + j * BI;
            ptx::tma_load_2d(smem_k + s * TILE_BYTES, &tma_qkv128, &kv_full[s], col_k, row);
            ptx::tma_load_2d(smem_v + s * TILE_BYTES, &tma_qkv128, &kv_full[s], col_v, row);
          }
        }
      }
    } else if (warp == EW_WARPS + 1) {
      // ---------------------------------------------------------------- MMA issuer
      // The whole warp runs the issuer loop (warp-uniform control flow keeps the descriptor arithmetic in the
      // uniform datapath); one elected lane issues.  With a single active lane every MMA cost ~100 clk of
      // address moves and the issuer, not the tensor pipe, set the pace (measured).
      const bool leader = ptx::elect_one();
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BT, BI, 0, 0);
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      // S / dP of global tile t (stage t % NS) from the Q / dO buffer qb; `last` = last inner tile of its outer tile
      auto issue_s_dp = [&](int t, int qb, bool last) {
        const int s = t % NS;
        ptx::mbar_wait(&kv_full[s], (t / NS) & 1);
        ptx::tc_fence_after();
        const uint32_t q_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_q + qb * TILE_BYTES), 16);
        const uint32_t do_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_do + qb * TILE_BYTES), 16);
        const uint32_t k_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_k + s * TILE_BYTES), 16);
        const uint32_t v_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_v + s * TILE_BYTES), 16);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          if (leader) ptx::umma_bf16_split(tmem_base + TM_S, q_lo + k * 2, k_lo + k * 2, HI, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          if (leader) ptx::umma_bf16_split(tmem_base + TM_DP, do_lo + k * 2, v_lo + k * 2, HI, idesc_s, k != 0 ? 1u : 0u);
        if (leader) {
          ptx::umma_commit(s_full);
          ptx::umma_commit(&kv_empty[s]);            // one of the two releases of the K / V stage (the dQ issuer: the other)
          if (last) ptx::umma_commit(&q_empty[qb]);  // Q / dO of this outer tile are not read again
        }
      };
      int it = 0;
      ptx::mbar_wait(&q_full[0], 0);
      issue_s_dp(0, 0, tile_range(o_begin * BT, len, p.window).n == 1);
#pragma unroll 1
      for (int oi = 0; oi < n_o; ++oi) {
        const int n = tile_range((o_begin + oi) * BT, len, p.window).n;
        const int qb = oi & 1;
#pragma unroll 1
        for (int j = 0; j < n; ++j, ++it) {
          // chain the S / dP GEMM of the next tile of the stream (same outer tile, or the next one's first)
          if (j + 1 < n) {
            ptx::mbar_wait(s_free, it & 1);  // S / dP of tile `it` are in registers
            issue_s_dp(it + 1, qb, j + 2 == n);
          } else if (oi + 1 < n_o) {
            ptx::mbar_wait(&q_full[qb ^ 1], ((oi + 1) >> 1) & 1);
            ptx::mbar_wait(s_free, it & 1);
            issue_s_dp(it + 1, qb ^ 1, tile_range((o_begin + oi + 1) * BT, len, p.window).n == 1);
          }
        }
      }
    } else if (warp == EW_WARPS + 2) {
      // ---------------------------------------------------------------- second issuer: dQ += dZ K
      // A warp of its own (on another scheduler than the S / dP issuer): the S / dP GEMM of the next tile is
      // never queued behind the wait for this tile's dZ slices, and the issuing work is spread over two
      // schedulers (with one issuer the four element-wise warps sharing its scheduler ran 20 % behind the
      // other twelve and paced every slice hand-over).  The two issuers write disjoint TMEM regions.
      const bool leader = ptx::elect_one();
      const uint32_t idesc_dq = ptx::umma_idesc_bf16(BT, D, 0, 1);  // K as MN-major B operand
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      int it = 0;
#pragma unroll 1
      for (int oi = 0; oi < n_o; ++oi) {
        const int n = tile_range((o_begin + oi) * BT, len, p.window).n;
        const int qb = oi & 1;
        const uint32_t t_acc = tmem_base + TM_DQ + qb * 64;
        // the accumulator buffer was last used two outer tiles ago: wait until it is copied out
        ptx::mbar_wait(&acc_free[qb], ((oi >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int j = 0; j < n; ++j, ++it) {
          const int s = it % NS;
          ptx::mbar_wait(&kv_full[s], (it / NS) & 1);  // already complete (S of this tile was computed from it)
          // K tile as the MN-major B operand of dQ += dZ K (LBO = 8192: distance of 64-element MN chunks, unused)
          const uint32_t kmn_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_k + s * TILE_BYTES), 8192);
#pragma unroll
          for (int c = 0; c < 4; ++c) {  // 32-column slices (their warps finish at about the same time)
            ptx::mbar_wait(&dz_full[c], it & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = 2 * c + kk;  // 16-key step: 8 TMEM columns of dZ, 16 rows of K
              if (leader)
                ptx::umma_bf16_ts(t_acc, tmem_base + TM_DZ + k * 8, kmn_lo + ((k * 2048) >> 4), HI, idesc_dq,
                                  (j | c | kk) != 0 ? 1u : 0u);
            }
            if (leader) ptx::umma_commit(&dz_free[c]);
          }
          if (leader) {
            ptx::umma_commit(&kv_empty[s]);
            if (j + 1 == n) ptx::umma_commit(&acc_full[qb]);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ element-wise warps
    ptx::setmaxnreg_inc<REG_EW>();
    const int c = warp >> 2;                       // 32-column slice
    const int t = (warp & 3) * 32 + lane;          // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
    const float2 sc2 = make_float2(p.scale, p.scale);
    const uint32_t t_dz = tmem_base + TM_DZ + lane_off + c * (HC / 2);
    const uint32_t t_s = tmem_base + TM_S + lane_off + c * HC;
    const uint32_t t_dp = tmem_base + TM_DP + lane_off + c * HC;
    // dQ of outer tile `po`: every slice takes 8 + 8 columns (a RoPE pair is columns k and k + 32), so the
    // write-out costs all 16 warps the same short detour
    auto write_out = [&](int po) {
      const int ab = po & 1;
      ptx::mbar_wait(&acc_full[ab], (po >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t r1[8], r2[8];
      const uint32_t t_acc = tmem_base + TM_DQ + ab * 64 + lane_off + c * 8;
      ptx::tmem_ld_32x32b_x8(t_acc, r1);
      ptx::tmem_ld_32x32b_x8(t_acc + 32, r2);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_free[ab]);
      const int qi = (o_begin + po) * BT + t;
      if (qi >= len) return;
      const int64_t row = static_cast<int64_t>(seq_start) + qi;
      float o1[8], o2[8];
      if (p.rope_table && p.positions) {
        const float4* tab = reinterpret_cast<const float4*>(p.rope_table + static_cast<int64_t>(p.positions[row]) * 32) + c * 4;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float4 f = __ldg(tab + k);
          const float a0 = __uint_as_float(r1[2 * k]), b0 = __uint_as_float(r2[2 * k]);
          const float a1 = __uint_as_float(r1[2 * k + 1]), b1 = __uint_as_float(r2[2 * k + 1]);
          o1[2 * k] = a0 * f.x + b0 * f.y;
          o2[2 * k] = b0 * f.x - a0 * f.y;
          o1[2 * k + 1] = a1 * f.z + b1 * f.w;
          o2[2 * k + 1] = b1 * f.z - a1 * f.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          o1[k] = __uint_as_float(r1[k]);
          o2[k] = __uint_as_float(r2[k]);
        }
      }
      __nv_bfloat16* dst = p.dqkv + row * 3 * p.hidden + head * D + c * 8;
      *reinterpret_cast<uint4*>(dst) = pack8f(o1);
      *reinterpret_cast<uint4*>(dst + 32) = pack8f(o2);
    };
    int it = 0;
#ifdef CM3P_ATTN_PROF
    const long long pf0 = clock64();
    __shared__ long long pf_ev[MAX_OUTER_PER_CTA];
#endif
    for (int oi = 0; oi < n_o; ++oi) {
#ifdef CM3P_ATTN_PROF
      if (threadIdx.x == 0) pf_ev[oi] = clock64();
#endif
      const int q0 = (o_begin + oi) * BT;
      const int qi = q0 + t;
      const bool valid = qi < len;
      const int warp_pos = q0 + (warp & 3) * 32;
      const TileRange tr = tile_range(q0, len, p.window);
      ptx::mbar_wait(&q_full[oi & 1], (oi >> 1) & 1);
      const float nlse = smem_vec[(oi & 1) * 2 * BT + t];
      const float ndsc = smem_vec[(oi & 1) * 2 * BT + BT + t];
      // the producer may refill this buffer two outer tiles ahead: with a single inner tile per outer tile
      // (tiny windows) the S / dP GEMM alone could release it before these reads
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&q_empty[oi & 1]);
      const float2 nlse2 = make_float2(nlse, nlse), ndsc2 = make_float2(ndsc, ndsc);
      for (int j = 0; j < tr.n; ++j, ++it) {
        const int kv0 = tr.base + j * BI + c * HC;
        ptx::mbar_wait(s_full, it & 1);
        ptx::tc_fence_after();
        uint32_t rs[32], rp[32];
        ptx::tmem_ld_32x32b_x32(t_s, rs);
        ptx::tmem_ld_32x32b_x32(t_dp, rp);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(s_free);
        const Band bd = band_of(qi, warp_pos, kv0, len, p.window, valid);
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) packed[i] = 0u;
        if (bd.state == 2) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 e = ptx::fma2(make_float2(__uint_as_float(rs[i]), __uint_as_float(rs[i + 1])), c2, nlse2);
            const float2 pr = ptx::ex2_pair(e, i >> 1);
            const float2 u = ptx::fma2(make_float2(__uint_as_float(rp[i]), __uint_as_float(rp[i + 1])), sc2, ndsc2);
            const float2 z = ptx::mul2(pr, u);
            packed[i >> 1] = ptx::pack_bf16x2(z.x, z.y);
          }
        } else if (bd.state == 1) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 e = ptx::fma2(make_float2(__uint_as_float(rs[i]), __uint_as_float(rs[i + 1])), c2, nlse2);
            const float p0 = (i >= bd.a && i < bd.b) ? ptx::ex2_approx(e.x) : 0.f;
            const float p1 = (i + 1 >= bd.a && i + 1 < bd.b) ? ptx::ex2_approx(e.y) : 0.f;
            const float2 u = ptx::fma2(make_float2(__uint_as_float(rp[i]), __uint_as_float(rp[i + 1])), sc2, ndsc2);
            packed[i >> 1] = ptx::pack_bf16x2(p0 * u.x, p1 * u.y);
          }
        }
        // the slice still feeds the dQ MMAs of the previous tile until dz_free fires (issued a whole tile ago)
        if (it > 0) ptx::mbar_wait(&dz_free[c], (it - 1) & 1);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32b_x16(t_dz, packed);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&dz_full[c]);
        // the previous outer tile's accumulator is written out one tile late: its MMAs have drained by now
        if (j == 0 && oi > 0) write_out(oi - 1);
      }
    }
    write_out(n_o - 1);
#ifdef CM3P_ATTN_PROF
    if (threadIdx.x == 0 && blockIdx.y == 0 && blockIdx.x < p.ctas_per_seq) {
      const long long pf1 = clock64();
      for (int oi = 0; oi < n_o; ++oi) printf("dq cta %d outer %d start +%lld\n", blockIdx.x, oi, pf_ev[oi] - pf0);
      printf("dq cta %d end +%lld tiles=%d\n", blockIdx.x, pf1 - pf0, it);
    }
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EW_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ================================================================================================
// dKV kernel.  smem: K 2x16K | V 2x16K | Q 3x16K | dO 3x16K | -lse / -delta*scale 3x1 KB
// TMEM: S^T = [0,128)  dP^T = [128,256)  P^T (bf16 pairs) = [256,320)  dZ^T = [320,384)  dK = [384,448)  dV = [448,512).
constexpr int DKV_TILES = 4 * TILE_BYTES + NS * 2 * TILE_BYTES;  // 160 KB
constexpr int DKV_VEC = NS * 2 * BI * 4;                         // 3 KB
constexpr int DKV_SMEM = DKV_TILES + DKV_VEC + 256;

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dkv_v3_kernel(const __grid_constant__ CUtensorMap tma_qkv128, const __grid_constant__ CUtensorMap tma_do128,
                       const BwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.ctas_per_seq, head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int o_begin = (blockIdx.x % p.ctas_per_seq) * p.outer_per_cta;
  const int n_o = min(p.outer_per_cta, (len + BT - 1) / BT - o_begin);
  if (n_o <= 0) return;

  uint8_t* smem_k = smem;                       // [2][16K]
  uint8_t* smem_v = smem + 2 * TILE_BYTES;      // [2][16K]
  uint8_t* smem_q = smem + 4 * TILE_BYTES;      // [NS][16K]
  uint8_t* smem_do = smem_q + NS * TILE_BYTES;  // [NS][16K]
  float* smem_vec = reinterpret_cast<float*>(smem + DKV_TILES);  // [stage][-lse 128 | -delta*scale 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + DKV_TILES + DKV_VEC);
  uint64_t* kv_full = bars;               // [2]
  uint64_t* kv_empty = kv_full + 2;       // [2] the S^T / dP^T MMAs of the outer tile have retired
  uint64_t* qdo_full = kv_empty + 2;      // [NS] TMA bytes + 32 staging-lane arrivals
  uint64_t* qdo_empty = qdo_full + NS;    // [NS]
  uint64_t* s_full = qdo_empty + NS;
  uint64_t* s_free = s_full + 1;          // 16 arrivals
  uint64_t* pz_full = s_free + 1;         // [slice] 128 arrivals: 32-column slices of P^T and dZ^T are in TMEM
  uint64_t* pz_free = pz_full + 4;        // [slice]
  uint64_t* acc_full = pz_free + 4;
  uint64_t* acc_free = acc_full + 1;      // 16 arrivals: dK and dV have been copied out
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_free + 1);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == EW_WARPS + 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    for (int s = 0; s < NS; ++s) {
      ptx::mbar_init(&qdo_full[s], 33);
      ptx::mbar_init(&qdo_empty[s], 2);  // S^T / dP^T issuer + dK / dV issuer
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_free, EW_WARPS);
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&pz_full[i], 128);
      ptx::mbar_init(&pz_free[i], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_free, EW_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp == EW_WARPS) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_qkv128);
      ptx::prefetch_tmap(&tma_do128);
    }
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_PT = 256, TM_DZT = 320, TM_DK = 384, TM_DV = 448;

  if (warp >= EW_WARPS) {
    ptx::setmaxnreg_dec<REG_AUX>();
    if (warp == EW_WARPS) {
      // ------------------------------------------------------------------ producer (whole warp)
      const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
      const float* lse_h = p.lse + static_cast<int64_t>(head) * p.total_tokens;
      const float* delta_h = p.delta + static_cast<int64_t>(head) * p.total_tokens;
      int it = 0;
      for (int oi = 0; oi < n_o; ++oi) {
        const int k0 = (o_begin + oi) * BT;
        const int kb = oi & 1;
        ptx::mbar_wait(&kv_empty[kb], ((oi >> 1) & 1) ^ 1);
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(&kv_full[kb], 2 * TILE_BYTES);
          ptx::tma_load_2d(smem_k + kb * TILE_BYTES, &tma_qkv128, &kv_full[kb], col_k, seq_start + k0);
          ptx::tma_load_2d(smem_v + kb * TILE_BYTES, &tma_qkv128, &kv_full[kb], col_v, seq_start + k0);
        }
        const TileRange tr = tile_range(k0, len, p.window);
        for (int i = 0; i < tr.n; ++i, ++it) {
          const int s = it % NS;
          ptx::mbar_wait(&qdo_empty[s], ((it / NS) & 1) ^ 1);
          const int64_t row = static_cast<int64_t>(seq_start) + tr.base + i * BI;
          if (lane == 0) {
            ptx::mbar_arrive_expect_tx(&qdo_full[s], 2 * TILE_BYTES);
            ptx::tma_load_2d(smem_q + s * TILE_BYTES, &tma_qkv128, &qdo_full[s], col_q, static_cast<int32_t>(row));
            ptx::tma_load_2d(smem_do + s * TILE_BYTES, &tma_do128, &qdo_full[s], col_q, static_cast<int32_t>(row));
          }
          float* vec = smem_vec + s * 2 * BI;
#pragma unroll
          for (int hh = 0; hh < BI; hh += 32) {
            const int64_t r = row + hh + lane;
            const bool ok = r < p.total_tokens;
            vec[hh + lane] = ok ? -lse_h[r] : 0.f;
            vec[BI + hh + lane] = ok ? -delta_h[r] * p.scale : 0.f;  // dZ = P * (dP*scale - delta*scale)
          }
          ptx::mbar_arrive(&qdo_full[s]);
        }
      }
    } else if (warp == EW_WARPS + 1) {
      // ------------------------------------------------------------------ MMA issuer (see the dQ kernel)
      const bool leader = ptx::elect_one();
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BT, BI, 0, 0);
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      auto issue_s_dp = [&](int t, int kb, bool last) {
        const int s = t % NS;
        ptx::mbar_wait(&qdo_full[s], (t / NS) & 1);
        ptx::tc_fence_after();
        const uint32_t k_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_k + kb * TILE_BYTES), 16);
        const uint32_t v_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_v + kb * TILE_BYTES), 16);
        const uint32_t q_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_q + s * TILE_BYTES), 16);
        const uint32_t do_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_do + s * TILE_BYTES), 16);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          if (leader) ptx::umma_bf16_split(tmem_base + TM_ST, k_lo + k * 2, q_lo + k * 2, HI, idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          if (leader) ptx::umma_bf16_split(tmem_base + TM_DPT, v_lo + k * 2, do_lo + k * 2, HI, idesc_s, k != 0 ? 1u : 0u);
        if (leader) {
          ptx::umma_commit(s_full);
          ptx::umma_commit(&qdo_empty[s]);            // one of the two releases of the Q / dO stage
          if (last) ptx::umma_commit(&kv_empty[kb]);  // K / V of this outer tile are not read again
        }
      };
      int it = 0;
      ptx::mbar_wait(&kv_full[0], 0);
      issue_s_dp(0, 0, tile_range(o_begin * BT, len, p.window).n == 1);
#pragma unroll 1
      for (int oi = 0; oi < n_o; ++oi) {
        const int n = tile_range((o_begin + oi) * BT, len, p.window).n;
        const int kb = oi & 1;
#pragma unroll 1
        for (int i = 0; i < n; ++i, ++it) {
          if (i + 1 < n) {
            ptx::mbar_wait(s_free, it & 1);
            issue_s_dp(it + 1, kb, i + 2 == n);
          } else if (oi + 1 < n_o) {
            ptx::mbar_wait(&kv_full[kb ^ 1], ((oi + 1) >> 1) & 1);
            ptx::mbar_wait(s_free, it & 1);
            issue_s_dp(it + 1, kb ^ 1, tile_range((o_begin + oi + 1) * BT, len, p.window).n == 1);
          }
        }
      }
    } else if (warp == EW_WARPS + 2) {
      // ------------------------------------------------------------------ second issuer: dV += P^T dO, dK += dZ^T Q
      // (see the dQ kernel: own warp on another scheduler, disjoint TMEM regions)
      const bool leader = ptx::elect_one();
      const uint32_t idesc_acc = ptx::umma_idesc_bf16(BT, D, 0, 1);
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      int it = 0;
#pragma unroll 1
      for (int oi = 0; oi < n_o; ++oi) {
        const int n = tile_range((o_begin + oi) * BT, len, p.window).n;
        if (oi > 0) {  // dK / dV of the previous outer tile must have been copied out
          ptx::mbar_wait(acc_free, (oi - 1) & 1);
          ptx::tc_fence_after();
        }
#pragma unroll 1
        for (int i = 0; i < n; ++i, ++it) {
          const int s = it % NS;
          ptx::mbar_wait(&qdo_full[s], (it / NS) & 1);  // already complete (S^T of this tile was computed from it)
          const uint32_t qmn_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_q + s * TILE_BYTES), 8192);
          const uint32_t domn_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_do + s * TILE_BYTES), 8192);
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            ptx::mbar_wait(&pz_full[c], it & 1);
            ptx::tc_fence_after();
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = 2 * c + kk;  // 16-query step: 8 TMEM columns of P^T, 16 rows of dO
              if (leader)
                ptx::umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_PT + k * 8, domn_lo + ((k * 2048) >> 4), HI,
                                  idesc_acc, (i | c | kk) != 0 ? 1u : 0u);
            }
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
              const int k = 2 * c + kk;
              if (leader)
                ptx::umma_bf16_ts(tmem_base + TM_DK, tmem_base + TM_DZT + k * 8, qmn_lo + ((k * 2048) >> 4), HI,
                                  idesc_acc, (i | c | kk) != 0 ? 1u : 0u);
            }
            if (leader) ptx::umma_commit(&pz_free[c]);
          }
          if (leader) {
            ptx::umma_commit(&qdo_empty[s]);
            if (i + 1 == n) ptx::umma_commit(acc_full);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ element-wise warps
    ptx::setmaxnreg_inc<REG_EW>();
    const int c = warp >> 2;                 // 32-column (query) slice
    const int t = (warp & 3) * 32 + lane;    // key row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const float2 c2 = make_float2(p.scale_log2, p.scale_log2);
    const float2 sc2 = make_float2(p.scale, p.scale);
    const uint32_t t_pt = tmem_base + TM_PT + lane_off + c * (HC / 2);
    const uint32_t t_dzt = tmem_base + TM_DZT + lane_off + c * (HC / 2);
    const uint32_t t_s = tmem_base + TM_ST + lane_off + c * HC;
    const uint32_t t_dp = tmem_base + TM_DPT + lane_off + c * HC;
    int it = 0;
    for (int oi = 0; oi < n_o; ++oi) {
      const int k0 = (o_begin + oi) * BT;
      const int kj = k0 + t;
      const bool valid = kj < len;
      const int warp_pos = k0 + (warp & 3) * 32;
      const TileRange tr = tile_range(k0, len, p.window);
      for (int i = 0; i < tr.n; ++i, ++it) {
        const int s = it % NS;
        const int q0i = tr.base + i * BI + c * HC;
        ptx::mbar_wait(s_full, it & 1);
        ptx::mbar_wait(&qdo_full[s], (it / NS) & 1);  // already complete: orders the lse / delta staging writes
        ptx::tc_fence_after();
        uint32_t rs[32], rp[32];
        ptx::tmem_ld_32x32b_x32(t_s, rs);
        ptx::tmem_ld_32x32b_x32(t_dp, rp);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(s_free);
        const Band bd = band_of(kj, warp_pos, q0i, len, p.window, valid);
        const float4* nlse4 = reinterpret_cast<const float4*>(smem_vec + s * 2 * BI + c * HC);
        const float4* ndel4 = reinterpret_cast<const float4*>(smem_vec + s * 2 * BI + BI + c * HC);
        uint32_t pp[16], pz[16];
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) pp[i2] = pz[i2] = 0u;
        if (bd.state == 2) {
          // every column allowed for every row of this warp: no predicates
#pragma unroll
          for (int i2 = 0; i2 < 32; i2 += 4) {
            const float4 l4 = nlse4[i2 >> 2];
            const float4 d4 = ndel4[i2 >> 2];
            const float2 e0 = ptx::fma2(make_float2(__uint_as_float(rs[i2]), __uint_as_float(rs[i2 + 1])), c2,
                                        make_float2(l4.x, l4.y));
            const float2 e1 = ptx::fma2(make_float2(__uint_as_float(rs[i2 + 2]), __uint_as_float(rs[i2 + 3])), c2,
                                        make_float2(l4.z, l4.w));
            const float2 p0 = ptx::ex2_pair(e0, i2 >> 1);
            const float2 p1 = ptx::ex2_pair(e1, (i2 >> 1) + 1);
            const float2 u0 = ptx::fma2(make_float2(__uint_as_float(rp[i2]), __uint_as_float(rp[i2 + 1])), sc2,
                                        make_float2(d4.x, d4.y));
            const float2 u1 = ptx::fma2(make_float2(__uint_as_float(rp[i2 + 2]), __uint_as_float(rp[i2 + 3])), sc2,
                                        make_float2(d4.z, d4.w));
            const float2 z0 = ptx::mul2(p0, u0), z1 = ptx::mul2(p1, u1);
            pp[i2 >> 1] = ptx::pack_bf16x2(p0.x, p0.y);
            pp[(i2 >> 1) + 1] = ptx::pack_bf16x2(p1.x, p1.y);
            pz[i2 >> 1] = ptx::pack_bf16x2(z0.x, z0.y);
            pz[(i2 >> 1) + 1] = ptx::pack_bf16x2(z1.x, z1.y);
          }
        } else if (bd.state == 1) {
          const float2* nlse2 = reinterpret_cast<const float2*>(nlse4);
          const float2* ndel2 = reinterpret_cast<const float2*>(ndel4);
#pragma unroll
          for (int i2 = 0; i2 < 32; i2 += 2) {
            const float2 l2 = nlse2[i2 >> 1];
            const float2 d2 = ndel2[i2 >> 1];
            const float2 e = ptx::fma2(make_float2(__uint_as_float(rs[i2]), __uint_as_float(rs[i2 + 1])), c2, l2);
            const float p0 = (i2 >= bd.a && i2 < bd.b) ? ptx::ex2_approx(e.x) : 0.f;
            const float p1 = (i2 + 1 >= bd.a && i2 + 1 < bd.b) ? ptx::ex2_approx(e.y) : 0.f;
            const float2 u = ptx::fma2(make_float2(__uint_as_float(rp[i2]), __uint_as_float(rp[i2 + 1])), sc2, d2);
            pp[i2 >> 1] = ptx::pack_bf16x2(p0, p1);
            pz[i2 >> 1] = ptx::pack_bf16x2(p0 * u.x, p1 * u.y);
          }
        }
        if (it > 0) ptx::mbar_wait(&pz_free[c], (it - 1) & 1);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32b_x16(t_pt, pp);
        ptx::tmem_st_32x32b_x16(t_dzt, pz);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(&pz_full[c]);
      }
      // dK (slices 0 / 1: 16 + 16 columns each, a RoPE pair is columns k and k + 32) and dV (slices 2 / 3: 32
      // columns each) of this outer tile, spread over all 16 warps so that no warp falls behind the others;
      // the accumulators are released as soon as they are in registers
      {
        ptx::mbar_wait(acc_full, oi & 1);
        ptx::tc_fence_after();
        uint32_t r1[16], r2[16];
        const uint32_t t_acc = tmem_base + lane_off + (c < 2 ? TM_DK + c * 16 : TM_DV + (c - 2) * 32);
        ptx::tmem_ld_32x32b_x16(t_acc, r1);
        ptx::tmem_ld_32x32b_x16(t_acc + (c < 2 ? 32 : 16), r2);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free);
        if (valid) {
          const int64_t row = static_cast<int64_t>(seq_start) + kj;
          __nv_bfloat16* base = p.dqkv + row * 3 * p.hidden + head * D;
          float o1[16], o2[16];
          if (c < 2 && p.rope_table && p.positions) {
            const float4* tab =
                reinterpret_cast<const float4*>(p.rope_table + static_cast<int64_t>(p.positions[row]) * 32) + c * 8;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float4 f = __ldg(tab + k);
              const float a0 = __uint_as_float(r1[2 * k]), b0 = __uint_as_float(r2[2 * k]);
              const float a1 = __uint_as_float(r1[2 * k + 1]), b1 = __uint_as_float(r2[2 * k + 1]);
              o1[2 * k] = a0 * f.x + b0 * f.y;
              o2[2 * k] = b0 * f.x - a0 * f.y;
              o1[2 * k + 1] = a1 * f.z + b1 * f.w;
              o2[2 * k + 1] = b1 * f.z - a1 * f.w;
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              o1[k] = __uint_as_float(r1[k]);
              o2[k] = __uint_as_float(r2[k]);
            }
          }
          // dK: columns c*16.. and 32 + c*16..; dV: columns (c-2)*32.. and (c-2)*32 + 16..
          __nv_bfloat16* d1 = base + (c < 2 ? p.hidden + c * 16 : 2 * p.hidden + (c - 2) * 32);
          __nv_bfloat16* d2 = d1 + (c < 2 ? 32 : 16);
          *reinterpret_cast<uint4*>(d1) = pack8f(o1);
          *reinterpret_cast<uint4*>(d1 + 8) = pack8f(o1 + 8);
          *reinterpret_cast<uint4*>(d2) = pack8f(o2);
          *reinterpret_cast<uint4*>(d2 + 8) = pack8f(o2 + 8);
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EW_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace v3
}  // namespace

int attn_varlen_bwd_v3(const AttnBwdArgs& a, cudaStream_t stream) {
  using namespace v3;
  const uint64_t H = static_cast<uint64_t>(a.heads) * 64;
  const uint64_t T = static_cast<uint64_t>(a.total_tokens);
  CUtensorMap qkv128, do128;
  int rc;
  if ((rc = encode_tmap_2d_bf16(&qkv128, a.qkv, 3 * H, T, 3 * H * 2, 64, BT)) != kOk) return rc;
  if ((rc = encode_tmap_2d_bf16(&do128, a.dout, H, T, H * 2, 64, BT)) != kOk) return rc;
  CM3P_ENSURE_DYN_SMEM(attn_bwd_dq_v3_kernel, DQ_SMEM);
  CM3P_ENSURE_DYN_SMEM(attn_bwd_dkv_v3_kernel, DKV_SMEM);
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
  const int64_t vec_units = a.total_tokens * a.heads * 8;
  attn_bwd_delta_kernel<<<static_cast<unsigned>((vec_units + 255) / 256), 256, 0, stream>>>(p.out, p.dout, p.delta,
                                                                                       a.total_tokens, a.heads);
  CM3P_CUDA_TRY(cudaGetLastError());
  const int forced_opc = get_option(kOptBwdOuterPerCta);  // 0 = heuristic (the tests sweep it)
  const int64_t units = (a.total_tokens / BT + a.batch / 2 + 1) * a.heads;  // ~ (sequence, head, outer tile) triples
  const int64_t target_ctas = static_cast<int64_t>(num_sms()) * (a.window >= 0 ? 4 : 16);
  int opc = static_cast<int>((units + target_ctas - 1) / target_ctas);
  opc = opc < 2 ? 2 : (opc > MAX_OUTER_PER_CTA ? MAX_OUTER_PER_CTA : opc);
  if (forced_opc > 0) opc = forced_opc > MAX_OUTER_PER_CTA ? MAX_OUTER_PER_CTA : forced_opc;
  p.outer_per_cta = opc;
  p.ctas_per_seq = (a.max_seqlen + BT * opc - 1) / (BT * opc);
  CM3P_REQUIRE(static_cast<int64_t>(p.ctas_per_seq) * a.batch <= 0x7fffffffLL && a.heads <= 65535, kBadShape,
               "attn_bwd: grid too large (batch=%d max_seqlen=%d heads=%d)", a.batch, a.max_seqlen, a.heads);
  dim3 grid(static_cast<unsigned>(p.ctas_per_seq) * a.batch, a.heads, 1);
  attn_bwd_dq_v3_kernel<<<grid, THREADS, DQ_SMEM, stream>>>(qkv128, do128, p);
  CM3P_CUDA_TRY(cudaGetLastError());
  attn_bwd_dkv_v3_kernel<<<grid, THREADS, DKV_SMEM, stream>>>(qkv128, do128, p);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
