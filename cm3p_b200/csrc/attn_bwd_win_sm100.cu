// Sliding-window attention backward as ONE kernel, 5 GEMMs and one exponentiation per (query tile, key tile) pair:
// the "band walk".
//
// ModernBERT's local layers attend to |i - j| <= 64 (14 of the 22 beatmap layers, 4 of the 6 audio layers;
// transformers/masking_utils.py:121-131, called from /root/reference/cm3p/modeling_cm3p.py:509-514,607-619).  With
// key tiles offset by -64 rows from the query tiles,
//     query tile i = rows [128 i, 128 i + 128)        key tile j = rows [128 j - 64, 128 j + 64)
// every query tile meets exactly two key tiles (j = i and j = i + 1) and every key tile exactly two query tiles
// (i = j - 1 and i = j): the score matrix is block bidiagonal.  A CTA walks down the band of one (sequence, head):
//     ... (Q_i, K_i) -> (Q_i, K_i+1) -> (Q_i+1, K_i+1) -> (Q_i+1, K_i+2) ...
// Consecutive pairs share either the query tile or the key tile, so each pair loads ONE new 32 KB operand pair, and
// the accumulators simply stay in TMEM across the two pairs they belong to: dQ_i over (Q_i, K_i), (Q_i, K_i+1);
// dK_j / dV_j over (Q_j-1, K_j), (Q_j, K_j).  Nothing is recomputed and nothing meets through atomics (the general
// two-kernel backward of attn_bwd_v3_sm100.cu recomputes S, P and dP in both kernels: 7 GEMMs and two exp passes over
// 256 columns for a 129-wide band).  Per pair, on the transposed tile (TMEM lanes = key rows) as in the packed kernel:
//     S^T = K Q^T, dP^T = V dO^T;  P^T = exp2(S^T c - lse_q) o band,  dZ^T = P^T o (dP^T/8 - delta_q/8)
//     dV += P^T dO (A from TMEM),  dK += dZ^T Q (A from smem, K-major),  dQ += dZ K (the same smem tile, MN-major)
// A CTA owns `tiles_per_cta` consecutive query tiles and the key tiles of the same indices; the pair above its first
// query tile is recomputed as a halo (only its dK / dV contribution), the pair below its last one only feeds dQ.
//
// Warps: 0-7 element-wise + epilogues (TMEM lane quadrant = warp & 3, 64-query column half = warp >> 2),
//        8 = TMA producer + lse / delta staging (+ TMEM alloc), 9 = MMA issuer.  1 CTA / SM (512 TMEM columns).
#include <cuda_bf16.h>
#include <math.h>

#include "attn.h"
#include "attn_bwd_common.cuh"
#include "common.h"
#include "ptx.cuh"

namespace cm3p {
namespace {
namespace win {
using namespace bwd_detail;

constexpr int BT = 128;   // tile rows (queries or keys)
constexpr int D = 64;
constexpr int KOFF = 64;  // key tile j starts at row 128 j - KOFF
constexpr int TILE = BT * D * 2;  // 16 KB
constexpr int EW_WARPS = 8;
constexpr int THREADS = (EW_WARPS + 2) * 32;
constexpr int SMEM_TILES = 8 * TILE + 2 * TILE;  // Q, dO, K, V double-buffered + dZ^T (2 blocks of 64 queries)
constexpr int VEC_BYTES = 2 * 2 * BT * 4;        // [buffer][lse | delta / 8][128 queries]
constexpr int CS_BYTES = 2 * BT * 256;             // rotation factors of one row per element-wise thread: [256][32 (cos, sin)] fp32
constexpr int SMEM_BYTES = SMEM_TILES + VEC_BYTES + CS_BYTES + 256;
constexpr uint32_t TM_ST = 0, TM_DPT = 128, TM_PT = 256, TM_DQ = 320, TM_DK = 384, TM_DV = 448;
constexpr int MAX_TILES_PER_CTA = 16;

struct WinParams {
  const int32_t* cu_seqlens;
  const float* lse;    // [heads, T] log2 domain
  const float* delta;  // [heads, T] <dO, O>
  __nv_bfloat16* dqkv;
  const int32_t* positions;
  const float2* rope_table;
  int64_t total_tokens;
  int heads;
  int hidden;
  int window;
  float scale_log2;
  float scale;
  int tiles_per_cta;
  int ctas_per_seq;
};

// One (query tile, key tile) pair of the walk and what happens to the accumulators around it (flags packed in one
// register: a struct of bools ended up in local memory, and with 1 CTA x 640 threads per SM the local-memory lines
// do not survive in the small L1 next to 227 KB of shared memory: every reload was an L2 round trip).
struct Part {
  int q, k;
  uint32_t f;
  // dQ_q: computed here / first pair of its accumulation / complete after this pair; dK_k, dV_k likewise
  __device__ __forceinline__ bool need_q() const { return f & 1u; }
  __device__ __forceinline__ bool q_first() const { return f & 2u; }
  __device__ __forceinline__ bool q_last() const { return f & 4u; }
  __device__ __forceinline__ bool need_kv() const { return f & 8u; }
  __device__ __forceinline__ bool kv_first() const { return f & 16u; }
  __device__ __forceinline__ bool kv_last() const { return f & 32u; }
};

struct Walk {
  int i0, i1, n_q, n_k;
  int step, phase;  // query tile, 0 = pair (step, step), 1 = pair (step, step + 1)
  __device__ __forceinline__ void init(int i0_, int i1_, int n_q_, int n_k_) {
    i0 = i0_; i1 = i1_; n_q = n_q_; n_k = n_k_;
    step = i0 > 0 ? i0 - 1 : 0;  // halo: the pair (Q_{i0-1}, K_{i0}) contributes to this CTA's first key tile
    phase = i0 > 0 ? 1 : 0;
  }
  __device__ __forceinline__ bool done() const { return step >= i1; }
  __device__ __forceinline__ Part get() const {
    Part p;
    p.q = step;
    p.k = step + phase;
    const bool halo = step < i0;
    bool need_kv, kv_first, kv_last;
    if (phase == 0) {
      need_kv = true;
      kv_first = step == 0;  // key tile 0 has no pair above it
      kv_last = true;
    } else {
      need_kv = halo || step + 1 < i1 || i1 == n_q;  // does this CTA own key tile step + 1 ?
      kv_first = true;
      kv_last = step + 1 >= n_q;  // the key tile below the last query tile has no second pair
    }
    p.f = (!halo ? 1u : 0u) | (phase == 0 ? 2u : 0u) | ((phase == 1 || step + 1 >= n_k) ? 4u : 0u) |
          (need_kv ? 8u : 0u) | (kv_first ? 16u : 0u) | (kv_last ? 32u : 0u);
    return p;
  }
  __device__ __forceinline__ void advance() {
    if (phase == 0 && step + 1 < n_k) {
      phase = 1;
    } else {
      phase = 0;
      ++step;
    }
  }
};

// One 32-query chunk of a pair for one key row (= TMEM lane): P^T = exp2(S^T c - lse_q), dZ^T = P^T o (dP^T / 8 -
// delta_q / 8), both packed to bf16.  `taddr` = this lane's row at the chunk's first column (S^T; dP^T sits TM_DPT
// columns further), `vec` = lse | delta / 8 of the chunk's queries in shared memory.  MASK: columns outside [a, b)
// are zeroed (chunks cut by the band edge or the end of the sequence); interior chunks skip the compares.
template <bool MASK>
__device__ __forceinline__ void chunk_math(uint32_t taddr, const float* vec, float c, float scale, int a, int b,
                                           uint32_t (&pp)[16], uint32_t (&pz)[16]) {
#pragma unroll
  for (int h = 0; h < 4; ++h) {  // 8 columns at a time: 16 live score registers next to the 32 packed results
    uint32_t rs[8], rp[8];
    ptx::tmem_ld_32x32b_x8(taddr + TM_ST + h * 8, rs);
    ptx::tmem_ld_32x32b_x8(taddr + TM_DPT + h * 8, rp);
    ptx::tmem_ld_wait();
    const float4* lse4 = reinterpret_cast<const float4*>(vec + h * 8);
    const float4* del4 = reinterpret_cast<const float4*>(vec + BT + h * 8);
#pragma unroll
    for (int i4 = 0; i4 < 8; i4 += 4) {
      const float4 l4 = lse4[i4 >> 2];
      const float4 d4 = del4[i4 >> 2];
      const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
      const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
      float pv[4], zv[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = h * 8 + i4 + e;
        const float ex = ptx::ex2_approx(__uint_as_float(rs[i4 + e]) * c - ls[e]);
        const float dz = ex * (__uint_as_float(rp[i4 + e]) * scale - dl[e]);
        const bool on = !MASK || (j >= a && j < b);
        pv[e] = on ? ex : 0.f;
        zv[e] = on ? dz : 0.f;
      }
      pp[h * 4 + (i4 >> 1)] = ptx::pack_bf16x2(pv[0], pv[1]);
      pp[h * 4 + (i4 >> 1) + 1] = ptx::pack_bf16x2(pv[2], pv[3]);
      pz[h * 4 + (i4 >> 1)] = ptx::pack_bf16x2(zv[0], zv[1]);
      pz[h * 4 + (i4 >> 1) + 1] = ptx::pack_bf16x2(zv[2], zv[3]);
    }
  }
}

__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_win_kernel(const __grid_constant__ CUtensorMap tma_qkv, const __grid_constant__ CUtensorMap tma_do,
                    const WinParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.ctas_per_seq, head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int n_q = (len + BT - 1) / BT;
  const int n_k = (len + KOFF + BT - 1) / BT;
  const int i0 = (blockIdx.x % p.ctas_per_seq) * p.tiles_per_cta;
  if (i0 >= n_q) return;  // also len == 0
  const int i1 = min(n_q, i0 + p.tiles_per_cta);

  uint8_t* smem_q = smem;                 // [2][16 KB]
  uint8_t* smem_do = smem + 2 * TILE;     // [2]
  uint8_t* smem_k = smem + 4 * TILE;      // [2]
  uint8_t* smem_v = smem + 6 * TILE;      // [2]
  uint8_t* smem_dzt = smem + 8 * TILE;    // [2 blocks of 64 queries][128 key rows][128 B]
  float* smem_vec = reinterpret_cast<float*>(smem + SMEM_TILES);  // [2][lse 128 | delta/8 128]
  uint8_t* smem_cs = smem + SMEM_TILES + VEC_BYTES;  // [128 rows][16 units of 16 B], unit u of row r at u ^ (r & 15)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_TILES + VEC_BYTES + CS_BYTES);
  uint64_t* q_full = bars;        // [2] TMA bytes + 32 staging-lane arrivals
  uint64_t* q_empty = bars + 2;   // [2]
  uint64_t* kv_full = bars + 4;   // [2]
  uint64_t* kv_empty = bars + 6;  // [2]
  uint64_t* s_full = bars + 8;    // S^T and dP^T of the pair are in TMEM
  uint64_t* pz_full = bars + 9;   // P^T (TMEM) and dZ^T (smem) written: 256 arrivals
  uint64_t* acc_done = bars + 10; // the accumulating MMAs of the pair have retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == EW_WARPS + 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 33);
      ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(pz_full, EW_WARPS * 32);
    ptx::mbar_init(acc_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == EW_WARPS) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tma_qkv);
      ptx::prefetch_tmap(&tma_do);
    }
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  Walk w;
  w.init(i0, i1, n_q, n_k);

  if (warp == EW_WARPS) {
    // ------------------------------------------------------------------ producer (whole warp: lse / delta staging)
    const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
    const float* lse_h = p.lse + static_cast<int64_t>(head) * p.total_tokens;
    const float* delta_h = p.delta + static_cast<int64_t>(head) * p.total_tokens;
    int prev_q = -1, prev_k = -1, qn = 0, kn = 0;
    for (; !w.done(); w.advance()) {
      const Part pt = w.get();
      if (pt.q != prev_q) {
        const int b = qn & 1;
        ptx::mbar_wait_trap(&q_empty[b], ((qn >> 1) & 1) ^ 1);
        const int row0 = seq_start + pt.q * BT;
        if (lane == 0) {
          ptx::mbar_arrive_expect_tx(&q_full[b], 2 * TILE);
          ptx::tma_load_2d(smem_q + b * TILE, &tma_qkv, &q_full[b], col_q, row0);
          ptx::tma_load_2d(smem_do + b * TILE, &tma_do, &q_full[b], col_q, row0);
        }
        float* vec = smem_vec + b * 2 * BT;
#pragma unroll
        for (int h = 0; h < BT; h += 32) {
          const int c = h + lane;
          const bool ok = pt.q * BT + c < len;
          vec[c] = ok ? __ldcg(lse_h + row0 + c) : 0.f;
          vec[BT + c] = ok ? __ldcg(delta_h + row0 + c) * p.scale : 0.f;
        }
        ptx::mbar_arrive(&q_full[b]);
        prev_q = pt.q;
        ++qn;
      }
      if (pt.k != prev_k) {
        const int b = kn & 1;
        ptx::mbar_wait_trap(&kv_empty[b], ((kn >> 1) & 1) ^ 1);
        if (lane == 0) {
          const int row0 = seq_start + pt.k * BT - KOFF;  // may be negative for the first tile: zero-filled / masked
          ptx::mbar_arrive_expect_tx(&kv_full[b], 2 * TILE);
          ptx::tma_load_2d(smem_k + b * TILE, &tma_qkv, &kv_full[b], col_k, row0);
          ptx::tma_load_2d(smem_v + b * TILE, &tma_qkv, &kv_full[b], col_v, row0);
        }
        prev_k = pt.k;
        ++kn;
      }
    }
  } else if (warp == EW_WARPS + 1) {
    // ------------------------------------------------------------------ MMA issuer
    // The whole warp runs the loop (warp-uniform control flow keeps the descriptor arithmetic in the uniform datapath:
    // a single diverged lane paid ~120 clk per MMA, 2850 clk per pair for the accumulating GEMMs alone); one elected
    // lane issues the MMAs and the commits.
    {
      const bool leader = ptx::elect_one();
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BT, BT, 0, 0);
      const uint32_t idesc_acc = ptx::umma_idesc_bf16(BT, D, 0, 1);  // A K-major (or TMEM), B MN-major
      const uint32_t idesc_dq = ptx::umma_idesc_bf16(BT, D, 1, 1);   // A = dZ read MN-major from the dZ^T tile
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      const uint32_t dzt_addr = ptx::smem_u32(smem_dzt);
      // S^T / dP^T of pair p+1 are issued AHEAD of the accumulating MMAs of pair p (both only need the element-wise
      // warps to be done with pair p), so the exponentials of pair p+1 run under the accumulating MMAs of pair p.
      int prev_q = -1, prev_k = -1, qn = 0, kn = 0;  // operand bookkeeping of the pair whose S^T was issued last
      int qb_cur = 0, kb_cur = 0;
      auto issue_scores = [&](const Part& pt) {
        if (pt.q != prev_q) {
          ptx::mbar_wait_trap(&q_full[qn & 1], (qn >> 1) & 1);
          prev_q = pt.q;
          ++qn;
        }
        if (pt.k != prev_k) {
          ptx::mbar_wait_trap(&kv_full[kn & 1], (kn >> 1) & 1);
          prev_k = pt.k;
          ++kn;
        }
        qb_cur = (qn - 1) & 1;
        kb_cur = (kn - 1) & 1;
        const uint32_t q_addr = ptx::smem_u32(smem_q + qb_cur * TILE), do_addr = ptx::smem_u32(smem_do + qb_cur * TILE);
        const uint32_t k_addr = ptx::smem_u32(smem_k + kb_cur * TILE), v_addr = ptx::smem_u32(smem_v + kb_cur * TILE);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          if (leader) ptx::umma_bf16(tmem_base + TM_ST, ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          if (leader) ptx::umma_bf16(tmem_base + TM_DPT, ptx::umma_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        if (leader) ptx::umma_commit(s_full);
      };
      int pi = 0;
#ifdef CM3P_ATTN_PROF
      long long pf_wait = 0, pf_sc = 0, pf_acc = 0, pf_t0 = clock64(), pf_a = pf_t0, pf_b;
#define PFI(acc) do { pf_b = clock64(); acc += pf_b - pf_a; pf_a = pf_b; } while (0)
#else
#define PFI(acc)
#endif
      issue_scores(w.get());
      for (; !w.done(); ++pi) {
        const Part pt = w.get();
        const int qb = qb_cur, kb = kb_cur;  // operand buffers of pair pi
        PFI(pf_acc);
        ptx::mbar_wait_trap(pz_full, pi & 1);  // P^T / dZ^T of this pair written, S^T / dP^T drained, previous epilogue read
        PFI(pf_wait);
        ptx::tc_fence_after();
        w.advance();
        const bool more = !w.done();
        const Part np = more ? w.get() : pt;
        if (more) issue_scores(np);
        PFI(pf_sc);
        const uint32_t q_addr = ptx::smem_u32(smem_q + qb * TILE), do_addr = ptx::smem_u32(smem_do + qb * TILE);
        const uint32_t k_addr = ptx::smem_u32(smem_k + kb * TILE);
        if (pt.need_kv()) {
          const uint32_t do_lo = ptx::umma_desc_lo(do_addr, 8192);
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)  // dV (+)= P^T dO: 16 queries per step = 8 TMEM columns of P^T
            if (leader) ptx::umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_PT + k * 8, do_lo + ((k * 2048) >> 4), HI, idesc_acc,
                              (k != 0 || !pt.kv_first()) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)  // dK (+)= dZ^T Q
            if (leader) ptx::umma_bf16(tmem_base + TM_DK,
                           ptx::umma_smem_desc_sw128(dzt_addr + (k >> 2) * (BT * 128) + (k & 3) * 32, 16, 1024),
                           ptx::umma_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_acc,
                           (k != 0 || !pt.kv_first()) ? 1u : 0u);
        }
        if (pt.need_q()) {
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)  // dQ (+)= dZ K: 16 keys per step = 16 rows of the dZ^T tile
            if (leader) ptx::umma_bf16(tmem_base + TM_DQ, ptx::umma_smem_desc_sw128(dzt_addr + k * 2048, BT * 128, 1024),
                           ptx::umma_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_dq,
                           (k != 0 || !pt.q_first()) ? 1u : 0u);
        }
        if (leader) {
          ptx::umma_commit(acc_done);
          // operand buffers go back to the producer after their last pair
          if (!more || np.q != pt.q) ptx::umma_commit(&q_empty[qb]);
          if (!more || np.k != pt.k) ptx::umma_commit(&kv_empty[kb]);
        }
        __syncwarp();
      }
#ifdef CM3P_ATTN_PROF
      PFI(pf_acc);
      if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0)
        printf("win bwd issuer: pairs=%d total=%lld wait_pz=%lld scores(incl operand wait)=%lld acc_issue=%lld\n", pi,
               clock64() - pf_t0, pf_wait, pf_sc, pf_acc);
#endif
    }
  } else if (warp < EW_WARPS) {
    // ------------------------------------------------------------------ element-wise warps + epilogues
    // 8 warps: TMEM lane quadrant (warp & 3) x 64-query column half (warp >> 2), two 32-column chunks each; chunks
    // outside the band are skipped, chunks entirely inside it skip the mask.  (16 one-chunk warps were tried: at 96
    // registers per thread the loop state spilled, and with 1 CTA x 640 threads per SM the local-memory lines do not
    // survive in the small L1 next to 227 KB of shared memory - every reload was an L2 round trip.)
    const int quad = warp & 3, half = warp >> 2;
    const int t = quad * 32 + lane;  // TMEM lane: key row of the key tile, query row of the dQ accumulator
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale_log2;
    const bool rope = p.rope_table != nullptr && p.positions != nullptr;
    // Epilogue of pair p (dK / dV of a finished key tile, dQ of a finished query tile) in "units" of 2 x 16 accumulator
    // columns for the warp's 32 rows:
    //   key tile finished:   half-0 warps write dK (unit s = columns {16s.., 32+16s..}, s = 0, 1: a rotation pair stays
    //                        in one thread), half-1 warps dV (unit s = columns [32s, 32s + 32))
    //   query tile finished: both halves write dQ, unit {16h.., 32+16h..} with h = half
    // The TMEM reads happen before this pair's pz_full arrive (the accumulating MMAs of the pair may then touch the
    // accumulators), the arithmetic and the global stores after it, under the score MMAs of the next pair.
    // Global traffic is issued by rows, not by lanes: with one table row / one gradient row per lane every load and
    // store instruction touched 32 different 128-byte lines, and the L1 tag stage set the pace of the whole pair
    // (measured: ~4000 of 11000 clk per pair).  Here 8 lanes fetch one 128-byte half row of rotation factors (4 rows
    // per cp.async instruction) into the warp's shared-memory slot one pair ahead; after the rotation each lane parks
    // its 64 output bytes in the same (now dead) row and the warp writes them out as 4 lanes per row, 8 rows per
    // store instruction.
    uint8_t* cs_warp = smem_cs + warp * 8192;  // [2 slots][32 rows][128 B], 16-byte unit u of row r at u ^ (r & 7)
    uint32_t cs_mask = 0;          // slots holding prefetched factors (warp-uniform)
    Part pend;                     // the pair whose epilogue is pending
    bool have_pend = false;
    auto kv_unit = [&](const Part& pt) { return pt.need_kv() && pt.kv_last(); };
    auto q_unit = [&](const Part& pt) { return pt.need_q() && pt.q_last(); };
    // sequence position of row 0 of this warp's 32 rows in the unit it writes after pair pt (warp-uniform)
    auto unit_row0 = [&](const Part& pt, bool want_q) { return (want_q ? pt.q * BT : pt.k * BT - KOFF) + quad * 32; };
    struct Unit {
      uint32_t acc;   // TMEM column base of the accumulator
      int x1, x2;     // first column of the two 16-column runs
      bool rotate;    // the runs are the two halves of rotation pairs (dK, dQ); dV is stored as is
      int row0;       // sequence position of the warp's first row
      int col0;       // column of the head's first element in the dqkv row (q | k | v block)
      int slot;       // shared-memory slot (rotation factors in, staged rows out)
      int hsel;       // which 16 frequencies the unit rotates
    };
    auto make_unit = [&](const Part& pt, bool want_q, int s) -> Unit {
      Unit u;
      u.row0 = unit_row0(pt, want_q);
      u.slot = s;
      if (want_q) {          // dQ
        u.hsel = half; u.acc = TM_DQ; u.x1 = half * 16; u.x2 = 32 + half * 16; u.rotate = rope; u.col0 = head * D;
      } else if (half == 0) {  // dK
        u.hsel = s; u.acc = TM_DK; u.x1 = s * 16; u.x2 = 32 + s * 16; u.rotate = rope; u.col0 = p.hidden + head * D;
      } else {                 // dV
        u.hsel = s; u.acc = TM_DV; u.x1 = s * 32; u.x2 = s * 32 + 16; u.rotate = false; u.col0 = 2 * p.hidden + head * D;
      }
      return u;
    };
    // position id of this lane's row in the unit(s) this warp rotates after pair pt (-1: none); issued at the top of
    // the pair so that the load completes under the score wait
    auto rotation_pos = [&](const Part& pt) -> int {
      if (!rope) return -1;
      const bool want_q = !kv_unit(pt);  // (a pair that finishes both: dQ's factors are fetched in the cold path)
      if (want_q ? !q_unit(pt) : (half != 0)) return -1;
      const int r_mine = unit_row0(pt, want_q) + lane;
      return (r_mine >= 0 && r_mine < len) ? __ldg(p.positions + seq_start + r_mine) : -1;
    };
    // rotation factors of those rows -> shared memory (asynchronously, L1 bypassed)
    auto prefetch_rotation = [&](const Part& pt, int pos) {
      __syncwarp();  // the slots are also the staging buffers of the previous units' stores
      cs_mask = 0;
      if (!__any_sync(0xffffffffu, pos >= 0)) return;
      const bool want_q = !kv_unit(pt);
      const int slots = want_q ? 1 : 2;
      for (int s = 0; s < slots; ++s) {
        const int hs = want_q ? half : s;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = 4 * i + (lane >> 3), u = lane & 7;
          const int pr = __shfl_sync(0xffffffffu, pos, r);
          if (pr >= 0)
            ptx::cp_async_16_cg(cs_warp + s * 4096 + r * 128 + ((u ^ (r & 7)) << 4),
                                reinterpret_cast<const uint8_t*>(p.rope_table + static_cast<int64_t>(pr) * 32) + hs * 128 + u * 16);
        }
        cs_mask |= 1u << s;
      }
    };
    auto unit_load = [&](const Unit& u, uint32_t (&r1)[16], uint32_t (&r2)[16]) {  // warp-collective
      ptx::tmem_ld_32x32b_x16(tmem_base + u.acc + lane_off + u.x1, r1);
      ptx::tmem_ld_32x32b_x16(tmem_base + u.acc + lane_off + u.x2, r2);
    };
    // accumulator registers -> (inverse rotation) -> 16 packed bf16 pairs: o[0..7] = first run, o[8..15] = second run
    auto unit_rotate = [&](const Unit& u, const uint32_t (&r1)[16], const uint32_t (&r2)[16], uint32_t (&o)[16]) {
      const uint8_t* my_row = cs_warp + u.slot * 4096 + lane * 128;
      const int rr = u.row0 + lane;
      if (u.rotate && rr >= 0 && rr < len) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dx1 = dy1 c + dy2 s, dx2 = dy2 c - dy1 s  (inverse of the forward rotation)
          const float4 f = *reinterpret_cast<const float4*>(my_row + ((k ^ (lane & 7)) << 4));
          const float a0 = __uint_as_float(r1[2 * k]), b0 = __uint_as_float(r2[2 * k]);
          const float a1 = __uint_as_float(r1[2 * k + 1]), b1 = __uint_as_float(r2[2 * k + 1]);
          o[k] = ptx::pack_bf16x2(a0 * f.x + b0 * f.y, a1 * f.z + b1 * f.w);
          o[8 + k] = ptx::pack_bf16x2(b0 * f.x - a0 * f.y, b1 * f.z - a1 * f.w);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {  // dV; rows outside the sequence are never written out
          o[k] = ptx::pack_bf16x2(__uint_as_float(r1[2 * k]), __uint_as_float(r1[2 * k + 1]));
          o[8 + k] = ptx::pack_bf16x2(__uint_as_float(r2[2 * k]), __uint_as_float(r2[2 * k + 1]));
        }
      }
    };
    // packed rows -> the warp's shared-memory slot (the factors are dead by now) -> global, 8 rows per instruction
    auto unit_write = [&](const Unit& u, const uint32_t (&o)[16]) {  // warp-collective
      uint8_t* slot = cs_warp + u.slot * 4096;
      uint8_t* my_row = slot + lane * 128;
      __syncwarp();  // every lane has read its factors
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(my_row + ((j ^ (lane & 7)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {  // 4 lanes (two 32-byte sectors) per row
        const int r = 8 * i + (lane >> 2), j = lane & 3;
        const int pos = u.row0 + r;
        if (pos >= 0 && pos < len) {
          const uint4 v = *reinterpret_cast<const uint4*>(slot + r * 128 + ((j ^ (r & 7)) << 4));
          __nv_bfloat16* dst = p.dqkv + (static_cast<int64_t>(seq_start) + pos) * 3 * p.hidden + u.col0 +
                               (j < 2 ? u.x1 + 8 * j : u.x2 + 8 * (j - 2));
          *reinterpret_cast<uint4*>(dst) = v;
        }
      }
    };
    // factors fetched by prefetch_rotation have landed and are visible to the whole warp
    auto rotation_ready = [&]() {
      if (cs_mask != 0) {
        ptx::cp_async_wait_all();
        __syncwarp();
      }
    };
    // All units of pair pt, start to finish, one after the other (cold path, one copy of the code: the last pair of
    // the walk, and a pair that finishes both its tiles - dQ's factors were not prefetched then, the slots held dK's)
    auto units_cold = [&](const Part& pt) {
      const bool kvu = kv_unit(pt), qu = q_unit(pt);
      rotation_ready();
#pragma unroll 1
      for (int i = 0; i < 3; ++i) {
        const bool want_q = i == 2;
        if (want_q ? !qu : !kvu) continue;
        if (want_q && kvu && rope) {
          __syncwarp();
          const int r_mine = unit_row0(pt, true) + lane;
          const int pos = (r_mine >= 0 && r_mine < len) ? __ldg(p.positions + seq_start + r_mine) : -1;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = 4 * j + (lane >> 3), u = lane & 7;
            const int pr = __shfl_sync(0xffffffffu, pos, r);
            if (pr >= 0)
              ptx::cp_async_16_cg(cs_warp + r * 128 + ((u ^ (r & 7)) << 4),
                               reinterpret_cast<const uint8_t*>(p.rope_table + static_cast<int64_t>(pr) * 32) + half * 128 + u * 16);
          }
          ptx::cp_async_wait_all();
          __syncwarp();
        }
        uint32_t r1[16], r2[16], o[16];
        const Unit u = make_unit(pt, want_q, want_q ? 0 : i);
        unit_load(u, r1, r2);
        ptx::tmem_ld_wait();
        unit_rotate(u, r1, r2, o);
        unit_write(u, o);
        __syncwarp();
      }
    };

    int prev_q = -1, qn = 0, pi = 0;
#ifdef CM3P_ATTN_PROF
    long long pe_q = 0, pe_s = 0, pe_c = 0, pe_acc = 0, pe_st = 0, pe_post = 0, pe_t0 = clock64(), pe_a = pe_t0, pe_b;
    long long pe_c1 = 0, pe_ul = 0, pe_ar = 0, pe_rr = 0, pe_us = 0, pe_pf = 0;
#define PFE(acc) do { pe_b = clock64(); acc += pe_b - pe_a; pe_a = pe_b; } while (0)
#else
#define PFE(acc)
#endif
    for (; !w.done(); w.advance(), ++pi) {
      const Part pt = w.get();
      PFE(pe_post);
      if (pt.q != prev_q) {
        ptx::mbar_wait_trap(&q_full[qn & 1], (qn >> 1) & 1);  // lse / delta of the query tile are staged
        prev_q = pt.q;
        ++qn;
      }
      const float* vec = smem_vec + ((qn - 1) & 1) * 2 * BT;
      const int kk = pt.k * BT - KOFF + t;  // position of this lane's key in the sequence
      const bool key_valid = kk >= 0 && kk < len;
      const int rot_pos = rotation_pos(pt);
      PFE(pe_q);
      ptx::mbar_wait_trap(s_full, pi & 1);
      PFE(pe_s);
      ptx::tc_fence_after();
      uint8_t* dz_tile = smem_dzt + half * (BT * 128);
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        const int c0 = half * 64 + ch * 32;  // first query column of the chunk inside the tile
        const int q0 = pt.q * BT + c0;       // its position in the sequence
        // allowed query columns [a, b) of this key row inside the chunk: |q - k| <= window, q < len
        int a = max(0, kk - p.window - q0);
        int b = min(min(32, len - q0), kk + p.window + 1 - q0);
        if (!key_valid) b = a;
        const bool skip = __all_sync(0xffffffffu, a >= b);               // chunk outside the band for the whole warp
        const bool inside = __all_sync(0xffffffffu, a <= 0 && b >= 32);  // chunk entirely inside it: no mask
        uint32_t pp[16], pz[16];
        if (skip) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pp[i] = 0u;
        } else if (inside) {
          chunk_math<false>(tmem_base + lane_off + c0, vec + c0, c, p.scale, a, b, pp, pz);
        } else {
          chunk_math<true>(tmem_base + lane_off + c0, vec + c0, c, p.scale, a, b, pp, pz);
        }
        if (ch == 0) {
          PFE(pe_c);
          if (pi > 0) {
            // P^T / dZ^T still belong to the accumulating MMAs of the previous pair until they have retired; its
            // accumulators are final then
            ptx::mbar_wait_trap(acc_done, (pi - 1) & 1);
            ptx::tc_fence_after();
          }
          PFE(pe_acc);
        }
        ptx::tmem_st_32x32b_x16(tmem_base + TM_PT + lane_off + half * 32 + ch * 16, pp);
        if (skip) store_zero_units(dz_tile, t, ch * 4);
        else store_row_units(dz_tile, t, ch * 4, pz);
      }
      PFE(pe_c1);
      // the previous pair's finished accumulators: into registers before this pair's MMAs may touch them
      uint32_t o[2][16];
      int n_units = 0;
      bool unit_q = false;
      if (have_pend) {
        const bool kvu = kv_unit(pend), qu = q_unit(pend);
        if (kvu && qu) {  // a pair that finished both its tiles (the last pair of a sequence)
          units_cold(pend);
        } else if (kvu || qu) {
          unit_q = qu;
          n_units = qu ? 1 : 2;
          rotation_ready();
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            if (s2 < n_units) {
              uint32_t r1[16], r2[16];
              const Unit u = make_unit(pend, unit_q, s2);
              unit_load(u, r1, r2);
              ptx::tmem_ld_wait();
              unit_rotate(u, r1, r2, o[s2]);
            }
          }
        }
      }
      PFE(pe_ul);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(pz_full);
      PFE(pe_ar);
      if (n_units > 0) {
        unit_write(make_unit(pend, unit_q, 0), o[0]);
        if (n_units == 2) unit_write(make_unit(pend, unit_q, 1), o[1]);
        PFE(pe_us);
      }
      have_pend = kv_unit(pt) || q_unit(pt);
      pend = pt;
      prefetch_rotation(pt, rot_pos);
      PFE(pe_pf);
    }
#ifdef CM3P_ATTN_PROF
    PFE(pe_post);
    if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0)
      printf("win bwd warp %d: pairs=%d total=%lld wait_q+setup=%lld wait_s=%lld chunk0=%lld wait_acc=%lld "
             "st0+chunk1+st1=%lld unit_load=%lld fences+arrive=%lld rotation_ready=%lld unit_store=%lld prefetch=%lld "
             "walk=%lld\n", warp, pi, clock64() - pe_t0, pe_q, pe_s, pe_c, pe_acc, pe_c1, pe_ul, pe_ar, pe_rr, pe_us, pe_pf,
             pe_post);
#endif
    // the last pair
    ptx::mbar_wait_trap(acc_done, (pi - 1) & 1);
    ptx::tc_fence_after();
    if (have_pend) units_cold(pend);
    ptx::tc_fence_before();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == EW_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// delta[head, row] = <dO[row, head, :], O[row, head, :]>: 8 lanes per (row, head), one 16-byte load of each operand
__global__ void __launch_bounds__(256)
attn_bwd_win_delta_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                          float* __restrict__ delta, int64_t T, int heads) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;  // 16-byte unit of the [T, heads*64] matrix
  const int64_t total = T * heads * 8;
  float s = 0.f;
  if (idx < total) {
    float a[8], b[8];
    unpack8f(__ldg(reinterpret_cast<const uint4*>(out) + idx), a);
    unpack8f(__ldg(reinterpret_cast<const uint4*>(dout) + idx), b);
#pragma unroll
    for (int k = 0; k < 8; ++k) s += a[k] * b[k];
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  if (idx < total && (threadIdx.x & 7) == 0) {
    const int64_t rh = idx >> 3;  // row * heads + head
    delta[(rh % heads) * T + rh / heads] = s;
  }
}

}  // namespace win
}  // namespace

int attn_varlen_bwd_window(const AttnBwdArgs& a, cudaStream_t stream) {
  using namespace win;
  const uint64_t H = static_cast<uint64_t>(a.heads) * 64;
  const uint64_t T = static_cast<uint64_t>(a.total_tokens);
  CM3P_REQUIRE(a.window >= 0 && a.window <= KOFF, kBadShape, "attn_bwd(window walk): window %d must be in [0, %d]",
               a.window, KOFF);
  CUtensorMap qkv128, do128;
  int rc;
  if ((rc = encode_tmap_2d_bf16(&qkv128, a.qkv, 3 * H, T, 3 * H * 2, 64, BT)) != kOk) return rc;
  if ((rc = encode_tmap_2d_bf16(&do128, a.dout, H, T, H * 2, 64, BT)) != kOk) return rc;
  CM3P_ENSURE_DYN_SMEM(attn_bwd_win_kernel, SMEM_BYTES);
  WinParams p;
  p.cu_seqlens = a.cu_seqlens;
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
  attn_bwd_win_delta_kernel<<<static_cast<unsigned>((vec_units + 255) / 256), 256, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(a.out), reinterpret_cast<const __nv_bfloat16*>(a.dout), a.delta,
      a.total_tokens, a.heads);
  CM3P_CUDA_TRY(cudaGetLastError());
  // query tiles per CTA: enough CTAs for ~8 waves, whole (sequence, head) walks when there is plenty of work
  const int forced = get_option(kOptBwdOuterPerCta);
  const int64_t units = (a.total_tokens / BT + a.batch / 2 + 1) * a.heads;
  const int64_t target_ctas = static_cast<int64_t>(num_sms()) * 8;
  int tpc = static_cast<int>((units + target_ctas - 1) / target_ctas);
  tpc = tpc < 2 ? 2 : (tpc > MAX_TILES_PER_CTA ? MAX_TILES_PER_CTA : tpc);
  if (forced > 0) tpc = forced > MAX_TILES_PER_CTA ? MAX_TILES_PER_CTA : forced;
  p.tiles_per_cta = tpc;
  const int max_q_tiles = (a.max_seqlen + BT - 1) / BT;
  p.ctas_per_seq = (max_q_tiles + tpc - 1) / tpc;
  CM3P_REQUIRE(static_cast<int64_t>(p.ctas_per_seq) * a.batch <= 0x7fffffffLL && a.heads <= 65535, kBadShape,
               "attn_bwd: grid too large (batch=%d max_seqlen=%d heads=%d)", a.batch, a.max_seqlen, a.heads);
  dim3 grid(static_cast<unsigned>(p.ctas_per_seq) * a.batch, a.heads, 1);
  attn_bwd_win_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(qkv128, do128, p);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
