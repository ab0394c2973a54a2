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

// One (query tile, key tile) pair of the walk and what happens to the accumulators around it.
struct Part {
  int q, k;
  bool need_q, q_first, q_last;     // dQ_q: computed here / first pair of its accumulation / complete after this pair
  bool need_kv, kv_first, kv_last;  // dK_k, dV_k likewise
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
    p.need_q = !halo;
    p.q_first = phase == 0;
    p.q_last = phase == 1 || step + 1 >= n_k;
    if (phase == 0) {
      p.need_kv = true;
      p.kv_first = step == 0;  // key tile 0 has no pair above it
      p.kv_last = true;
    } else {
      p.need_kv = halo || step + 1 < i1 || i1 == n_q;  // does this CTA own key tile step + 1 ?
      p.kv_first = true;
      p.kv_last = step + 1 >= n_q;  // the key tile below the last query tile has no second pair
    }
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
        ptx::mbar_wait(&q_empty[b], ((qn >> 1) & 1) ^ 1);
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
          vec[c] = ok ? lse_h[row0 + c] : 0.f;
          vec[BT + c] = ok ? delta_h[row0 + c] * p.scale : 0.f;
        }
        ptx::mbar_arrive(&q_full[b]);
        prev_q = pt.q;
        ++qn;
      }
      if (pt.k != prev_k) {
        const int b = kn & 1;
        ptx::mbar_wait(&kv_empty[b], ((kn >> 1) & 1) ^ 1);
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
    if (lane == 0) {
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
          ptx::mbar_wait(&q_full[qn & 1], (qn >> 1) & 1);
          prev_q = pt.q;
          ++qn;
        }
        if (pt.k != prev_k) {
          ptx::mbar_wait(&kv_full[kn & 1], (kn >> 1) & 1);
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
          ptx::umma_bf16(tmem_base + TM_ST, ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TM_DPT, ptx::umma_smem_desc_sw128(v_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(do_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
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
        ptx::mbar_wait(pz_full, pi & 1);  // P^T / dZ^T of this pair written, S^T / dP^T drained, previous epilogue read
        PFI(pf_wait);
        ptx::tc_fence_after();
        w.advance();
        const bool more = !w.done();
        const Part np = more ? w.get() : pt;
        if (more) issue_scores(np);
        PFI(pf_sc);
        const uint32_t q_addr = ptx::smem_u32(smem_q + qb * TILE), do_addr = ptx::smem_u32(smem_do + qb * TILE);
        const uint32_t k_addr = ptx::smem_u32(smem_k + kb * TILE);
        if (pt.need_kv) {
          const uint32_t do_lo = ptx::umma_desc_lo(do_addr, 8192);
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)  // dV (+)= P^T dO: 16 queries per step = 8 TMEM columns of P^T
            ptx::umma_bf16_ts(tmem_base + TM_DV, tmem_base + TM_PT + k * 8, do_lo + ((k * 2048) >> 4), HI, idesc_acc,
                              (k != 0 || !pt.kv_first) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)  // dK (+)= dZ^T Q
            ptx::umma_bf16(tmem_base + TM_DK,
                           ptx::umma_smem_desc_sw128(dzt_addr + (k >> 2) * (BT * 128) + (k & 3) * 32, 16, 1024),
                           ptx::umma_smem_desc_sw128(q_addr + k * 2048, 8192, 1024), idesc_acc,
                           (k != 0 || !pt.kv_first) ? 1u : 0u);
        }
        if (pt.need_q) {
#pragma unroll
          for (int k = 0; k < BT / 16; ++k)  // dQ (+)= dZ K: 16 keys per step = 16 rows of the dZ^T tile
            ptx::umma_bf16(tmem_base + TM_DQ, ptx::umma_smem_desc_sw128(dzt_addr + k * 2048, BT * 128, 1024),
                           ptx::umma_smem_desc_sw128(k_addr + k * 2048, 8192, 1024), idesc_dq,
                           (k != 0 || !pt.q_first) ? 1u : 0u);
        }
        ptx::umma_commit(acc_done);
        // operand buffers go back to the producer after their last pair
        if (!more || np.q != pt.q) ptx::umma_commit(&q_empty[qb]);
        if (!more || np.k != pt.k) ptx::umma_commit(&kv_empty[kb]);
      }
#ifdef CM3P_ATTN_PROF
      PFI(pf_acc);
      if (blockIdx.x == 0 && blockIdx.y == 0)
        printf("win bwd issuer: pairs=%d total=%lld wait_pz=%lld scores(incl operand wait)=%lld acc_issue=%lld\n", pi,
               clock64() - pf_t0, pf_wait, pf_sc, pf_acc);
#endif
    }
  } else {
    // ------------------------------------------------------------------ element-wise warps + epilogues
    const int quad = warp & 3, half = warp >> 2;
    const int t = quad * 32 + lane;  // TMEM lane: key row of the key tile, query row of the dQ accumulator
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const float c = p.scale_log2;
    const bool rope = p.rope_table != nullptr && p.positions != nullptr;
    // Epilogue of pair p (dK / dV of a finished key tile, dQ of a finished query tile) runs one pair late, in the
    // shadow of the accumulating MMAs: its rotation factors travel global -> shared memory with cp.async right after
    // the element-wise work of pair p (two dependent global loads per row otherwise sit in the serial chain of every
    // pair), and its TMEM reads sit between the two chunks of pair p+1.
    uint8_t* cs_row = smem_cs + (half * BT + t) * 256;  // this thread's row of rotation factors
    bool cs_loaded = false;
    Part pend;                     // the pair whose epilogue is pending
    bool have_pend = false;
    int pend_kk = 0;
    // row whose gradient this warp rotates in the epilogue of pair pt (-1: none): half 0 warps write dK of a finished
    // key tile, half 1 warps dQ of a finished query tile (and the un-rotated dV)
    auto rotation_row = [&](const Part& pt, int kk) -> int64_t {
      if (!rope) return -1;
      if (half == 0) {
        if (pt.need_kv && pt.kv_last && kk >= 0 && kk < len) return static_cast<int64_t>(seq_start) + kk;
      } else {
        if (pt.need_q && pt.q_last && pt.q * BT + t < len) return static_cast<int64_t>(seq_start) + pt.q * BT + t;
      }
      return -1;
    };
    auto prefetch_rotation = [&](int pos) {  // pos = positions[rotation_row], fetched a pair earlier; -1: nothing to do
      cs_loaded = false;
      if (pos >= 0) {
        const uint8_t* tab = reinterpret_cast<const uint8_t*>(p.rope_table + static_cast<int64_t>(pos) * 32);
#pragma unroll
        for (int k = 0; k < 16; ++k) ptx::cp_async_16(cs_row + ((k ^ (t & 15)) << 4), tab + k * 16);
        cs_loaded = true;
      }
    };
#ifdef CM3P_ATTN_PROF
    long long pq_ld = 0, pq_rot = 0, pq_stg = 0, pq_tl = 0, pq_math = 0, pq_a, pq_b;
#define PFQ0() pq_a = clock64()
#define PFQ(acc) do { pq_b = clock64(); acc += pq_b - pq_a; pq_a = pq_b; } while (0)
#else
#define PFQ0()
#define PFQ(acc)
#endif
    auto rotate_store = [&](uint32_t taddr, __nv_bfloat16* dst, bool valid, bool use_cs) {
      uint32_t r1[32], r2[32];
      PFQ0();
      ptx::tmem_ld_32x32b_x32(taddr, r1);
      ptx::tmem_ld_32x32b_x32(taddr + 32, r2);
      ptx::tmem_ld_wait();
      PFQ(pq_ld);
      if (!valid) return;
      uint32_t o1[16], o2[16];
      if (use_cs) {
        if (cs_loaded) ptx::cp_async_wait_all();  // this lane's own copies: no cross-thread visibility needed
#pragma unroll
        for (int k = 0; k < 16; ++k) {  // dx1 = dy1 c + dy2 s, dx2 = dy2 c - dy1 s  (inverse of the forward rotation)
          const float4 f = *reinterpret_cast<const float4*>(cs_row + ((k ^ (t & 15)) << 4));
          const float a0 = __uint_as_float(r1[2 * k]), b0 = __uint_as_float(r2[2 * k]);
          const float a1 = __uint_as_float(r1[2 * k + 1]), b1 = __uint_as_float(r2[2 * k + 1]);
          o1[k] = ptx::pack_bf16x2(a0 * f.x + b0 * f.y, a1 * f.z + b1 * f.w);
          o2[k] = ptx::pack_bf16x2(b0 * f.x - a0 * f.y, b1 * f.z - a1 * f.w);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          o1[k] = ptx::pack_bf16x2(__uint_as_float(r1[2 * k]), __uint_as_float(r1[2 * k + 1]));
          o2[k] = ptx::pack_bf16x2(__uint_as_float(r2[2 * k]), __uint_as_float(r2[2 * k + 1]));
        }
      }
      PFQ(pq_rot);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<uint4*>(dst + i * 8) = make_uint4(o1[4 * i], o1[4 * i + 1], o1[4 * i + 2], o1[4 * i + 3]);
        *reinterpret_cast<uint4*>(dst + 32 + i * 8) = make_uint4(o2[4 * i], o2[4 * i + 1], o2[4 * i + 2], o2[4 * i + 3]);
      }
      PFQ(pq_stg);
    };
    auto epilogue = [&](const Part& pt, int kk) {  // accumulators of pair pt are final (acc_done waited by the caller)
      if (pt.need_kv && pt.kv_last) {
        const bool key_valid = kk >= 0 && kk < len;
        const int64_t row = static_cast<int64_t>(seq_start) + kk;
        __nv_bfloat16* base = p.dqkv + row * 3 * p.hidden + head * D;
        if (half == 0) rotate_store(tmem_base + TM_DK + lane_off, base + p.hidden, key_valid, rope);
        else rotate_store(tmem_base + TM_DV + lane_off, base + 2 * p.hidden, key_valid, false);
      }
      if (pt.need_q && pt.q_last && half == 1) {
        const int qq = pt.q * BT + t;
        const bool q_valid = qq < len;
        const int64_t row = static_cast<int64_t>(seq_start) + qq;
        rotate_store(tmem_base + TM_DQ + lane_off, p.dqkv + row * 3 * p.hidden + head * D, q_valid, rope);
      }
    };

    int prev_q = -1, qn = 0, pi = 0;
#ifdef CM3P_ATTN_PROF
    long long pe_s = 0, pe_c = 0, pe_acc = 0, pe_epi = 0, pe_st = 0, pe_t0 = clock64(), pe_a = pe_t0, pe_b;
#define PFE(acc) do { pe_b = clock64(); acc += pe_b - pe_a; pe_a = pe_b; } while (0)
#else
#define PFE(acc)
#endif
    for (; !w.done(); w.advance(), ++pi) {
      const Part pt = w.get();
      if (pt.q != prev_q) {
        ptx::mbar_wait(&q_full[qn & 1], (qn >> 1) & 1);  // lse / delta of the query tile are staged
        prev_q = pt.q;
        ++qn;
      }
      const float* vec = smem_vec + ((qn - 1) & 1) * 2 * BT;
      const int kk = pt.k * BT - KOFF + t;  // position of this lane's key in the sequence
      const bool key_valid = kk >= 0 && kk < len;
      // position id of the row this warp rotates after this pair: the load completes under the score wait below
      const int64_t rot_row = rotation_row(pt, kk);
      const int rot_pos = rot_row >= 0 ? __ldg(p.positions + rot_row) : -1;
      PFE(pe_st);
      ptx::mbar_wait(s_full, pi & 1);
      PFE(pe_s);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int ch = 0; ch < 2; ++ch) {
        const int c0 = half * 64 + ch * 32;  // first query column of the chunk inside the tile
        const int q0 = pt.q * BT + c0;       // its position in the sequence
        // allowed query columns [a, b) of this key row inside the chunk: |q - k| <= window, q < len
        int a = max(0, kk - p.window - q0);
        int b = min(min(32, len - q0), kk + p.window + 1 - q0);
        if (!key_valid) b = a;
        const bool skip = __all_sync(0xffffffffu, a >= b);  // chunk outside the band for the whole warp
        uint8_t* dz_tile = smem_dzt + half * (BT * 128);
        uint32_t pp[16], pz[16];
        if (skip) {
#pragma unroll
          for (int i = 0; i < 16; ++i) pp[i] = 0u;
        } else {
          uint32_t rs[32], rp[32];
          PFQ0();
          ptx::tmem_ld_32x32b_x32(tmem_base + TM_ST + lane_off + c0, rs);
          ptx::tmem_ld_32x32b_x32(tmem_base + TM_DPT + lane_off + c0, rp);
          ptx::tmem_ld_wait();
          PFQ(pq_tl);
          const float4* lse4 = reinterpret_cast<const float4*>(vec + c0);
          const float4* del4 = reinterpret_cast<const float4*>(vec + BT + c0);
#pragma unroll
          for (int i4 = 0; i4 < 32; i4 += 4) {
            const float4 l4 = lse4[i4 >> 2];
            const float4 d4 = del4[i4 >> 2];
            const float ls[4] = {l4.x, l4.y, l4.z, l4.w};
            const float dl[4] = {d4.x, d4.y, d4.z, d4.w};
            float pv[4], zv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int j = i4 + e;
              const bool on = j >= a && j < b;
              const float ex = ptx::ex2_approx(__uint_as_float(rs[j]) * c - ls[e]);
              pv[e] = on ? ex : 0.f;
              zv[e] = on ? ex * (__uint_as_float(rp[j]) * p.scale - dl[e]) : 0.f;
            }
            pp[i4 >> 1] = ptx::pack_bf16x2(pv[0], pv[1]);
            pp[(i4 >> 1) + 1] = ptx::pack_bf16x2(pv[2], pv[3]);
            pz[i4 >> 1] = ptx::pack_bf16x2(zv[0], zv[1]);
            pz[(i4 >> 1) + 1] = ptx::pack_bf16x2(zv[2], zv[3]);
          }
          PFQ(pq_math);
        }
        PFE(pe_c);
        if (ch == 0 && pi > 0) {
          // P^T / dZ^T still belong to the accumulating MMAs of the previous pair until they have retired; its
          // accumulators are final then: read them out before this pair's MMAs can touch them
          ptx::mbar_wait(acc_done, (pi - 1) & 1);
          PFE(pe_acc);
          ptx::tc_fence_after();
          if (have_pend) epilogue(pend, pend_kk);
          PFE(pe_epi);
        }
        ptx::tmem_st_32x32b_x16(tmem_base + TM_PT + lane_off + half * 32 + ch * 16, pp);
        if (skip) store_zero_units(dz_tile, t, ch * 4);
        else store_row_units(dz_tile, t, ch * 4, pz);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(pz_full);
      have_pend = (pt.need_kv && pt.kv_last) || (pt.need_q && pt.q_last);
      pend = pt;
      pend_kk = kk;
      prefetch_rotation(rot_pos);
    }
    // the last pair
    ptx::mbar_wait(acc_done, (pi - 1) & 1);
    ptx::tc_fence_after();
    if (have_pend) epilogue(pend, pend_kk);
    ptx::tc_fence_before();
#ifdef CM3P_ATTN_PROF
    PFE(pe_st);
    if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0)
      printf("win bwd warp %d: pairs=%d total=%lld wait_s=%lld chunk_compute=%lld (tmem_ld %lld math %lld) wait_acc=%lld "
             "epilogue=%lld (tmem_ld %lld rotate %lld stg %lld) store+arrive+prefetch=%lld\n", warp, pi,
             clock64() - pe_t0, pe_s, pe_c, pq_tl, pq_math, pe_acc, pe_epi, pq_ld, pq_rot, pq_stg, pe_st);
#endif
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
