// Unpadded (varlen) bidirectional attention forward for sm_100a, head_dim 64, covering ModernBERT's
// global layers and its +-window sliding layers with one kernel.
//
// Reference semantics (third-party transformers ModernBertAttention as called from
// /root/reference/cm3p/modeling_cm3p.py:359-369,509-514,607-619; window rule
// transformers/masking_utils.py:121-131): o = softmax_j(q_i.k_j / 8 over allowed j) v, with
// allowed(i, j) = j is a real token of the same sequence and (global or |i - j| <= window).
// q/k arrive already rotated (RoPE is fused into the Wqkv GEMM epilogue, gemm_sm100.cu EPI_ROPE).
//
// One CTA = one (sequence, head, 128-query tile).  KV is streamed in 128-row tiles by TMA
// (2 stages).  Both GEMMs run on tcgen05: S = Q K^T (M128 N128 K64) into TMEM, the softmax
// warps read S with tcgen05.ld (thread t owns query row t), write P (bf16) into 128B-swizzled
// smem, and O_j = P V_j (M128 N64 K128, V as an MN-major operand) lands in a second TMEM region
// that the softmax threads fold into their register accumulator with the usual online rescale.
// 112 KB smem + 256 TMEM columns per CTA -> two CTAs per SM overlap each other's MMA and softmax.
//
// Warps: 0..3 softmax/epilogue (TMEM lane quadrant = warp), 4 = TMA producer (+TMEM alloc),
//        5 = MMA issuer.
#include <cuda_bf16.h>
#include <math.h>
#include <stdlib.h>

#include "attn.h"
#include "common.h"
#include "ptx.cuh"

namespace cm3p {
namespace {

constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int D = 64;
constexpr int Q_BYTES = BQ * D * 2;       // 16 KB
constexpr int KV_TILE_BYTES = BKV * D * 2;  // 16 KB (K or V)
constexpr int KV_STAGES = 2;
constexpr int P_BYTES = BQ * BKV * 2;  // 32 KB, two 64-wide K blocks of 16 KB
constexpr int SMEM_TILES = Q_BYTES + KV_STAGES * 2 * KV_TILE_BYTES + P_BYTES;  // 112 KB
constexpr int SMEM_BYTES = SMEM_TILES + 256;
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 256;  // S: [0,128)  O_tmp: [128,192)
constexpr int TMEM_S = 0;
constexpr int TMEM_O = 128;

struct Params {
  const int32_t* cu_seqlens;  // [B+1]
  __nv_bfloat16* out;         // [T, H]
  float* lse;                 // [heads, T] log2-domain logsumexp of scaled scores, or nullptr
  int64_t total_tokens;
  int heads;
  int hidden;  // H = heads * 64
  int window;  // < 0: global
  float scale_log2;  // (1/sqrt(64)) * log2(e)
  int blocks_per_cta;  // v2: consecutive 256-query blocks streamed by one CTA
  int ctas_per_seq;    // grid.x = batch * ctas_per_seq (the sequence index lives in grid.x: grid.z stops at 65535)
};

__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tma_qkv, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int seq = blockIdx.x / p.ctas_per_seq;
  const int head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int q0 = (blockIdx.x % p.ctas_per_seq) * BQ;
  if (q0 >= len) return;  // whole CTA exits together, before any barrier / TMEM use

  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + Q_BYTES;
  uint8_t* smem_v = smem_k + KV_STAGES * KV_TILE_BYTES;
  uint8_t* smem_p = smem_v + KV_STAGES * KV_TILE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_TILES);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_full = bars + 6;
  uint64_t* o_full = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();  // swizzle-128B needs 1 KB alignment

  // KV tile range
  int kv_lo = 0, kv_hi = len - 1;
  if (p.window >= 0) {
    kv_lo = max(0, q0 - p.window);
    kv_hi = min(len - 1, q0 + BQ - 1 + p.window);
  }
  const int tile_lo = kv_lo / BKV;
  const int n_tiles = kv_hi / BKV - tile_lo + 1;

  if (warp == 5 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < KV_STAGES; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) ptx::prefetch_tmap(&tma_qkv);
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
      ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_2d(smem_q, &tma_qkv, q_full, col_q, seq_start + q0);
      for (int j = 0; j < n_tiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait_trap(&kv_empty[s], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * KV_TILE_BYTES);
        const int row = seq_start + (tile_lo + j) * BKV;
        ptx::tma_load_2d(smem_k + s * KV_TILE_BYTES, &tma_qkv, &kv_full[s], col_k, row);
        ptx::tma_load_2d(smem_v + s * KV_TILE_BYTES, &tma_qkv, &kv_full[s], col_v, row);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ---------------------------------------------------------------- MMA issuer
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, D, 0, 1);  // V is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem_q);
      const uint32_t p_addr = ptx::smem_u32(smem_p);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        ptx::mbar_wait_trap(&kv_full[s], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem_k + s * KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TMEM_S, ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      };
      ptx::mbar_wait_trap(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        ptx::mbar_wait_trap(p_full, j & 1);  // P_j in smem, S_j and O_{j-1} consumed
        ptx::tc_fence_after();
        if (j + 1 < n_tiles) issue_s(j + 1);
        const int s = j & 1;
        const uint32_t v_addr = ptx::smem_u32(smem_v + s * KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          ptx::umma_bf16(tmem_base + TMEM_O,
                         ptx::umma_smem_desc_sw128(p_addr + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(v_addr + k * 2048, 8192, 1024), idesc_o, k != 0 ? 1u : 0u);
        ptx::umma_commit(o_full);
        ptx::umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax + epilogue
    const int t = threadIdx.x;  // query row inside the tile == TMEM lane
    const int qi = q0 + t;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    float m_prev = -INFINITY, l = 0.f;
    float o[D];
#pragma unroll
    for (int i = 0; i < D; ++i) o[i] = 0.f;

    for (int j = 0; j < n_tiles; ++j) {
      const int kv0 = (tile_lo + j) * BKV;
      ptx::mbar_wait_trap(s_full, j & 1);
      ptx::tc_fence_after();
      // allowed kv range for this row inside the tile: [a, b)
      int a = 0, b = min(BKV, len - kv0);
      if (p.window >= 0) {
        a = max(a, qi - p.window - kv0);
        b = min(b, qi + p.window + 1 - kv0);
      }
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int cidx = 0; cidx < BKV; cidx += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_S + lane_off + cidx, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int kj = cidx + i;
          const float sv = (kj >= a && kj < b) ? __uint_as_float(r[i]) : -INFINITY;
          mx = fmaxf(mx, sv);
        }
      }
      const float m_new = fmaxf(m_prev, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ptx::ex2_approx((m_prev - m_use) * c);  // 0 when m_prev = -inf
      // fold O_{j-1} (relative to m_prev) and rescale to m_new
      if (j > 0) {
        ptx::mbar_wait_trap(o_full, (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int h = 0; h < D; h += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_O + lane_off + h, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[h + i] = (o[h + i] + __uint_as_float(r[i])) * alpha;
        }
      }
      // pass 2: p = exp2(s*c - m*c), row sum, bf16 P into swizzled smem (A operand of P.V)
      float rs = 0.f;
      const float mc = m_use * c;
#pragma unroll 1
      for (int cidx = 0; cidx < BKV; cidx += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_S + lane_off + cidx, r);
        ptx::tmem_ld_wait();
        uint32_t packed[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const int kj = cidx + i;
          float p0 = (kj >= a && kj < b) ? ptx::ex2_approx(__uint_as_float(r[i]) * c - mc) : 0.f;
          float p1 = (kj + 1 >= a && kj + 1 < b) ? ptx::ex2_approx(__uint_as_float(r[i + 1]) * c - mc) : 0.f;
          packed[i >> 1] = ptx::pack_bf16x2(p0, p1);
          // sum what the tensor core will actually see (bf16-rounded probabilities)
          const float2 pr = ptx::unpack_bf16x2(packed[i >> 1]);
          rs += pr.x + pr.y;
        }
        uint8_t* prow = smem_p + (cidx >> 6) * (BQ * 128) + t * 128;
        const int u0 = (cidx & 32) ? 4 : 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int unit = (u0 + u) ^ (t & 7);
          *reinterpret_cast<uint4*>(prow + unit * 16) =
              make_uint4(packed[u * 4], packed[u * 4 + 1], packed[u * 4 + 2], packed[u * 4 + 3]);
        }
      }
      l = l * alpha + rs;
      m_prev = m_new;
      ptx::tc_fence_before();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(p_full);
    }
    // last P.V
    ptx::mbar_wait_trap(o_full, (n_tiles - 1) & 1);
    ptx::tc_fence_after();
#pragma unroll
    for (int h = 0; h < D; h += 32) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_O + lane_off + h, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[h + i] += __uint_as_float(r[i]);
    }
    if (qi < len) {
      const float inv = 1.f / l;
      const int64_t row = static_cast<int64_t>(seq_start) + qi;
      __nv_bfloat16* dst = p.out + row * p.hidden + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) {
        uint4 u;
        u.x = ptx::pack_bf16x2(o[i] * inv, o[i + 1] * inv);
        u.y = ptx::pack_bf16x2(o[i + 2] * inv, o[i + 3] * inv);
        u.z = ptx::pack_bf16x2(o[i + 4] * inv, o[i + 5] * inv);
        u.w = ptx::pack_bf16x2(o[i + 6] * inv, o[i + 7] * inv);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
      if (p.lse) p.lse[static_cast<int64_t>(head) * p.total_tokens + row] = m_prev * c + log2f(l);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}


// ================================================================================================
// Packed short sequences (metadata tower: ~17-25 real tokens per sequence, B*V up to 65 536 sequences).
// One 128-row tile per SEQUENCE wastes >80 % of every MMA and exp pass and pays a CTA's fixed cost per 21
// tokens.  Here consecutive sequences are packed into groups of <= 128 tokens (attn_pack_groups_kernel:
// greedy inside chunks of 64 sequences, one warp per chunk, group order irrelevant), and one CTA handles one
// (group, head): Q = K = V rows are the same 128-row slab of the token matrix, S = Q K^T is ONE 128x128 MMA
// whose mask is block diagonal (row i attends to the columns of its own sequence), O = P V one more.
// 80 KB smem, 256 TMEM columns: two CTAs per SM.
namespace packed {

constexpr int CHUNK_SEQS = 64;       // sequences walked by one warp of the group builder
constexpr int SB_INTS = CHUNK_SEQS + 4;

__global__ void __launch_bounds__(256)
attn_pack_groups_kernel(const int32_t* __restrict__ cu, int batch, int2* __restrict__ groups,
                        int32_t* __restrict__ n_groups, int max_groups) {
  __shared__ int32_t s_cu[8][CHUNK_SEQS + 1];
  __shared__ int2 s_grp[8][CHUNK_SEQS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int base = (blockIdx.x * 8 + warp) * CHUNK_SEQS;
  if (base >= batch) return;  // warp-uniform
  const int n = min(CHUNK_SEQS, batch - base);
  for (int i = lane; i <= n; i += 32) s_cu[warp][i] = cu[base + i];
  __syncwarp();
  int s = 0, k = 0;
  while (s < n) {
    const int c0 = s_cu[warp][s];
    // largest e in (s, n] with cu[e] - cu[s] <= 128 (monotone in e); at least s + 1
    const int e1 = lane + 1, e2 = lane + 33;
    const bool ok1 = e1 > s && e1 <= n && s_cu[warp][e1] - c0 <= BQ;
    const bool ok2 = e2 > s && e2 <= n && s_cu[warp][e2] - c0 <= BQ;
    const unsigned m1 = __ballot_sync(0xffffffffu, ok1), m2 = __ballot_sync(0xffffffffu, ok2);
    int e = s + 1;
    if (m2) e = 33 + (31 - __clz(m2));
    else if (m1) e = 1 + (31 - __clz(m1));
    if (s_cu[warp][e] - c0 > 0) {  // groups without tokens (empty sequences only) are dropped
      if (lane == 0) s_grp[warp][k] = make_int2(base + s, base + e);
      ++k;
    }
    s = e;
  }
  __syncwarp();
  int off = 0;
  if (lane == 0 && k > 0) off = atomicAdd(n_groups, k);
  off = __shfl_sync(0xffffffffu, off, 0);
  for (int i = lane; i < k; i += 32)
    if (off + i < max_groups) groups[off + i] = s_grp[warp][i];
}

constexpr int P_SMEM_TILES = 3 * Q_BYTES + P_BYTES;  // Q, K, V (16 KB each) + P (32 KB) = 80 KB
constexpr int P_SMEM_BYTES = P_SMEM_TILES + SB_INTS * 4 + 128;

struct PackedParams {
  const int32_t* cu_seqlens;
  const int2* groups;
  const int32_t* n_groups;
  __nv_bfloat16* out;
  float* lse;
  int64_t total_tokens;
  int hidden;
  float scale_log2;
};

// bounds [lo, hi) (rows relative to the group's first token) of the sequence that holds row t
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
attn_fwd_packed_kernel(const __grid_constant__ CUtensorMap tma_qkv, const PackedParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int g = blockIdx.x, head = blockIdx.y;
  if (g >= *p.n_groups) return;  // the grid is sized for the worst case
  const int2 grp = p.groups[g];
  const int tok0 = p.cu_seqlens[grp.x];
  const int rows = p.cu_seqlens[grp.y] - tok0;  // 1..128
  const int nseq = grp.y - grp.x;
  if (rows <= 0) return;

  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + Q_BYTES;
  uint8_t* smem_v = smem + 2 * Q_BYTES;
  uint8_t* smem_p = smem + 3 * Q_BYTES;
  int* sb = reinterpret_cast<int*>(smem + P_SMEM_TILES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P_SMEM_TILES + SB_INTS * 4);
  uint64_t* ld_full = bars;
  uint64_t* s_full = bars + 1;
  uint64_t* p_full = bars + 2;
  uint64_t* o_full = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();
  for (int i = threadIdx.x; i <= nseq; i += THREADS) sb[i] = p.cu_seqlens[grp.x + i] - tok0;
  if (warp == 5 && lane == 0) {
    ptx::mbar_init(ld_full, 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) ptx::prefetch_tmap(&tma_qkv);
    ptx::tmem_alloc(tmem_slot, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int ksteps = (rows + 15) >> 4;  // 16-key steps of P.V that can hold a non-zero probability

  if (warp == 4) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(ld_full, 3 * Q_BYTES);
      ptx::tma_load_2d(smem_q, &tma_qkv, ld_full, head * D, tok0);
      ptx::tma_load_2d(smem_k, &tma_qkv, ld_full, p.hidden + head * D, tok0);
      ptx::tma_load_2d(smem_v, &tma_qkv, ld_full, 2 * p.hidden + head * D, tok0);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, D, 0, 1);  // V is MN-major
      const uint32_t q_addr = ptx::smem_u32(smem_q), k_addr = ptx::smem_u32(smem_k);
      const uint32_t v_addr = ptx::smem_u32(smem_v), p_addr = ptx::smem_u32(smem_p);
      ptx::mbar_wait_trap(ld_full, 0);
      ptx::tc_fence_after();
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        ptx::umma_bf16(tmem_base + TMEM_S, ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                       ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
      ptx::umma_commit(s_full);
      ptx::mbar_wait_trap(p_full, 0);
      ptx::tc_fence_after();
      for (int k = 0; k < ksteps; ++k)
        ptx::umma_bf16(tmem_base + TMEM_O,
                       ptx::umma_smem_desc_sw128(p_addr + (k >> 2) * (BQ * 128) + (k & 3) * 32, 16, 1024),
                       ptx::umma_smem_desc_sw128(v_addr + k * 2048, 8192, 1024), idesc_o, k != 0 ? 1u : 0u);
      ptx::umma_commit(o_full);
    }
  } else {
    const int t = threadIdx.x;  // row inside the group == TMEM lane
    const bool valid = t < rows;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    const float c = p.scale_log2;
    int lo = 0, hi = 0;
    if (valid) row_bounds(sb, nseq, t, lo, hi);
    // columns any row of this warp may attend to: chunks outside are skipped (zeros, no TMEM read, no exp)
    const int wa = __reduce_min_sync(0xffffffffu, valid ? lo : BKV);
    const int wb = __reduce_max_sync(0xffffffffu, valid ? hi : 0);
    ptx::mbar_wait_trap(s_full, 0);
    ptx::tc_fence_after();
    float mx = -INFINITY;
#pragma unroll 1
    for (int c0 = 0; c0 < BKV; c0 += 32) {
      if (c0 + 32 <= wa || c0 >= wb) continue;  // warp-uniform
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_S + lane_off + c0, r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const int kj = c0 + i;
        mx = fmaxf(mx, (kj >= lo && kj < hi) ? __uint_as_float(r[i]) : -INFINITY);
      }
    }
    const float mc = (mx == -INFINITY) ? 0.f : mx * c;
    float rs = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < BKV; c0 += 32) {
      uint8_t* prow = smem_p + (c0 >> 6) * (BQ * 128) + t * 128;
      const int u0 = (c0 & 32) ? 4 : 0;
      if (c0 + 32 <= wa || c0 >= wb) {
#pragma unroll
        for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ (t & 7)) << 4)) = make_uint4(0, 0, 0, 0);
        continue;
      }
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_S + lane_off + c0, r);
      ptx::tmem_ld_wait();
      uint32_t packed[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const int kj = c0 + i;
        const float p0 = (kj >= lo && kj < hi) ? ptx::ex2_approx(__uint_as_float(r[i]) * c - mc) : 0.f;
        const float p1 = (kj + 1 >= lo && kj + 1 < hi) ? ptx::ex2_approx(__uint_as_float(r[i + 1]) * c - mc) : 0.f;
        packed[i >> 1] = ptx::pack_bf16x2(p0, p1);
        const float2 pr = ptx::unpack_bf16x2(packed[i >> 1]);  // sum what the tensor core multiplies
        rs += pr.x + pr.y;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        *reinterpret_cast<uint4*>(prow + (((u0 + u) ^ (t & 7)) << 4)) =
            make_uint4(packed[u * 4], packed[u * 4 + 1], packed[u * 4 + 2], packed[u * 4 + 3]);
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(p_full);
    ptx::mbar_wait_trap(o_full, 0);
    ptx::tc_fence_after();
    uint32_t o[64];
    {
      uint32_t r1[32], r2[32];
      ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_O + lane_off, r1);
      ptx::tmem_ld_32x32b_x32(tmem_base + TMEM_O + lane_off + 32, r2);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) { o[i] = r1[i]; o[32 + i] = r2[i]; }
    }
    if (valid) {
      const float inv = 1.f / rs;
      const int64_t row = static_cast<int64_t>(tok0) + t;
      __nv_bfloat16* dst = p.out + row * p.hidden + head * D;
#pragma unroll
      for (int i = 0; i < D; i += 8) {
        uint4 u;
        u.x = ptx::pack_bf16x2(__uint_as_float(o[i]) * inv, __uint_as_float(o[i + 1]) * inv);
        u.y = ptx::pack_bf16x2(__uint_as_float(o[i + 2]) * inv, __uint_as_float(o[i + 3]) * inv);
        u.z = ptx::pack_bf16x2(__uint_as_float(o[i + 4]) * inv, __uint_as_float(o[i + 5]) * inv);
        u.w = ptx::pack_bf16x2(__uint_as_float(o[i + 6]) * inv, __uint_as_float(o[i + 7]) * inv);
        *reinterpret_cast<uint4*>(dst + i) = u;
      }
      if (p.lse) p.lse[static_cast<int64_t>(head) * p.total_tokens + row] = mc + log2f(rs);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace packed

// ================================================================================================
// v2: one CTA per SM streams `blocks_per_cta` consecutive 256-query blocks of one (sequence, head); a block
// is two 128-row Q tiles A and B.  20 warps:
//   warps 0-7 softmax group A, warps 8-15 softmax group B: TMEM lane quadrant (warp & 3) x 64-key column half,
//   so a query row is shared by two threads that exchange their row maximum through shared memory once per
//   tile (four light warps per scheduler hide exp / FMA latency that two heavy ones could not);
//   warp 16 TMA producer (+ TMEM alloc), warps 17 / 18 MMA issuers of tile A / B, warp 19 idle.
// K/V tiles (128 rows) stream through a 4-stage ring shared by both Q tiles; Q is double-buffered across
// blocks, so the next block's loads and first S GEMM run under the current block's epilogue.  Per Q tile the
// pipeline is decoupled in both directions: S_x(t+1) is issued as soon as group x has copied S_x(t) into
// registers (s_free), and PV_x(t) is issued per 64-key half of P as soon as the four warps of that half have
// written it (p_full / pv_done per half), so neither the group nor the tensor pipe waits for the other in
// steady state.  P never touches shared memory: the softmax threads write it to TMEM as packed bf16
// (tcgen05.st, thread = row = lane) and PV reads its A operand from TMEM; with P in shared memory a K tile
// moved 256 KB through the SM's 128 B/clk (S and PV operand reads, P writes, TMA): 2048 of ~3000 clk.
// O_A / O_B stay in TMEM for the whole KV loop (tcgen05.mma accumulate); the running maximum is only raised
// when it grows by more than 2^8 (then each warp rescales its 32 rows x 32 columns of O in TMEM with
// tcgen05.ld/st), so the common iteration is: one TMEM read of 64 scores, max (3-input), exchange, exp2
// (packed f32x2 arithmetic, a quarter of the exponentials as a polynomial on the FMA pipe), bf16 pack,
// tcgen05.st, mbarrier arrive.  Masking (-inf) is only applied on tiles that need it (tail, window band).
// TMEM columns: S_A [0,128)  S_B [128,256)  O_A [256,320)  O_B [320,384)  P_A [384,448)  P_B [448,512).
namespace v2 {

constexpr int KV_STAGES2 = 4;
constexpr int THREADS2 = 640;  // 5 warpgroups: softmax A (2), softmax B (2), {TMA, MMA issuer A, MMA issuer B, idle}
constexpr int SM_WARPS = 16;    // softmax warps: 8 per Q tile = 4 TMEM lane quadrants x 2 column halves
constexpr int REG_SM = 104, REG_AUX2 = 64;
constexpr int SMEM_TILES2 = 2 * 2 * Q_BYTES + KV_STAGES2 * 2 * KV_TILE_BYTES;  // 64 + 128 = 192 KB
constexpr int SMEM_XCHG = 3 * 2 * 2 * BQ * 4;  // row maxima (one plane per tile parity) and row sums exchanged between the two warps of a row
constexpr int SMEM_BYTES2 = SMEM_TILES2 + 512 + SMEM_XCHG;
constexpr uint32_t TM_S = 0, TM_O = 256, TM_P = 384;  // S: + 128 x;  O: + 64 x;  P (bf16 pairs): + 64 x
constexpr float RESCALE_LOG2 = 8.0f;
constexpr int MAX_BLOCKS_PER_CTA = 16;

// KV stream of one 256-query block: tiles u = 0..U-1 at sequence rows kv_base + 128 u; Q tile x consumes
// u in [lo[x], hi[x])
struct BlockRange {
  int kv_base, U;
  int lo[2], hi[2];
};
__device__ __forceinline__ BlockRange block_range(int q0, int len, int window) {
  BlockRange r;
  r.kv_base = window >= 0 ? max(0, q0 - window) : 0;
#pragma unroll
  for (int x = 0; x < 2; ++x) {
    const int qx = q0 + x * BQ;
    int first = 0, last = len - 1;
    if (window >= 0) {
      first = max(0, qx - window);
      last = min(len - 1, qx + BQ - 1 + window);
    }
    r.lo[x] = (first - r.kv_base) / BKV;
    r.hi[x] = qx < len ? (last - r.kv_base) / BKV + 1 : r.lo[x];
  }
  r.U = max(r.hi[0], r.hi[1]);
  return r;
}

__global__ void __launch_bounds__(THREADS2, 1)
attn_fwd_v2_kernel(const __grid_constant__ CUtensorMap tma_qkv, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int seq = blockIdx.x / p.ctas_per_seq, head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int b_begin = (blockIdx.x % p.ctas_per_seq) * p.blocks_per_cta;
  const int n_b = min(p.blocks_per_cta, (len + 2 * BQ - 1) / (2 * BQ) - b_begin);  // 256-query blocks of this CTA
  if (n_b <= 0) return;

  uint8_t* smem_q = smem;                                   // [2 blocks][2 tiles][16 KB]
  uint8_t* smem_k = smem + 4 * Q_BYTES;                     // [4][16 KB]
  uint8_t* smem_v = smem_k + KV_STAGES2 * KV_TILE_BYTES;    // [4][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SMEM_TILES2);
  uint64_t* q_full = bars;                 // [2]
  uint64_t* q_empty = q_full + 2;          // [2] 2 arrivals: both issuers have retired their S MMAs of the block
  uint64_t* kv_full = q_empty + 2;         // [4]
  uint64_t* kv_empty = kv_full + KV_STAGES2;  // [4] 2 arrivals (both issuers)
  uint64_t* s_full = kv_empty + KV_STAGES2;   // [2]
  uint64_t* s_free = s_full + 2;           // [2] 8 arrivals (one per warp): S_x is in registers
  uint64_t* p_full = s_free + 2;           // [2][2] per 64-key half of P_x (in TMEM), 128 arrivals each
  uint64_t* pv_done = p_full + 4;          // [2][2] PV_x(u, half) retired: that half of the P_x buffer is reusable
  uint64_t* o_full = pv_done + 4;          // [2] per Q tile
  uint64_t* o_free = o_full + 2;           // [2] 8 arrivals: O_x of the block is in registers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
  float* smem_xchg = reinterpret_cast<float*>(smem + SMEM_TILES2 + 512);  // [max even tiles | max odd tiles | sums][2 Q tiles][2 halves][128 rows]

  if (threadIdx.x == 0 && (ptx::smem_u32(smem) & 1023u) != 0) __trap();

  if (warp == SM_WARPS + 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);
      ptx::mbar_init(&q_empty[i], 2);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_free[i], 8);
    }
    for (int s = 0; s < KV_STAGES2; ++s) {
      ptx::mbar_init(&kv_full[s], 1);
      ptx::mbar_init(&kv_empty[s], 2);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&p_full[i], 128);
      ptx::mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&o_free[i], 8);
    }
    ptx::fence_barrier_init();
  }
  if (warp == SM_WARPS) {
    if (lane == 0) ptx::prefetch_tmap(&tma_qkv);
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // the data-movement warpgroup hands registers to the softmax ones
  if (warp >= SM_WARPS) {
    ptx::setmaxnreg_dec<REG_AUX2>();
    if (warp == SM_WARPS && lane == 0) {
      // ---------------------------------------------------------------- TMA producer
      const int col_q = head * D, col_k = p.hidden + head * D, col_v = 2 * p.hidden + head * D;
      int ring = 0;
      for (int bi = 0; bi < n_b; ++bi) {
        const int q0 = (b_begin + bi) * 2 * BQ;
        const int qb = bi & 1;
        ptx::mbar_wait_trap(&q_empty[qb], ((bi >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[qb], 2 * Q_BYTES);
        ptx::tma_load_2d(smem_q + qb * 2 * Q_BYTES, &tma_qkv, &q_full[qb], col_q, seq_start + q0);
        ptx::tma_load_2d(smem_q + qb * 2 * Q_BYTES + Q_BYTES, &tma_qkv, &q_full[qb], col_q, seq_start + q0 + BQ);
        const BlockRange br = block_range(q0, len, p.window);
        for (int u = 0; u < br.U; ++u, ++ring) {
          const int s = ring % KV_STAGES2;
          ptx::mbar_wait_trap(&kv_empty[s], ((ring / KV_STAGES2) & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&kv_full[s], 2 * KV_TILE_BYTES);
          const int row = seq_start + br.kv_base + u * BKV;
          ptx::tma_load_2d(smem_k + s * KV_TILE_BYTES, &tma_qkv, &kv_full[s], col_k, row);
          ptx::tma_load_2d(smem_v + s * KV_TILE_BYTES, &tma_qkv, &kv_full[s], col_v, row);
        }
      }
    } else if (warp == SM_WARPS + 1 || warp == SM_WARPS + 2) {
      // ---------------------------------------------------------------- MMA issuers (one per Q tile)
      // Each Q tile has its own issuing warp, so its MMAs follow the order in which its softmax group
      // produces the events: "S_x buffer drained into registers" -> S_x(t+1) (runs on the tensor pipe WHILE
      // the group is still exponentiating S_x(t)), "first / second 64-key half of P_x(t) written" -> PV on
      // that half.  All waits are blocking mbarrier waits: a polling issuer steals issue slots from the
      // softmax warps that share its scheduler (measured: -20 %).  The whole warp runs the loop (warp-uniform
      // control flow keeps the descriptor arithmetic in the uniform datapath: back-to-back UTCHMMA instead
      // of ~10 address-move instructions per MMA); one elected lane issues.
      const bool leader = ptx::elect_one();
      const int x = warp - (SM_WARPS + 1);
      const uint32_t idesc_s = ptx::umma_idesc_bf16(BQ, BKV, 0, 0);
      const uint32_t idesc_o = ptx::umma_idesc_bf16(BQ, D, 0, 1);  // V is MN-major
      const uint32_t t_p = tmem_base + TM_P + x * 64;
      constexpr uint32_t HI = ptx::umma_desc_hi_sw128(1024);
      int blk = 0;  // blocks of this Q tile that had work so far (phase of its o_full / o_free)
      const uint32_t t_s = tmem_base + TM_S + x * 128;
      int ring_base = 0;  // ring position of tile 0 of the current block
      int it = 0;         // tiles this Q-tile stream has consumed so far (phase of its s / p barriers)
#pragma unroll 1
      for (int bi = 0; bi < n_b; ++bi) {
        const BlockRange br = block_range((b_begin + bi) * 2 * BQ, len, p.window);
        const int lo_x = br.lo[x], hi_x = br.hi[x];
        const int qb = bi & 1;
        const uint32_t q_addr = ptx::smem_u32(smem_q + qb * 2 * Q_BYTES + x * Q_BYTES);
        const uint32_t t_o = tmem_base + TM_O + x * 64;
        auto issue_s = [&](int t) {
          const int r = ring_base + t;
          const int s = r % KV_STAGES2;
          ptx::mbar_wait_trap(&kv_full[s], (r / KV_STAGES2) & 1);
          ptx::tc_fence_after();
          const uint32_t k_addr = ptx::smem_u32(smem_k + s * KV_TILE_BYTES);
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            if (leader)
              ptx::umma_bf16(t_s, ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                             ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
          if (leader) {
            ptx::umma_commit(&s_full[x]);
            if (t + 1 == hi_x) ptx::umma_commit(&q_empty[qb]);  // last S of the block: Q is not read again
          }
        };
        auto ack = [&](int u) {  // tile outside this Q tile's range: hand the stage back once it has landed
          const int r = ring_base + u;
          ptx::mbar_wait_trap(&kv_full[r % KV_STAGES2], (r / KV_STAGES2) & 1);
          if (leader) ptx::umma_commit(&kv_empty[r % KV_STAGES2]);
        };
        ptx::mbar_wait_trap(&q_full[qb], (bi >> 1) & 1);
        for (int u = 0; u < lo_x; ++u) ack(u);
        if (hi_x > lo_x) {
          // the S buffer was drained when the group loaded the last tile of the previous block (s_free of tile
          // it-1 has completed long ago)
          issue_s(lo_x);
          // O_x of the previous block must have been copied out before the first PV overwrites it
          if (blk > 0) ptx::mbar_wait_trap(&o_free[x], (blk - 1) & 1);
          ptx::tc_fence_after();
        } else if (leader) {
          ptx::umma_commit(&q_empty[qb]);
        }
#pragma unroll 1
        for (int t = lo_x; t < hi_x; ++t, ++it) {
          const int j = t - lo_x;
          if (t + 1 < hi_x) {
            ptx::mbar_wait_trap(&s_free[x], it & 1);
            issue_s(t + 1);
          }
          // V tile as the MN-major B operand (LBO = 8192: distance of 64-element MN chunks, unused)
          const uint32_t v_lo = ptx::umma_desc_lo(ptx::smem_u32(smem_v + ((ring_base + t) % KV_STAGES2) * KV_TILE_BYTES), 8192);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            ptx::mbar_wait_trap(&p_full[2 * x + h], it & 1);  // this half of P_x(t) is in TMEM (O_x rescaled if needed)
            ptx::tc_fence_after();
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 16-key steps: 8 TMEM columns of P (A operand), 16 rows of V
              if (leader)
                ptx::umma_bf16_ts(t_o, t_p + (4 * h + k) * 8, v_lo + (((4 * h + k) * 2048) >> 4), HI, idesc_o,
                                  (j | h | k) != 0 ? 1u : 0u);
            if (leader) {
              ptx::umma_commit(&pv_done[2 * x + h]);
              if (h == 1 && t + 1 == hi_x) ptx::umma_commit(&o_full[x]);
            }
          }
          if (leader) ptx::umma_commit(&kv_empty[(ring_base + t) % KV_STAGES2]);
        }
        for (int u = max(hi_x, lo_x); u < br.U; ++u) ack(u);
        ring_base += br.U;
        if (hi_x > lo_x) ++blk;
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax groups
    // 8 warps per Q tile: TMEM lane quadrant (warp & 3) x 64-key column half ((warp >> 2) & 1).  A query row
    // is shared by two threads (one per half); they exchange their partial row maximum through shared memory
    // once per tile (named barrier of the two warps) and their partial row sums once per block.  Four light
    // warps per scheduler instead of two heavy ones: the exp / FMA latencies of one warp are covered by the
    // others (the exp phase of the 2-warp version reached 62 % of the MUFU rate).
    ptx::setmaxnreg_inc<REG_SM>();
    const int x = warp >> 3;            // 0 = A, 1 = B
    const int hc = (warp >> 2) & 1;     // column half of the 128-key tile
    const int r = (warp & 3) * 32 + lane;  // query row inside the tile == TMEM lane
    const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_s = tmem_base + TM_S + x * 128 + lane_off + hc * 64;
    const uint32_t t_p = tmem_base + TM_P + x * 64 + lane_off + hc * 32;  // this thread's 64 keys of P_x as bf16 pairs
    int blk = 0;  // blocks with work so far (phase of o_full / o_free)
    constexpr int XS = 2 * 2 * BQ;  // floats per exchange plane
    float* my_x = smem_xchg + (x * 2 + hc) * BQ + r;
    const float* peer_x = smem_xchg + (x * 2 + (hc ^ 1)) * BQ + r;
    const int pair_bar = 1 + x * 4 + (warp & 3);  // named barrier of the two warps that share these 32 rows
    const float c = p.scale_log2;
    const float2 c2 = make_float2(c, c);
    int it = 0;  // tiles consumed so far by this group (phase of its barriers)
    // -DCM3P_ATTN_PROF: clock64 spent per phase by one thread of CTA (0,0,0), printed at the end (tools/attn_one.py)
#ifdef CM3P_ATTN_PROF
    long long pf_s = 0, pf_ld = 0, pf_max = 0, pf_exp = 0, pf_epi = 0, pf_t0 = clock64(), pf_a = pf_t0, pf_b;
#define PF_B(acc) do { pf_b = clock64(); acc += pf_b - pf_a; pf_a = pf_b; } while (0)
#else
#define PF_B(acc)
#endif
    for (int bi = 0; bi < n_b; ++bi) {
      const int q0 = (b_begin + bi) * 2 * BQ;
      const BlockRange br = block_range(q0, len, p.window);
      const int qi = q0 + x * BQ + r;
      const bool valid = qi < len;
      const uint32_t t_o = tmem_base + TM_O + x * 64 + lane_off + hc * 32;  // this thread's 32 of the 64 O columns
      float m_run = -INFINITY, l = 0.f;
      const int lo_x = br.lo[x];
      const int n_iter = br.hi[x] - lo_x;

      for (int jj = 0; jj < n_iter; ++jj, ++it) {
        const int kv0 = br.kv_base + (lo_x + jj) * BKV;
        PF_B(pf_epi);
        ptx::mbar_wait_trap(&s_full[x], it & 1);
        PF_B(pf_s);
        ptx::tc_fence_after();
        uint32_t sr[2][32];
        ptx::tmem_ld_32x32b_x32(t_s, sr[0]);
        ptx::tmem_ld_32x32b_x32(t_s + 32, sr[1]);
        ptx::tmem_ld_wait();
        // the scores are in registers: hand the S buffer back so S_x(jj+1) overlaps this tile's softmax
        ptx::tc_fence_before();
        if (lane == 0) ptx::mbar_arrive(&s_free[x]);
        PF_B(pf_ld);
        // Masking state per 32-column chunk and per warp (32 consecutive query rows): 0 = no allowed key
        // for any row of the warp (skip: no exp, P = 0), 1 = some rows partially masked, 2 = fully allowed.
        // Interior tiles of global layers take the branch-free path (every chunk fully allowed).
        const bool masked_tile = (p.window >= 0) || (kv0 + BKV > len);  // CTA-uniform
        int state[2] = {2, 2};
        float mxq[2] = {-INFINITY, -INFINITY};
        if (!masked_tile) {
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              ma = ptx::max3(ma, __uint_as_float(sr[q][i]), __uint_as_float(sr[q][i + 1]));
              mb = ptx::max3(mb, __uint_as_float(sr[q][i + 2]), __uint_as_float(sr[q][i + 3]));
            }
            mxq[q] = fmaxf(ma, mb);
          }
        } else {
          int a = 0, b = min(BKV, len - kv0);
          const int qw = q0 + x * BQ + (warp & 3) * 32;  // first query row of this warp
          int wa = 0, wb = b;                            // union of the warp's allowed ranges
          int ia = 0, ib = b;                            // intersection
          if (p.window >= 0) {
            a = max(a, qi - p.window - kv0);
            b = min(b, qi + p.window + 1 - kv0);
            wa = max(wa, qw - p.window - kv0);
            wb = min(wb, qw + 31 + p.window + 1 - kv0);
            ia = max(ia, qw + 31 - p.window - kv0);
            ib = min(ib, qw + p.window + 1 - kv0);
          }
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int c0 = hc * 64 + q * 32, c1 = c0 + 32;
            state[q] = (c1 <= wa || c0 >= wb) ? 0 : ((c0 >= ia && c1 <= ib) ? 2 : 1);
            if (state[q] == 1) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int kj = c0 + i;
                if (kj < a || kj >= b) sr[q][i] = 0xff800000u;  // -inf
              }
            }
            if (state[q] != 0) {
              float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                ma = ptx::max3(ma, __uint_as_float(sr[q][i]), __uint_as_float(sr[q][i + 1]));
                mb = ptx::max3(mb, __uint_as_float(sr[q][i + 2]), __uint_as_float(sr[q][i + 3]));
              }
              mxq[q] = fmaxf(ma, mb);
            }
          }
        }
        // row maximum over all 128 keys: exchange the two halves.  One plane per tile parity: the peer may still
        // be reading this tile's slot when the next tile's maximum is written; the slot of tile t is rewritten at
        // tile t + 2, after both warps have passed the barrier of tile t + 1 (and so have read tile t's value)
        my_x[(it & 1) * XS] = fmaxf(mxq[0], mxq[1]);
        ptx::named_bar_sync(pair_bar, 64);
        const float m_new = ptx::max3(m_run, fmaxf(mxq[0], mxq[1]), peer_x[(it & 1) * XS]);
        PF_B(pf_max);
        if (jj == 0) {
          m_run = m_new;
        } else {
          const bool grow = (m_new - m_run) * c > RESCALE_LOG2;  // also true for -inf -> finite
          if (__any_sync(0xffffffffu, grow)) {  // same vote in both warps of the pair: same rows, same maxima
            // both halves of PV(jj-1) must have retired before O is rescaled
            ptx::mbar_wait_trap(&pv_done[2 * x], (it - 1) & 1);
            ptx::mbar_wait_trap(&pv_done[2 * x + 1], (it - 1) & 1);
            ptx::tc_fence_after();
            const float alpha = grow ? ptx::ex2_approx((m_run - m_new) * c) : 1.f;
#pragma unroll 1
            for (int h = 0; h < 32; h += 16) {
              uint32_t o[16];
              ptx::tmem_ld_32x32b_x16(t_o + h, o);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
              ptx::tmem_st_32x32b_x16(t_o + h, o);
            }
            ptx::tmem_st_wait();
            l *= alpha;
            if (grow) m_run = m_new;
            // PV of this tile accumulates into all 64 columns of O as soon as EITHER half of P is complete: the
            // peer warp (same vote, same branch) must not get there before this warp's 32 columns are rescaled
            ptx::tc_fence_before();
            ptx::named_bar_sync(pair_bar, 64);
            ptx::tc_fence_after();
          }
        }
        const float mc = (m_run == -INFINITY) ? 0.f : m_run * c;
        const float2 nmc2 = make_float2(-mc, -mc);
        float2 rs[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};  // independent partial row sums (ILP)
        // this half of P_x in TMEM is still being read by the PV of the previous tile until pv_done fires; that
        // MMA was issued a whole exp phase ago, so this wait is normally already satisfied
        if (it > 0) ptx::mbar_wait_trap(&pv_done[2 * x + hc], (it - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          uint32_t packed[16];
          if (masked_tile && state[q] == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) packed[i] = 0u;
          } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const float2 t = ptx::fma2(make_float2(__uint_as_float(sr[q][i]), __uint_as_float(sr[q][i + 1])), c2, nmc2);
              const float2 e = ptx::ex2_pair(t, i >> 1);
              rs[(i >> 1) & 3] = ptx::add2(rs[(i >> 1) & 3], e);
              packed[i >> 1] = ptx::pack_bf16x2(e.x, e.y);
            }
          }
          ptx::tmem_st_32x32b_x16(t_p + q * 16, packed);
        }
        ptx::tmem_st_wait();
        // this 64-key half of P is complete: its PV can start while the other half is still being computed
        ptx::tc_fence_before();
        ptx::mbar_arrive(&p_full[2 * x + hc]);
        const float2 rsum = ptx::add2(ptx::add2(rs[0], rs[1]), ptx::add2(rs[2], rs[3]));
        l += rsum.x + rsum.y;
        PF_B(pf_exp);
      }
      if (n_iter > 0) {
        // epilogue of the block; meanwhile the issuer already runs S of the next block's first tile.
        // Row sum over both halves (third plane of the exchange area).
        my_x[2 * XS] = l;
        ptx::named_bar_sync(pair_bar, 64);
        l += peer_x[2 * XS];
        ptx::named_bar_sync(pair_bar, 64);  // both sums are read before the next block's are written
        ptx::mbar_wait_trap(&o_full[x], blk & 1);
        ptx::tc_fence_after();
        uint32_t rr[32];
        ptx::tmem_ld_32x32b_x32(t_o, rr);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&o_free[x]);
        ++blk;
        if (valid) {
          const float inv = 1.f / l;
          const int64_t row = static_cast<int64_t>(seq_start) + qi;
          __nv_bfloat16* dst = p.out + row * p.hidden + head * D + hc * 32;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            uint4 uo;
            uo.x = ptx::pack_bf16x2(__uint_as_float(rr[i]) * inv, __uint_as_float(rr[i + 1]) * inv);
            uo.y = ptx::pack_bf16x2(__uint_as_float(rr[i + 2]) * inv, __uint_as_float(rr[i + 3]) * inv);
            uo.z = ptx::pack_bf16x2(__uint_as_float(rr[i + 4]) * inv, __uint_as_float(rr[i + 5]) * inv);
            uo.w = ptx::pack_bf16x2(__uint_as_float(rr[i + 6]) * inv, __uint_as_float(rr[i + 7]) * inv);
            *reinterpret_cast<uint4*>(dst + i) = uo;
          }
          if (p.lse && hc == 0) p.lse[static_cast<int64_t>(head) * p.total_tokens + row] = m_run * c + log2f(l);
        }
      }
    }
#ifdef CM3P_ATTN_PROF
    PF_B(pf_epi);
    if (lane == 0 && blockIdx.x == 0 && blockIdx.y == 0)
      printf("attn fwd warp %2d: blocks=%d tiles=%d total=%lld wait_s=%lld ld=%lld max+exchange=%lld exp+store=%lld "
             "epilogue=%lld\n", warp, n_b, it, clock64() - pf_t0, pf_s, pf_ld, pf_max, pf_exp, pf_epi);
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == SM_WARPS) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace v2
}  // namespace

int attn_pack_groups(const int32_t* cu_seqlens, int batch, int32_t* groups, int32_t* n_groups, int max_groups,
                     cudaStream_t stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  CM3P_REQUIRE(cu_seqlens && groups && n_groups && batch > 0 && max_groups > 0, kBadShape,
               "attn_pack_groups: null pointer or empty batch");
  CM3P_REQUIRE((reinterpret_cast<uintptr_t>(groups) & 7) == 0, kBadAlignment, "attn_pack_groups: groups must be 8-byte aligned");
  CM3P_CUDA_TRY(cudaMemsetAsync(n_groups, 0, sizeof(int32_t), stream));
  const int chunks = (batch + packed::CHUNK_SEQS - 1) / packed::CHUNK_SEQS;
  packed::attn_pack_groups_kernel<<<(chunks + 7) / 8, 256, 0, stream>>>(cu_seqlens, batch, reinterpret_cast<int2*>(groups),
                                                                      n_groups, max_groups);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

int attn_varlen_fwd(const AttnFwdArgs& a, cudaStream_t stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  CM3P_REQUIRE(a.head_dim == 64, kBadShape, "attn: head_dim %d unsupported (kernel is specialised for 64)", a.head_dim);
  CM3P_REQUIRE(a.batch > 0 && a.heads > 0 && a.total_tokens > 0 && a.max_seqlen > 0, kBadShape,
               "attn: empty problem (batch=%d heads=%d tokens=%lld max_seqlen=%d)", a.batch, a.heads,
               (long long)a.total_tokens, a.max_seqlen);
  CM3P_REQUIRE(a.qkv && a.out && a.cu_seqlens, kBadShape, "attn: null pointer");
  CM3P_REQUIRE(a.heads <= 65535, kBadShape, "attn: heads=%d exceeds the grid limit", a.heads);
  const int H = a.heads * 64;
  CUtensorMap tmap;
  rc = encode_tmap_2d_bf16(&tmap, a.qkv, 3 * (uint64_t)H, (uint64_t)a.total_tokens, 3 * (uint64_t)H * 2, 64, BKV);
  if (rc != kOk) return rc;
  if (a.groups) {
    // packed short sequences: one CTA per (group of sequences with <= 128 tokens in total, head)
    CM3P_REQUIRE(a.n_groups && a.max_groups > 0, kBadShape, "attn(packed): n_groups / max_groups missing");
    CM3P_REQUIRE(a.max_seqlen <= BQ && a.window < 0, kBadShape,
                 "attn(packed): needs max_seqlen <= 128 (got %d) and a global layer (window %d)", a.max_seqlen, a.window);
    CM3P_ENSURE_DYN_SMEM(packed::attn_fwd_packed_kernel, packed::P_SMEM_BYTES);
    packed::PackedParams pp;
    pp.cu_seqlens = a.cu_seqlens;
    pp.groups = reinterpret_cast<const int2*>(a.groups);
    pp.n_groups = a.n_groups;
    pp.out = reinterpret_cast<__nv_bfloat16*>(a.out);
    pp.lse = a.lse;
    pp.total_tokens = a.total_tokens;
    pp.hidden = H;
    pp.scale_log2 = 0.125f * 1.4426950408889634f;
    dim3 grid(a.max_groups, a.heads, 1);
    packed::attn_fwd_packed_kernel<<<grid, THREADS, packed::P_SMEM_BYTES, stream>>>(tmap, pp);
    CM3P_CUDA_TRY(cudaGetLastError());
    return kOk;
  }
  Params p;
  p.cu_seqlens = a.cu_seqlens;
  p.out = reinterpret_cast<__nv_bfloat16*>(a.out);
  p.lse = a.lse;
  p.total_tokens = a.total_tokens;
  p.heads = a.heads;
  p.hidden = H;
  p.window = a.window;
  p.scale_log2 = 0.125f * 1.4426950408889634f;
  p.blocks_per_cta = 1;
  // short sequences fit one 128-row tile: the light 2-CTA/SM kernel wins there
  if (get_option(kOptAttnForceTileKernels) || a.max_seqlen <= BQ) {
    CM3P_ENSURE_DYN_SMEM(attn_fwd_sm100_kernel, SMEM_BYTES);
    p.ctas_per_seq = (a.max_seqlen + BQ - 1) / BQ;
    CM3P_REQUIRE(static_cast<int64_t>(p.ctas_per_seq) * a.batch <= 0x7fffffffLL, kBadShape, "attn: grid too large");
    dim3 grid(static_cast<unsigned>(p.ctas_per_seq) * a.batch, a.heads, 1);
    attn_fwd_sm100_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmap, p);
  } else {
    CM3P_ENSURE_DYN_SMEM(v2::attn_fwd_v2_kernel, v2::SMEM_BYTES2);
    // blocks per CTA: enough CTAs for ~16 (global) / ~4 (window) waves of uneven work
    const int forced_bpc = get_option(kOptFwdBlocksPerCta);
    const int64_t units = (a.total_tokens / (2 * BQ) + a.batch / 2 + 1) * a.heads;
    const int64_t target_ctas = static_cast<int64_t>(num_sms()) * (a.window >= 0 ? 4 : 16);
    int bpc = static_cast<int>((units + target_ctas - 1) / target_ctas);
    bpc = bpc < 1 ? 1 : (bpc > v2::MAX_BLOCKS_PER_CTA ? v2::MAX_BLOCKS_PER_CTA : bpc);
    if (forced_bpc > 0) bpc = forced_bpc > v2::MAX_BLOCKS_PER_CTA ? v2::MAX_BLOCKS_PER_CTA : forced_bpc;
    p.blocks_per_cta = bpc;
    p.ctas_per_seq = (a.max_seqlen + 2 * BQ * bpc - 1) / (2 * BQ * bpc);
    CM3P_REQUIRE(static_cast<int64_t>(p.ctas_per_seq) * a.batch <= 0x7fffffffLL, kBadShape, "attn: grid too large");
    dim3 grid(static_cast<unsigned>(p.ctas_per_seq) * a.batch, a.heads, 1);
    v2::attn_fwd_v2_kernel<<<grid, v2::THREADS2, v2::SMEM_BYTES2, stream>>>(tmap, p);
  }
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
