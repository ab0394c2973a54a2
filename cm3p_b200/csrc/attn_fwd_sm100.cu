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
};

__global__ void __launch_bounds__(THREADS, 2)
attn_fwd_sm100_kernel(const __grid_constant__ CUtensorMap tma_qkv, const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int seq = blockIdx.z;
  const int head = blockIdx.y;
  const int seq_start = p.cu_seqlens[seq];
  const int len = p.cu_seqlens[seq + 1] - seq_start;
  const int q0 = blockIdx.x * BQ;
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
        ptx::mbar_wait(&kv_empty[s], ph ^ 1);
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
        ptx::mbar_wait(&kv_full[s], (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t k_addr = ptx::smem_u32(smem_k + s * KV_TILE_BYTES);
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16(tmem_base + TMEM_S, ptx::umma_smem_desc_sw128(q_addr + k * 32, 16, 1024),
                         ptx::umma_smem_desc_sw128(k_addr + k * 32, 16, 1024), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      };
      ptx::mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < n_tiles; ++j) {
        ptx::mbar_wait(p_full, j & 1);  // P_j in smem, S_j and O_{j-1} consumed
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
      ptx::mbar_wait(s_full, j & 1);
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
        ptx::mbar_wait(o_full, (j - 1) & 1);
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
    ptx::mbar_wait(o_full, (n_tiles - 1) & 1);
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

}  // namespace

int attn_varlen_fwd(const AttnFwdArgs& a, cudaStream_t stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  CM3P_REQUIRE(a.head_dim == 64, kBadShape, "attn: head_dim %d unsupported (kernel is specialised for 64)", a.head_dim);
  CM3P_REQUIRE(a.batch > 0 && a.heads > 0 && a.total_tokens > 0 && a.max_seqlen > 0, kBadShape,
               "attn: empty problem (batch=%d heads=%d tokens=%lld max_seqlen=%d)", a.batch, a.heads,
               (long long)a.total_tokens, a.max_seqlen);
  CM3P_REQUIRE(a.qkv && a.out && a.cu_seqlens, kBadShape, "attn: null pointer");
  const int H = a.heads * 64;
  CUtensorMap tmap;
  rc = encode_tmap_2d_bf16(&tmap, a.qkv, 3 * (uint64_t)H, (uint64_t)a.total_tokens, 3 * (uint64_t)H * 2, 64, BKV);
  if (rc != kOk) return rc;
  static bool configured = false;
  if (!configured) {
    CM3P_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_sm100_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
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
  dim3 grid((a.max_seqlen + BQ - 1) / BQ, a.heads, a.batch);
  attn_fwd_sm100_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmap, p);
  CM3P_CUDA_TRY(cudaGetLastError());
  return kOk;
}

}  // namespace cm3p
