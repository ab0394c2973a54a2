// Internal C++ interface of the tcgen05 GEMM (the C ABI wrapper is in api.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

// Fused epilogues.  Values are part of the C ABI (include/cm3p_b200.h: CM3P_EPI_*).
enum Epilogue : int {
  EPI_STORE = 0,       // C = acc                                  (bf16)
  EPI_RESIDUAL = 1,    // C = acc + aux[M,N]                       (bf16 residual stream update)
  EPI_GELU = 2,        // C = gelu_erf(acc)                        (audio projector linear_1)
  EPI_BIAS_GELU = 3,   // C = gelu_erf(acc + bias[N])              (conv1d as GEMM over im2col rows)
  EPI_BIAS = 4,        // C = acc + bias[N]                        (MLM decoder)
  EPI_GEGLU = 5,       // C[M,N/2] = gelu_erf(u) * g, B rows interleaved (16 u, 16 g)
  EPI_GEGLU_SAVE = 6,  // same, and C2[M,N] = raw acc (kept for the backward pass)
  EPI_ROPE = 7,        // rotate-half RoPE on columns [0, rope_cols) per 64-wide head, rest copied
  EPI_SCALE_F32 = 8,   // C(fp32) = scale * acc (+ C when accumulate)   (logits, weight gradients)
  EPI_COUNT = 9,
};

struct GemmArgs {
  const void* a = nullptr;  // trans_a == 0: [M, K] K contiguous;  trans_a == 1: [K, M] M contiguous
  int64_t lda = 0;          // row pitch in elements
  const void* b = nullptr;  // trans_b == 0: [N, K] K contiguous (nn.Linear weight);  1: [K, N] N contiguous
  int64_t ldb = 0;
  void* c = nullptr;
  int64_t ldc = 0;
  int64_t M = 0, N = 0, K = 0;
  int epilogue = EPI_STORE;
  const void* aux = nullptr;
  int64_t ld_aux = 0;
  void* c2 = nullptr;
  int64_t ldc2 = 0;
  float scale = 1.f;
  const int32_t* positions = nullptr;
  const float* rope_table = nullptr;
  int64_t rope_cols = 0;
  int trans_a = 0, trans_b = 0;
  int accumulate = 0;
  // LayerNorm folded into the GEMMs on both sides of it (bf16 staged epilogues only):
  //   producer (EPI_RESIDUAL): stats_out[M][2] += (sum, sum of squares) of the bf16 rows it writes
  //   consumer (EPI_ROPE / EPI_GEGLU / EPI_GEGLU_SAVE), B = W . diag(gamma):
  //       acc <- rstd_row * (acc - mean_row * col_corr[n]),  col_corr[n] = sum_k B[n][k]
  float* stats_out = nullptr;
  const float* row_stats = nullptr;
  const float* col_corr = nullptr;
  float ln_eps = 1e-5f;
  // Ordered split-K accumulation (EPI_SCALE_F32 with accumulate): `tile_sem` points to `tile_sem_count` zeroed
  // int32 counters that the kernel uses as per-output-tile turnstiles and leaves zeroed; the K splits of a tile then
  // add their partial sums in split order and the result is bit-reproducible.  null: fp32 atomics (red.global.add).
  int32_t* tile_sem = nullptr;
  int64_t tile_sem_count = 0;
  // Grouped GEMM: M / group_m independent problems (group_m x N x K each) in one launch.  Operands are stacked
  // along their outer dimension: A [M, K] (trans_a: [groups*K, group_m]), B [groups*N, K] (trans_b: [groups*K, N]),
  // C / aux [M, N].  group_m % 256 == 0, K % 64 == 0.  0 = one problem.
  int64_t group_m = 0;
  // Implicit-GEMM conv1d, kernel 3, padding 1 (cm3p/modeling_cm3p.py:488-489, :501-502).  x = channels-last bf16
  // input [conv_batch, conv_frames, conv_cin]; conv_cpad = conv_cin rounded up to 64 = channels per tap in the
  // packed weight [C_out, 3 * conv_cpad] (k index = tap * conv_cpad + c, zero columns for c >= conv_cin).
  //   conv_mode 1 (forward):  a = x, b = packed weight, c = out [conv_batch, conv_frames / stride, N] bf16,
  //                           epilogue EPI_BIAS / EPI_BIAS_GELU (aux = bias); M, K are derived.
  //   conv_mode 2 (weight gradient): a = dz [conv_batch, conv_frames / stride, M] bf16, b = x,
  //                           c = dW [M, 3 * conv_cpad] fp32 (+=), epilogue EPI_SCALE_F32 with accumulate; N, K derived.
  int conv_mode = 0;
  int conv_stride = 1;
  int conv_batch = 0, conv_frames = 0, conv_cin = 0, conv_cpad = 0;
};

int gemm_bf16(const GemmArgs& args, cudaStream_t stream);

}  // namespace cm3p
