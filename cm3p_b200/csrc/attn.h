// Internal C++ interface of the varlen attention kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

struct AttnFwdArgs {
  const void* qkv = nullptr;        // [T, 3, heads, 64] bf16, q/k already rotated
  void* out = nullptr;              // [T, heads*64] bf16
  float* lse = nullptr;             // [heads, T] fp32 (log2 domain) or nullptr
  const int32_t* cu_seqlens = nullptr;  // [batch+1] device
  int64_t total_tokens = 0;
  int batch = 0;
  int heads = 0;
  int head_dim = 64;
  int max_seqlen = 0;
  int window = -1;  // < 0: global attention; otherwise attend iff |i-j| <= window
  // packed short sequences (all three or none): group table from attn_pack_groups, needs max_seqlen <= 128
  const int32_t* groups = nullptr;    // [max_groups][2] (first sequence, end sequence)
  const int32_t* n_groups = nullptr;  // device scalar
  int max_groups = 0;
};

// Packs consecutive sequences into groups of <= 128 tokens (the unit of work of the packed attention kernels).
// groups: [max_groups][2] int32, n_groups: device scalar (zeroed here).  max_groups must be at least
// min(batch, 2 * (total_tokens / 128) + ceil(batch / 64) + 1).
int attn_pack_groups(const int32_t* cu_seqlens, int batch, int32_t* groups, int32_t* n_groups, int max_groups,
                     cudaStream_t stream);

int attn_varlen_fwd(const AttnFwdArgs& args, cudaStream_t stream);

struct AttnBwdArgs {
  const void* qkv = nullptr;    // [T, 3, heads, 64] bf16 as given to the forward (q/k rotated)
  const void* out = nullptr;    // [T, heads*64] bf16 forward output
  const void* dout = nullptr;   // [T, heads*64] bf16
  const float* lse = nullptr;   // [heads, T] fp32 from the forward (log2 domain)
  float* delta = nullptr;       // [heads, T] fp32 workspace
  void* dqkv = nullptr;         // [T, 3, heads, 64] bf16 gradient w.r.t. the un-rotated Wqkv output
  const int32_t* cu_seqlens = nullptr;
  const int32_t* positions = nullptr;  // [T] + rope_table: undo the forward's RoPE on dq/dk (both or neither)
  const float* rope_table = nullptr;
  int64_t total_tokens = 0;
  int batch = 0;
  int heads = 0;
  int head_dim = 64;
  int max_seqlen = 0;
  int window = -1;
  const int32_t* groups = nullptr;  // packed short sequences, as in AttnFwdArgs
  const int32_t* n_groups = nullptr;
  int max_groups = 0;
};

int attn_varlen_bwd(const AttnBwdArgs& args, cudaStream_t stream);
int attn_varlen_bwd_v3(const AttnBwdArgs& args, cudaStream_t stream);
int attn_varlen_bwd_window(const AttnBwdArgs& args, cudaStream_t stream);  // band walk: window <= 64, one kernel, 5 GEMMs  // 128x128 tiles, early S/dP issue (long sequences)

}  // namespace cm3p
