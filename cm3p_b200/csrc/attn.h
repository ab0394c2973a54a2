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
};

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
};

int attn_varlen_bwd(const AttnBwdArgs& args, cudaStream_t stream);
int attn_varlen_bwd_v2(const AttnBwdArgs& args, cudaStream_t stream);  // 1 CTA/SM ping-pong kernels (CM3P_ATTN_BWD=v2)
int attn_varlen_bwd_v3(const AttnBwdArgs& args, cudaStream_t stream);  // 128x128 tiles, early S/dP issue (long sequences)

}  // namespace cm3p
