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

}  // namespace cm3p
