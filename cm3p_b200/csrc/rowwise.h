// Internal C++ interface of the HBM-bound kernels (rowwise_fwd.cu / rowwise_bwd.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cm3p {

int layernorm_fwd(const void* x, const float* gamma, void* y, float* stats, int64_t rows, int H, float eps,
                  cudaStream_t stream);
int embed_gather_ln(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot, const void* tok_emb,
                    const void* audio_embeds, const float* gamma, void* y, float* stats, int64_t rows, int H,
                    int vocab, float eps, cudaStream_t stream);
int transpose_cast(const float* x, void* out, int B, int C, int F, cudaStream_t stream);  // [B,C,F] fp32 -> [B,F,C] bf16
int gather_rows(const void* x, const int32_t* index, void* out, int rows, int H, cudaStream_t stream);
int mean_pool(const void* x, const int32_t* cu_seqlens, void* out, int batch, int H, cudaStream_t stream);
int l2norm_rows(const float* e, float* out_f32, void* out_bf16, float* inv_norm, int rows, int P,
                cudaStream_t stream);
int clip_loss_fwd(const float* S, const int32_t* true_idx, float* row_lse, float* col_lse, float* loss, int Bm, int V,
                  int Bb, cudaStream_t stream);

int segment_accumulate(const float* e, const int32_t* slot, float* sums, float* counts, int rows, int P,
                       cudaStream_t stream);
int mean_renormalize(const float* sums, const float* counts, float* out, int rows, int P, cudaStream_t stream);

// ---- backward (rowwise_bwd.cu)
int layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* dres, void* dx, float* dgamma,
                  int64_t rows, int H, float eps, cudaStream_t stream);
int embed_gather_ln_bwd(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot, const void* tok_emb,
                        const void* audio_embeds, const float* gamma, const void* dy, float* d_tok_emb,
                        void* d_audio_embeds, float* dgamma, int64_t rows, int H, int vocab, float eps,
                        cudaStream_t stream);
int geglu_bwd(const void* ug, const void* dh, void* dug, void* h, int64_t rows, int I, cudaStream_t stream);
int gelu_fwd(const void* z, void* y, int64_t n, cudaStream_t stream);
int gelu_bwd(const void* z, const void* dy, void* dz, int64_t n, cudaStream_t stream);
int colsum_f32(const void* dy, float* out, int64_t rows, int N, cudaStream_t stream);
int pool_bwd(const void* dpooled, const int32_t* cu_seqlens, void* dhidden, int mode, int accumulate, int batch, int H,
             cudaStream_t stream);
int l2norm_bwd(const float* e, const float* inv_norm, const float* dembeds, void* de_bf16, int rows, int P,
               cudaStream_t stream);
int clip_loss_bwd(const float* S, const int32_t* true_idx, const float* row_lse, const float* col_lse,
                  const float* grad_out, void* dS, int64_t ld_ds, float* dlogit_scale, int Bm, int V, int Bb,
                  cudaStream_t stream);
int vocab_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index, int ignore_index,
                 float* row_lse, float* loss_sum, float* count, int64_t rows, int V, cudaStream_t stream);
int vocab_ce_bwd(void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index, int ignore_index,
                 const float* row_lse, const float* scale, int64_t rows, int V, cudaStream_t stream);
int gather_rows_i32(const void* x, const int32_t* index, void* out, int64_t rows, int H, cudaStream_t stream);
int scatter_add_rows(const void* dx_rows, const int32_t* index, void* dx, int64_t rows, int H, cudaStream_t stream);
int conv2_col2im_gelu_bwd(const void* da2, const void* z1, void* dz1, int B, int F, int C, cudaStream_t stream);

}  // namespace cm3p
