// extern "C" boundary of libcm3p_b200.so (declared in include/cm3p_b200.h).  Nothing here throws;
// every entry returns a status code and records a message for cm3p_last_error().
#include "../../include/cm3p_b200.h"

#include "attn.h"
#include "common.h"
#include "embed_tools.h"
#include "gemm.h"
#include "logmel.h"
#include "optim.h"
#include "rowwise.h"

using namespace cm3p;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

extern "C" {

const char* cm3p_last_error(void) { return last_error(); }
int cm3p_version(void) { return CM3P_B200_VERSION; }
int cm3p_num_sms(void) { return num_sms(); }
int cm3p_set_option(int option, int value) { return set_option(option, value); }
int cm3p_get_option(int option) { return get_option(option); }

int cm3p_gemm_bf16(const void* a, int64_t lda, int trans_a, const void* b, int64_t ldb, int trans_b, void* c,
                   int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* aux, int64_t ld_aux,
                   void* c2, int64_t ldc2, float scale, int accumulate, const int32_t* positions,
                   const float* rope_table, int64_t rope_cols, int32_t* tile_sem, int64_t tile_sem_count,
                   int64_t group_m, void* stream) {
  GemmArgs g;
  g.tile_sem = tile_sem; g.tile_sem_count = tile_sem_count; g.group_m = group_m;
  g.a = a; g.lda = lda; g.trans_a = trans_a;
  g.b = b; g.ldb = ldb; g.trans_b = trans_b;
  g.c = c; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.epilogue = epilogue;
  g.aux = aux; g.ld_aux = ld_aux;
  g.c2 = c2; g.ldc2 = ldc2;
  g.scale = scale; g.accumulate = accumulate;
  g.positions = positions; g.rope_table = rope_table; g.rope_cols = rope_cols;
  return gemm_bf16(g, as_stream(stream));
}

int cm3p_gemm_bf16_ln(const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t M, int64_t N,
                      int64_t K, int epilogue, const void* aux, int64_t ld_aux, void* c2, int64_t ldc2,
                      const int32_t* positions, const float* rope_table, int64_t rope_cols, float* stats_out,
                      const float* row_stats, const float* col_corr, float ln_eps, void* stream) {
  GemmArgs g;
  g.a = a; g.lda = lda;
  g.b = b; g.ldb = ldb;
  g.c = c; g.ldc = ldc;
  g.M = M; g.N = N; g.K = K;
  g.epilogue = epilogue;
  g.aux = aux; g.ld_aux = ld_aux;
  g.c2 = c2; g.ldc2 = ldc2;
  g.positions = positions; g.rope_table = rope_table; g.rope_cols = rope_cols;
  g.stats_out = stats_out; g.row_stats = row_stats; g.col_corr = col_corr; g.ln_eps = ln_eps;
  return gemm_bf16(g, as_stream(stream));
}

int cm3p_attn_pack_groups(const int32_t* cu_seqlens, int batch, int32_t* groups, int32_t* n_groups, int max_groups,
                          void* stream) {
  return attn_pack_groups(cu_seqlens, batch, groups, n_groups, max_groups, as_stream(stream));
}

int cm3p_attn_varlen_fwd(const void* qkv, void* out, float* lse, const int32_t* cu_seqlens, int64_t total_tokens,
                         int batch, int heads, int head_dim, int max_seqlen, int window, const int32_t* groups,
                         const int32_t* n_groups, int max_groups, void* stream) {
  AttnFwdArgs a;
  a.groups = groups; a.n_groups = n_groups; a.max_groups = max_groups;
  a.qkv = qkv; a.out = out; a.lse = lse; a.cu_seqlens = cu_seqlens;
  a.total_tokens = total_tokens; a.batch = batch; a.heads = heads; a.head_dim = head_dim;
  a.max_seqlen = max_seqlen; a.window = window;
  return attn_varlen_fwd(a, as_stream(stream));
}

int cm3p_layernorm_fwd(const void* x, const float* gamma, void* y, float* stats, int64_t rows, int hidden, float eps,
                       void* stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  return layernorm_fwd(x, gamma, y, stats, rows, hidden, eps, as_stream(stream));
}

int cm3p_embed_gather_ln(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot, const void* tok_emb,
                         const void* audio_embeds, const float* gamma, void* y, float* stats, int64_t rows, int hidden,
                         int vocab, float eps, void* stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  return embed_gather_ln(ids, src_index, audio_slot, tok_emb, audio_embeds, gamma, y, stats, rows, hidden, vocab, eps,
                         as_stream(stream));
}

int cm3p_transpose_cast_bf16(const float* x, void* out, int batch, int channels, int frames, void* stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  return transpose_cast(x, out, batch, channels, frames, as_stream(stream));
}

int cm3p_conv1d_k3_fwd(const void* x, const void* weight, const float* bias, void* out, int batch, int c_in, int c_pad,
                       int frames, int c_out, int stride, int gelu, void* stream) {
  GemmArgs g;
  g.conv_mode = 1; g.conv_stride = stride; g.conv_batch = batch; g.conv_frames = frames; g.conv_cin = c_in;
  g.conv_cpad = c_pad;
  g.a = x; g.b = weight; g.c = out;
  g.N = c_out;
  g.M = 1; g.K = 1;  // derived by the launcher
  g.epilogue = gelu ? EPI_BIAS_GELU : EPI_BIAS;
  g.aux = bias;
  return gemm_bf16(g, as_stream(stream));
}

int cm3p_conv1d_k3_wgrad(const void* dz, const void* x, float* dw, int batch, int c_in, int c_pad, int frames, int c_out,
                         int stride, int32_t* tile_sem, int64_t tile_sem_count, void* stream) {
  GemmArgs g;
  g.conv_mode = 2; g.conv_stride = stride; g.conv_batch = batch; g.conv_frames = frames; g.conv_cin = c_in;
  g.conv_cpad = c_pad;
  g.a = dz; g.b = x; g.c = dw;
  g.M = c_out;
  g.N = 1; g.K = 1;  // derived by the launcher
  g.epilogue = EPI_SCALE_F32;
  g.accumulate = 1;
  g.scale = 1.f;
  g.tile_sem = tile_sem; g.tile_sem_count = tile_sem_count;
  return gemm_bf16(g, as_stream(stream));
}

int cm3p_pool_project_normalize(const void* hidden_states, const int32_t* cu_seqlens, int mode, const void* proj_w,
                                void* pooled, float* proj_f32, float* inv_norm, float* embeds_f32, void* embeds_bf16,
                                int batch, int hidden, int proj_dim, void* stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  cudaStream_t s = as_stream(stream);
  if (mode == 0)
    rc = gather_rows(hidden_states, cu_seqlens, pooled, batch, hidden, s);
  else
    rc = mean_pool(hidden_states, cu_seqlens, pooled, batch, hidden, s);
  if (rc != kOk) return rc;
  if (!proj_w) return kOk;  // pooling only
  GemmArgs g;
  g.a = pooled; g.lda = hidden;
  g.b = proj_w; g.ldb = hidden;
  g.c = proj_f32; g.ldc = proj_dim;
  g.M = batch; g.N = proj_dim; g.K = hidden;
  g.epilogue = EPI_SCALE_F32;
  g.scale = 1.f;
  rc = gemm_bf16(g, s);
  if (rc != kOk) return rc;
  if (!embeds_f32 && !embeds_bf16) return kOk;
  return l2norm_rows(proj_f32, embeds_f32, embeds_bf16, inv_norm, batch, proj_dim, s);
}

int cm3p_clip_loss_fwd(const float* S, const int32_t* true_idx, float* row_lse, float* col_lse, float* loss, int Bm,
                       int V, int Bb, void* stream) {
  int rc = check_arch();
  if (rc != kOk) return rc;
  return clip_loss_fwd(S, true_idx, row_lse, col_lse, loss, Bm, V, Bb, as_stream(stream));
}

// ------------------------------------------------------------------------------------------ backward
int cm3p_attn_varlen_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                         const int32_t* cu_seqlens, const int32_t* positions, const float* rope_table,
                         int64_t total_tokens, int batch, int heads, int head_dim, int max_seqlen, int window,
                         const int32_t* groups, const int32_t* n_groups, int max_groups, void* stream) {
  AttnBwdArgs a;
  a.groups = groups; a.n_groups = n_groups; a.max_groups = max_groups;
  a.qkv = qkv; a.out = out; a.dout = dout; a.lse = lse; a.delta = delta; a.dqkv = dqkv;
  a.cu_seqlens = cu_seqlens; a.positions = positions; a.rope_table = rope_table;
  a.total_tokens = total_tokens; a.batch = batch; a.heads = heads; a.head_dim = head_dim;
  a.max_seqlen = max_seqlen; a.window = window;
  return attn_varlen_bwd(a, as_stream(stream));
}

#define CM3P_ARCH_GUARD()      \
  do {                         \
    int _rc = check_arch();    \
    if (_rc != kOk) return _rc; \
  } while (0)

int cm3p_layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* dres, void* dx, float* dgamma,
                       int64_t rows, int hidden, float eps, void* stream) {
  CM3P_ARCH_GUARD();
  return layernorm_bwd(x, dy, gamma, dres, dx, dgamma, rows, hidden, eps, as_stream(stream));
}

int cm3p_embed_gather_ln_bwd(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot,
                             const void* tok_emb, const void* audio_embeds, const float* gamma, const void* dy,
                             float* d_tok_emb, void* d_audio_embeds, float* dgamma, int64_t rows, int hidden, int vocab,
                             float eps, void* stream) {
  CM3P_ARCH_GUARD();
  return embed_gather_ln_bwd(ids, src_index, audio_slot, tok_emb, audio_embeds, gamma, dy, d_tok_emb, d_audio_embeds,
                             dgamma, rows, hidden, vocab, eps, as_stream(stream));
}

int cm3p_geglu_bwd(const void* ug, const void* dh, void* dug, void* h, int64_t rows, int intermediate, void* stream) {
  CM3P_ARCH_GUARD();
  return geglu_bwd(ug, dh, dug, h, rows, intermediate, as_stream(stream));
}

int cm3p_gelu_fwd(const void* z, void* y, int64_t n, void* stream) {
  CM3P_ARCH_GUARD();
  return gelu_fwd(z, y, n, as_stream(stream));
}

int cm3p_gelu_bwd(const void* z, const void* dy, void* dz, int64_t n, void* stream) {
  CM3P_ARCH_GUARD();
  return gelu_bwd(z, dy, dz, n, as_stream(stream));
}

int cm3p_colsum_f32(const void* dy, float* out, int64_t rows, int n, void* stream) {
  CM3P_ARCH_GUARD();
  return colsum_f32(dy, out, rows, n, as_stream(stream));
}

int cm3p_pool_bwd(const void* dpooled, const int32_t* cu_seqlens, void* dhidden, int mode, int accumulate, int batch,
                  int hidden, void* stream) {
  CM3P_ARCH_GUARD();
  return pool_bwd(dpooled, cu_seqlens, dhidden, mode, accumulate, batch, hidden, as_stream(stream));
}

int cm3p_l2norm_bwd(const float* proj_f32, const float* inv_norm, const float* dembeds, void* dproj_bf16, int rows,
                    int proj_dim, void* stream) {
  CM3P_ARCH_GUARD();
  return l2norm_bwd(proj_f32, inv_norm, dembeds, dproj_bf16, rows, proj_dim, as_stream(stream));
}

int cm3p_clip_loss_bwd(const float* S, const int32_t* true_idx, const float* row_lse, const float* col_lse,
                       const float* grad_out, void* dS, int64_t ld_ds, float* dlogit_scale, int Bm, int V, int Bb,
                       void* stream) {
  CM3P_ARCH_GUARD();
  return clip_loss_bwd(S, true_idx, row_lse, col_lse, grad_out, dS, ld_ds, dlogit_scale, Bm, V, Bb, as_stream(stream));
}

int cm3p_conv2_col2im_gelu_bwd(const void* da2, const void* z1, void* dz1, int batch, int frames, int channels,
                               void* stream) {
  CM3P_ARCH_GUARD();
  return conv2_col2im_gelu_bwd(da2, z1, dz1, batch, frames, channels, as_stream(stream));
}

int cm3p_vocab_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index,
                      int ignore_index, float* row_lse, float* loss_sum, float* count, int64_t rows, int vocab,
                      void* stream) {
  CM3P_ARCH_GUARD();
  return vocab_ce_fwd(logits, ld, labels, src_index, ignore_index, row_lse, loss_sum, count, rows, vocab,
                      as_stream(stream));
}

int cm3p_vocab_ce_bwd(void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index, int ignore_index,
                      const float* row_lse, const float* scale, int64_t rows, int vocab, void* stream) {
  CM3P_ARCH_GUARD();
  return vocab_ce_bwd(logits, ld, labels, src_index, ignore_index, row_lse, scale, rows, vocab, as_stream(stream));
}

int cm3p_gather_rows(const void* x, const int32_t* index, void* out, int64_t rows, int hidden, void* stream) {
  CM3P_ARCH_GUARD();
  return gather_rows_i32(x, index, out, rows, hidden, as_stream(stream));
}

int cm3p_scatter_add_rows(const void* dx_rows, const int32_t* index, void* dx, int64_t rows, int hidden,
                          void* stream) {
  CM3P_ARCH_GUARD();
  return scatter_add_rows(dx_rows, index, dx, rows, hidden, as_stream(stream));
}

int cm3p_segment_accumulate(const float* embeds, const int32_t* slot, float* sums, float* counts, int rows,
                            int proj_dim, void* stream) {
  CM3P_ARCH_GUARD();
  return segment_accumulate(embeds, slot, sums, counts, rows, proj_dim, as_stream(stream));
}
int cm3p_mean_renormalize(const float* sums, const float* counts, float* out, int rows, int proj_dim, void* stream) {
  CM3P_ARCH_GUARD();
  return mean_renormalize(sums, counts, out, rows, proj_dim, as_stream(stream));
}

// ------------------------------------------------------------------------------------------ log-mel
int cm3p_logmel_frames(const float* wave, const float* window, void* frames_hi, void* frames_lo, int batch,
                       int64_t samples, int frames, int n_fft, int hop, int ld, void* stream) {
  CM3P_ARCH_GUARD();
  return logmel_frames(wave, window, frames_hi, frames_lo, batch, samples, frames, n_fft, hop, ld, as_stream(stream));
}
int cm3p_logmel_power_mel(const float* spec, int64_t ld_spec, const float* mel_filters, float* out, float* clip_max,
                          int batch, int frames, int bins, int mels, void* stream) {
  CM3P_ARCH_GUARD();
  return logmel_power_mel(spec, ld_spec, mel_filters, out, clip_max, batch, frames, bins, mels, as_stream(stream));
}
int cm3p_logmel_finalize(float* out, const float* clip_max, int batch, int64_t per_clip, void* stream) {
  CM3P_ARCH_GUARD();
  return logmel_finalize(out, clip_max, batch, per_clip, as_stream(stream));
}

// ------------------------------------------------------------------------------------------ optimizer
int cm3p_muon_momentum(const float* grad, float* momentum_buffer, void* x_bf16, int64_t n, float momentum,
                       int nesterov, float* sumsq, void* stream) {
  CM3P_ARCH_GUARD();
  return muon_momentum(grad, momentum_buffer, x_bf16, n, momentum, nesterov, sumsq, as_stream(stream));
}
int cm3p_bf16_normalize(void* x_bf16, int64_t n, const float* sumsq, float eps, void* stream) {
  CM3P_ARCH_GUARD();
  return bf16_normalize(x_bf16, n, sumsq, eps, as_stream(stream));
}
int cm3p_bf16_axpy(void* out, int64_t ld_out, float a, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
                   int64_t rows, int64_t cols, void* stream) {
  CM3P_ARCH_GUARD();
  return bf16_axpy(out, ld_out, a, x, ld_x, y, ld_y, rows, cols, as_stream(stream));
}
int cm3p_muon_apply(float* param, const void* x_bf16, int64_t n, float post_scale, float alpha, void* stream) {
  CM3P_ARCH_GUARD();
  return muon_apply(param, x_bf16, n, post_scale, alpha, as_stream(stream));
}
int cm3p_adamw_step(float* param, const float* grad, float* moment1, float* moment2, int64_t n, float beta1,
                    float beta2, float eps, float decay, float step_size, void* stream) {
  CM3P_ARCH_GUARD();
  return adamw_step(param, grad, moment1, moment2, n, beta1, beta2, eps, decay, step_size, as_stream(stream));
}

// ------------------------------------------------------------------------------------------ embedding-table analysis
int cm3p_normalize_vectors(const float* x, float* out, int64_t n, int d, void* stream) {
  CM3P_ARCH_GUARD();
  return embed_normalize(x, out, n, d, as_stream(stream));
}
int64_t cm3p_pca2_workspace_floats(int64_t n, int d) { return embed_pca_workspace_floats(n, d); }
int cm3p_pca2(const float* x, int64_t n, int d, const float* init, int iterations, float* mean, float* components,
              float* proj, float* ws, int64_t ws_floats, void* stream) {
  CM3P_ARCH_GUARD();
  CM3P_REQUIRE(ws_floats >= embed_pca_workspace_floats(n, d), kBadShape, "pca2: workspace too small");
  return embed_pca2(x, n, d, init, iterations, mean, components, proj, ws, as_stream(stream));
}
int64_t cm3p_knn_workspace_bytes(int64_t n, int k) { return embed_knn_workspace_bytes(n, k); }
int cm3p_knn_cosine(const float* xn, int64_t n, int d, int64_t query, int k, int64_t* out_idx, float* out_dist,
                    void* ws, int64_t ws_bytes, void* stream) {
  CM3P_ARCH_GUARD();
  CM3P_REQUIRE(ws_bytes >= embed_knn_workspace_bytes(n, k), kBadShape, "knn: workspace too small");
  return embed_knn(xn, n, d, query, k, out_idx, out_dist, ws, as_stream(stream));
}
int64_t cm3p_kmeans_workspace_bytes(int64_t n, int d, int k) { return embed_kmeans_workspace_bytes(n, d, k); }
int cm3p_kmeans(const float* x, int64_t n, int d, int k, int64_t first_index, int iterations, float* centroids,
                int8_t* labels, int32_t* changed_per_iter, void* ws, int64_t ws_bytes, void* stream) {
  CM3P_ARCH_GUARD();
  CM3P_REQUIRE(ws_bytes >= embed_kmeans_workspace_bytes(n, d, k), kBadShape, "kmeans: workspace too small");
  return embed_kmeans(x, n, d, k, first_index, iterations, centroids, labels, changed_per_iter, ws, as_stream(stream));
}

}  // extern "C"
