/* cm3p_b200 — C ABI of the B200-native CM3P hot path (libcm3p_b200.so).
 *
 * The reference (OliBomby/CM3P) is pure Python/PyTorch and has no FFI of its own; every device op on
 * its hot path is a torch library call.  Each entry point below therefore cites the reference *call
 * site(s)* it replaces (paths relative to the reference checkout; "MB:" is the third-party
 * transformers/models/modernbert/modeling_modernbert.py that holds the encoder arithmetic).
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise
 *   - the caller owns all memory (inputs, outputs, workspaces); the library never allocates or frees
 *     device memory and keeps no state besides cached device properties
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns without
 *     synchronising; 0 = success, negative = error (see CM3P_ERR_*), message via cm3p_last_error()
 *   - bf16 activations / weights, fp32 statistics, LayerNorm weights, biases, logits and losses
 *   - sm_100a only: on any other device every compute entry returns CM3P_ERR_ARCH (no fallback)
 */
#ifndef CM3P_B200_H_
#define CM3P_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM3P_B200_VERSION 200 /* 0.2.0 */

#define CM3P_OK 0
#define CM3P_ERR_SHAPE (-1)
#define CM3P_ERR_ALIGN (-2)
#define CM3P_ERR_ARCH (-3)
#define CM3P_ERR_CUDA (-4)
#define CM3P_ERR_DRIVER (-5)

/* GEMM epilogues (cm3p_gemm_bf16.epilogue) */
#define CM3P_EPI_STORE 0      /* C = acc                                   bf16 */
#define CM3P_EPI_RESIDUAL 1   /* C = acc + aux[M,N]                        bf16, aux may alias C */
#define CM3P_EPI_GELU 2       /* C = gelu_erf(acc) */
#define CM3P_EPI_BIAS_GELU 3  /* C = gelu_erf(acc + bias[N]), bias fp32 in aux */
#define CM3P_EPI_BIAS 4       /* C = acc + bias[N] */
#define CM3P_EPI_GEGLU 5      /* C[M,N/2] = gelu_erf(u)*g; B rows interleaved in groups of 16 (u,g) */
#define CM3P_EPI_GEGLU_SAVE 6 /* as GEGLU, and C2[M,N] = acc (pre-activation kept for backward) */
#define CM3P_EPI_ROPE 7       /* rotate-half RoPE on columns [0, rope_cols) per 64-wide head */
#define CM3P_EPI_SCALE_F32 8  /* C(fp32) = scale*acc (aux != NULL: scale *= exp(*aux), aux a device fp32 scalar such as
                                 logit_scale); accumulate != 0: C += scale*acc with fp32 atomics, K split
                                 across CTAs (weight gradients: K = number of tokens); with tile_sem the splits of a
                                 tile add in split order instead (bit-reproducible) */

const char* cm3p_last_error(void);
int cm3p_version(void);
int cm3p_num_sms(void); /* host query; 0 if no CUDA device */

/* Run-time knobs (process-wide, thread-safe).  The defaults are the product; the tests sweep the streaming depths
 * and the benchmarks use the others for A/B measurements.  Nothing is read from the environment. */
#define CM3P_OPT_FWD_BLOCKS_PER_CTA 0      /* attention forward: 256-query blocks streamed per CTA; 0 = heuristic */
#define CM3P_OPT_BWD_OUTER_PER_CTA 1       /* attention backward: outer tiles streamed per CTA; 0 = heuristic */
#define CM3P_OPT_GEMM_CLUSTER 2            /* 2 = CTA pairs run one tcgen05.mma.cta_group::2 of M = 256 (default), 3 = CTA pairs
                                              share B tiles through TMA multicast (each its own MMA), 1 = single CTAs */
#define CM3P_OPT_ATTN_FORCE_TILE_KERNELS 3 /* 1 = one-tile-per-CTA attention kernels for every sequence length */
#define CM3P_OPT_WGRAD_DETERMINISTIC 4     /* 1 = ordered split-K accumulation (bit-reproducible weight gradients; the
                                              split-K GEMMs get ~1.7x slower), 0 = fp32 atomics (default) */
#define CM3P_OPT_TMAP_CACHE 5              /* 1 = cache encoded CUtensorMaps by (pointer, shape, pitch, box) (default) */
#define CM3P_OPT_ATTN_WINDOW_WALK 6         /* 1 = fused band-walk backward for sliding-window layers (default) */
int cm3p_set_option(int option, int value);
int cm3p_get_option(int option);

/* C[M,N] = epilogue(A . B^T) on tcgen05 tensor cores, fp32 accumulation in TMEM.
 * Replaces nn.Linear everywhere on the path: MB:74-91 (Wi/Wo), MB:232-310 (Wqkv/Wo),
 * cm3p/modeling_cm3p.py:470-481 (projector), :959/:971 (projections), :976-977 (logits matmul and
 * logit-scale multiply), :1229-1238 (MLM head); with CM3P_EPI_ROPE also MB:197-228.
 *   a: trans_a == 0 -> [M,K] K-contiguous, else [K,M] M-contiguous; lda = row pitch in elements
 *   b: trans_b == 0 -> [N,K] K-contiguous (nn.Linear.weight layout), else [K,N]
 *   positions [M] int32 + rope_table [max_pos][32][2] fp32 (cos,sin) only for CM3P_EPI_ROPE
 *   tile_sem: tile_sem_count zeroed int32 counters (left zeroed) for ordered split-K accumulation, or NULL
 *   group_m > 0: grouped GEMM, M / group_m independent (group_m x N x K) problems in one launch whose operands are
 *     stacked along their outer dimension (A [M,K] or, transposed, [groups*K, group_m]; B [groups*N, K] or
 *     [groups*K, N]; C / aux [M, N]); group_m % 256 == 0 and K % 64 == 0.  Used for the Newton-Schulz iterations of
 *     Muon over all same-shaped weight matrices at once (utils/muon_utils.py:35-57). */
int cm3p_gemm_bf16(const void* a, int64_t lda, int trans_a, const void* b, int64_t ldb, int trans_b, void* c,
                   int64_t ldc, int64_t M, int64_t N, int64_t K, int epilogue, const void* aux, int64_t ld_aux,
                   void* c2, int64_t ldc2, float scale, int accumulate, const int32_t* positions,
                   const float* rope_table, int64_t rope_cols, int32_t* tile_sem, int64_t tile_sem_count,
                   int64_t group_m, void* stream);

/* The same GEMM with a LayerNorm folded into the GEMMs on both sides of it, so that the pre-norm blocks of
 * ModernBERT (MB:313-342: `attn(attn_norm(x))`, `mlp(mlp_norm(x))`) need no LayerNorm pass at all:
 *   producer, epilogue CM3P_EPI_RESIDUAL: stats_out [ceil(N/256)][M][2] fp32 = (sum, sum of squares) of the bf16
 *     values each 256-column tile wrote to a row (one partial per tile: no atomics, no zero-fill, results do not
 *     depend on the order in which CTAs finish)
 *   consumer, epilogue CM3P_EPI_ROPE / CM3P_EPI_GEGLU(_SAVE), with b = W . diag(gamma) (bf16) and
 *     col_corr [N] = row sums of b:   acc <- rstd_m * (acc - mean_m * col_corr[n])
 *     where mean / rstd come from row_stats [ceil(K/256)][M][2] (the producer's stats_out, partials summed in
 *     tile order; width = K, eps = ln_eps).
 * Operands K-major only; outputs must be 16-byte aligned bf16 rows. */
int cm3p_gemm_bf16_ln(const void* a, int64_t lda, const void* b, int64_t ldb, void* c, int64_t ldc, int64_t M, int64_t N,
                      int64_t K, int epilogue, const void* aux, int64_t ld_aux, void* c2, int64_t ldc2,
                      const int32_t* positions, const float* rope_table, int64_t rope_cols, float* stats_out,
                      const float* row_stats, const float* col_corr, float ln_eps, void* stream);

/* Unpadded bidirectional attention, head_dim 64: out = softmax(q k^T / 8 | mask) v per sequence/head.
 * Replaces ALL_ATTENTION_FUNCTIONS[...] in MB:286-300 (sdpa / flash_attention_2 / eager) together
 * with the padding + sliding-window masks (transformers/masking_utils.py:121-131) and the
 * unpad/repad helpers cm3p/modeling_cm3p.py:65-134.
 *   qkv [T,3,heads,64] bf16 (q,k already rotated), out [T,heads*64] bf16, lse [heads,T] fp32 or NULL
 *   window < 0: global layer; window = w: attend iff |i-j| <= w (ModernBERT: local_attention/2)
 *   groups / n_groups / max_groups (all or none): packed short sequences.  When every sequence has <= 128 tokens
 *     (the metadata tower: B*V sequences of ~21 tokens, cm3p/modeling_cm3p.py:351-380, tokenization_cm3p.py:632-654)
 *     consecutive sequences are packed into groups of <= 128 tokens by cm3p_attn_pack_groups and one CTA serves a
 *     whole group with a block-diagonal mask, instead of one 128-row tile per sequence. */
int cm3p_attn_varlen_fwd(const void* qkv, void* out, float* lse, const int32_t* cu_seqlens, int64_t total_tokens,
                         int batch, int heads, int head_dim, int max_seqlen, int window, const int32_t* groups,
                         const int32_t* n_groups, int max_groups, void* stream);

/* Group table for the packed attention kernels: groups [max_groups][2] int32 = (first sequence, end sequence) of
 * consecutive sequences holding <= 128 tokens together, in no particular order; *n_groups (device) = their number.
 * max_groups >= min(batch, 2 * (total_tokens / 128) + (batch + 63) / 64 + 1).  Integer work only. */
int cm3p_attn_pack_groups(const int32_t* cu_seqlens, int batch, int32_t* groups, int32_t* n_groups, int max_groups,
                          void* stream);

/* y = LayerNorm(x) * gamma (no bias), rows of H bf16; stats [rows][2] = (mean, rstd) or NULL.
 * Replaces nn.LayerNorm at MB:63 (embeddings.norm), MB:318-323 (attn_norm / mlp_norm), final_norm. */
int cm3p_layernorm_fwd(const void* x, const float* gamma, void* y, float* stats, int64_t rows, int hidden, float eps,
                       void* stream);

/* x0[t] = LayerNorm(audio_slot[t] >= 0 ? audio_embeds[audio_slot[t]] : tok_emb[ids[src_index[t]]]).
 * Replaces cm3p/modeling_cm3p.py:591-592 (embedding gather), :603-605 (audio scatter) and MB:63.
 *   ids: padded int64 [B*L]; src_index [T] int32 flat index of every real token (NULL = identity);
 *   audio_slot [T] int32 running count of [AUDIO] tokens, -1 elsewhere (NULL = no audio) */
int cm3p_embed_gather_ln(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot, const void* tok_emb,
                         const void* audio_embeds, const float* gamma, void* y, float* stats, int64_t rows, int hidden,
                         int vocab, float eps, void* stream);

/* (gelu of) conv1d(x, kernel 3, padding 1, stride 1 | 2) + bias as an IMPLICIT GEMM on the tensor cores: the
 * im2col matrix is never formed.  For every 64-channel K block of a tap the producer fetches a {64 channels x 128
 * frames} box of the channels-last input through a 4-D tensor map (channel, frame parity, frame / stride, window),
 * shifted by the tap; the zero padding at the window edges is the tensor map's out-of-bounds fill, stride 2 is the
 * parity dimension; the result is written channels-last [B, F/stride, C_out] (no permute pass).
 * Replaces cm3p/modeling_cm3p.py:488-489, :501-504 (conv1 + gelu, conv2 + gelu, permute(0,2,1).contiguous()).
 *   x: bf16 [batch, frames, c_in] channels-last (the log-mel input goes through cm3p_transpose_cast_bf16 once)
 *   weight: bf16 [c_out, 3 * c_pad], column tap * c_pad + c = conv.weight[:, c, tap], zero for c >= c_in;
 *           c_pad = c_in rounded up to a multiple of 64;  bias fp32 [c_out];  gelu != 0: exact-erf GELU fused
 *   out: bf16 [batch, frames / stride, c_out] */
int cm3p_conv1d_k3_fwd(const void* x, const void* weight, const float* bias, void* out, int batch, int c_in, int c_pad,
                       int frames, int c_out, int stride, int gelu, void* stream);

/* Weight gradient of the same convolution, again without an im2col matrix: dw [c_out, 3 * c_pad] fp32 +=
 * sum over (window, frame) of dz[b, t, :]^T x[b, stride * t + tap - 1, :]  (K = frames, split across CTAs).
 *   dz: bf16 [batch, frames / stride, c_out];  x as in the forward;  tile_sem as in cm3p_gemm_bf16 (may be NULL) */
int cm3p_conv1d_k3_wgrad(const void* dz, const void* x, float* dw, int batch, int c_in, int c_pad, int frames, int c_out,
                         int stride, int32_t* tile_sem, int64_t tile_sem_count, void* stream);

/* [batch, channels, frames] fp32 (the processor's log-mel layout, cm3p/processing_cm3p.py:284-304) ->
 * [batch, frames, channels] bf16. */
int cm3p_transpose_cast_bf16(const float* x, void* out, int batch, int channels, int frames, void* stream);

/* pooled = first token (mode 0) or masked mean (mode 1) of every sequence; e = pooled . W^T;
 * embeds = e / sqrt(sum e^2) (no epsilon).  Replaces cm3p/modeling_cm3p.py:624-642 / :382-396
 * (pooling), :959-960 / :971-972 (projection + _get_vector_norm :54-62).
 *   pooled [B,H] bf16 (output), proj_f32 [B,P] fp32 (output, un-normalised), inv_norm [B] or NULL,
 *   embeds_f32 [B,P] / embeds_bf16 [B,P] outputs (either may be NULL) */
int cm3p_pool_project_normalize(const void* hidden_states, const int32_t* cu_seqlens, int mode, const void* proj_w,
                                void* pooled, float* proj_f32, float* inv_norm, float* embeds_f32, void* embeds_bf16,
                                int batch, int hidden, int proj_dim, void* stream);

/* CLIP-style symmetric cross-entropy on S = logits_per_metadata [Bm*V, Bb] fp32 (already scaled).
 * Replaces cm3p_loss / contrastive_loss, cm3p/modeling_cm3p.py:27-51.
 *   true_idx [Bm] int32 = argmax(classes == 0) (all zeros for the 2-D case)
 *   row_lse [Bm], col_lse [Bb] fp32 outputs (kept for backward); loss: 1 fp32 */
int cm3p_clip_loss_fwd(const float* S, const int32_t* true_idx, float* row_lse, float* col_lse, float* loss, int Bm,
                       int V, int Bb, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Backward entry points.  The reference has no backward code of its own: these are the gradients
 * torch autograd derives for the forward call sites cited above (loss.backward() in
 * transformers.Trainer.training_step, reached from train.py:360-375).  Weight gradients are
 * cm3p_gemm_bf16 calls on transposed operands (trans_a/trans_b, CM3P_EPI_SCALE_F32, accumulate=1).
 */

/* Gradient of cm3p_attn_varlen_fwd w.r.t. the un-rotated Wqkv output: dqkv [T,3,heads,64] bf16.
 *   out/dout [T,heads*64] bf16; lse [heads,T] from the forward; delta [heads,T] fp32 workspace;
 *   positions + rope_table (both or neither): dq/dk are rotated back (inverse of CM3P_EPI_ROPE).
 *   groups / n_groups / max_groups: packed short sequences as in the forward (one kernel, 5 GEMMs per group). */
int cm3p_attn_varlen_bwd(const void* qkv, const void* out, const void* dout, const float* lse, float* delta, void* dqkv,
                         const int32_t* cu_seqlens, const int32_t* positions, const float* rope_table,
                         int64_t total_tokens, int batch, int heads, int head_dim, int max_seqlen, int window,
                         const int32_t* groups, const int32_t* n_groups, int max_groups, void* stream);

/* dx = LayerNorm'(x; gamma) . dy (+ dres, the gradient arriving through the residual connection);
 * dgamma[H] fp32 += sum_rows dy * xhat (NULL to skip).  Statistics are recomputed from x. */
int cm3p_layernorm_bwd(const void* x, const void* dy, const float* gamma, const void* dres, void* dx, float* dgamma,
                       int64_t rows, int hidden, float eps, void* stream);

/* Backward of cm3p_embed_gather_ln: token rows are atomically added into d_tok_emb [vocab,H] fp32,
 * audio rows written to d_audio_embeds [n_audio,H] bf16 (either may be NULL); dgamma as above. */
int cm3p_embed_gather_ln_bwd(const int64_t* ids, const int32_t* src_index, const int32_t* audio_slot,
                             const void* tok_emb, const void* audio_embeds, const float* gamma, const void* dy,
                             float* d_tok_emb, void* d_audio_embeds, float* dgamma, int64_t rows, int hidden, int vocab,
                             float eps, void* stream);

/* GeGLU backward on the interleaved pre-activation ug [rows,2I] kept by CM3P_EPI_GEGLU_SAVE:
 * dug [rows,2I] (same interleaving), h [rows,I] = gelu(u)*g recomputed (NULL to skip). */
int cm3p_geglu_bwd(const void* ug, const void* dh, void* dug, void* h, int64_t rows, int intermediate, void* stream);

/* y = gelu_erf(z) and dz = dy * gelu_erf'(z) over n bf16 elements (n % 8 == 0). */
int cm3p_gelu_fwd(const void* z, void* y, int64_t n, void* stream);
int cm3p_gelu_bwd(const void* z, const void* dy, void* dz, int64_t n, void* stream);

/* out[n] fp32 += sum_rows dy[rows,n] bf16 (bias gradients). */
int cm3p_colsum_f32(const void* dy, float* out, int64_t rows, int n, void* stream);

/* Backward of the pooling in cm3p_pool_project_normalize: dhidden [T,H] bf16 (overwritten, or added to
 * when accumulate != 0) from dpooled [B,H] bf16; mode as in the forward. */
int cm3p_pool_bwd(const void* dpooled, const int32_t* cu_seqlens, void* dhidden, int mode, int accumulate, int batch,
                  int hidden, void* stream);

/* Backward of embeds = e / |e|: dproj (bf16 [rows,P]) from proj_f32, inv_norm and dembeds (fp32). */
int cm3p_l2norm_bwd(const float* proj_f32, const float* inv_norm, const float* dembeds, void* dproj_bf16, int rows,
                    int proj_dim, void* stream);

/* Backward of cm3p_clip_loss_fwd: dS [Bm*V, ld_ds] bf16 (ld_ds >= Bb) and dlogit_scale (fp32, += sum dS*S).
 * grad_out: device scalar (upstream gradient of the loss) or NULL for 1. */
int cm3p_clip_loss_bwd(const float* S, const int32_t* true_idx, const float* row_lse, const float* col_lse,
                       const float* grad_out, void* dS, int64_t ld_ds, float* dlogit_scale, int Bm, int V, int Bb,
                       void* stream);

/* conv2 input gradient: dA2 [B*F/2, 3*C] = dz2 . W2 (a plain cm3p_gemm_bf16), scattered back onto [B,F,C] (col2im)
 * and fused with conv1's GELU backward:
 * dz1 = col2im(dA2) * gelu'(z1).  Replaces autograd through cm3p/modeling_cm3p.py:501-502. */
int cm3p_conv2_col2im_gelu_bwd(const void* da2, const void* z1, void* dz1, int batch, int frames, int channels,
                               void* stream);

/* Cross-entropy over a vocabulary / class set on bf16 logits [rows, ld] (ld >= vocab).  Replaces
 * `self.loss_function(logits, labels, vocab_size)` = transformers ForMaskedLMLoss (cm3p/modeling_cm3p.py:994-996,
 * :1365-1367; ignore_index -100) and the classifier CrossEntropyLoss (:1207-1209).
 *   target(row) = labels[src_index ? src_index[row] : row]  (labels: padded int64; src_index as in cm3p_embed_gather_ln)
 *   fwd: row_lse [rows]; loss_sum (fp32, +=) = sum over non-ignored rows of (lse - x[target]); count (fp32, +=)
 *   bwd: logits are overwritten by (softmax - onehot) * (*scale) (device scalar), zeros for ignored rows / pad columns */
int cm3p_vocab_ce_fwd(const void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index,
                      int ignore_index, float* row_lse, float* loss_sum, float* count, int64_t rows, int vocab,
                      void* stream);
int cm3p_vocab_ce_bwd(void* logits, int64_t ld, const int64_t* labels, const int32_t* src_index, int ignore_index,
                      const float* row_lse, const float* scale, int64_t rows, int vocab, void* stream);

/* out[r] = x[index[r]];  dx[index[r]] += dx_rows[r] (unique indices).  Sparse MLM prediction,
 * cm3p/modeling_cm3p.py:1349-1357 (`last_hidden_state[mask_tokens]`) and its gradient. */
int cm3p_gather_rows(const void* x, const int32_t* index, void* out, int64_t rows, int hidden, void* stream);
int cm3p_scatter_add_rows(const void* dx_rows, const int32_t* index, void* dx, int64_t rows, int hidden,
                          void* stream);

/* Per-beatmap aggregation of window embeddings (extract_beatmap_embeddings.py:243-262):
 *   cm3p_segment_accumulate: sums[slot[i]] += embeds[i] (fp32 [rows,P]), counts[slot[i]] += 1; slot < 0 skipped
 *   cm3p_mean_renormalize  : out[b] = mean_b / |mean_b|  (mean_b = sums[b]/counts[b]; left un-normalised if 0) */
int cm3p_segment_accumulate(const float* embeds, const int32_t* slot, float* sums, float* counts, int rows,
                            int proj_dim, void* stream);
int cm3p_mean_renormalize(const float* sums, const float* counts, float* out, int rows, int proj_dim, void* stream);

/* Log-mel front-end on the GPU: replaces `WhisperFeatureExtractor.__call__` as used by the reference's
 * processor (cm3p/processing_cm3p.py:284-304; n_fft 400, hop 160, 80 slaney mel bins).  The DFT between the two
 * kernels is three cm3p_gemm_bf16 calls on hi/lo-split bf16 operands (CM3P_EPI_SCALE_F32, accumulate).
 *   cm3p_logmel_frames   : reflect-centred, Hann-windowed frames [batch*frames, ld] as bf16 hi + lo parts
 *   cm3p_logmel_power_mel: spec [batch*frames, ld_spec] fp32 = (re[0..bins) | im[0..bins)) -> out [batch, mels, frames]
 *                          = log10(max(mel(|.|^2), 1e-10)); clip_max[batch] = running max (init to -inf)
 *   cm3p_logmel_finalize : out = (max(out, clip_max - 8) + 4) / 4 */
int cm3p_logmel_frames(const float* wave, const float* window, void* frames_hi, void* frames_lo, int batch,
                       int64_t samples, int frames, int n_fft, int hop, int ld, void* stream);
int cm3p_logmel_power_mel(const float* spec, int64_t ld_spec, const float* mel_filters, float* out, float* clip_max,
                          int batch, int frames, int bins, int mels, void* stream);
int cm3p_logmel_finalize(float* out, const float* clip_max, int batch, int64_t per_clip, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Analysis of the embedding table [n, d] fp32 (row-major): what the reference's browser visualizer computes in its
 * Rust -> WASM core over the parquet written by extract_beatmap_embeddings.py (visualizer/wasm/src/lib.rs).  All
 * HBM-bound streaming kernels with fixed-order reductions (bit-reproducible); ties go to the lower index.
 *   cm3p_normalize_vectors : out[i] = x[i] / |x[i]|, zero rows stay zero                       (lib.rs:371-432)
 *   cm3p_knn_cosine        : the k rows nearest to row `query` of a NORMALISED table by 1 - <x_i, x_q>, the query
 *                            excluded, ascending; k <= min(n - 1, 1024) and ceil(n / 4096) * k <= 4096  (lib.rs:448-488)
 *   cm3p_pca2              : mean[d]; components[2][d] by `iterations` (reference: 8) power iterations of X_c^T X_c
 *                            from the start vectors `init[2][d]` (the reference draws them from Math.random / an LCG),
 *                            second one orthogonalised against the first at the end; proj[n][2]       (lib.rs:82-237)
 *   cm3p_kmeans            : farthest-point seeding from row `first_index` (reference: LCG(seed) %% n), `iterations`
 *                            (reference: up to 10) Lloyd steps, first-minimum assignment, empty clusters keep their
 *                            centroid; labels int8 [n], centroids [k][d], changed_per_iter[iterations] = labels that
 *                            changed in each step (a converged run stops changing: same labels as the reference's early
 *                            stop); k <= 127                                                          (lib.rs:242-365)
 * Workspaces are caller-owned; the *_workspace_* queries give their sizes. */
int cm3p_normalize_vectors(const float* x, float* out, int64_t n, int d, void* stream);
int64_t cm3p_knn_workspace_bytes(int64_t n, int k);
int cm3p_knn_cosine(const float* xn, int64_t n, int d, int64_t query, int k, int64_t* out_idx, float* out_dist,
                    void* ws, int64_t ws_bytes, void* stream);
int64_t cm3p_pca2_workspace_floats(int64_t n, int d);
int cm3p_pca2(const float* x, int64_t n, int d, const float* init, int iterations, float* mean, float* components,
              float* proj, float* ws, int64_t ws_floats, void* stream);
int64_t cm3p_kmeans_workspace_bytes(int64_t n, int d, int k);
int cm3p_kmeans(const float* x, int64_t n, int d, int k, int64_t first_index, int iterations, float* centroids,
                int8_t* labels, int32_t* changed_per_iter, void* ws, int64_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizer step of the reference's Muon (utils/muon_utils.py).  The three GEMMs of every Newton-Schulz
 * iteration (:50-53) are cm3p_gemm_bf16 calls; these are the element-wise pieces, with the reference's
 * bf16 roundings.
 *   cm3p_muon_momentum : buf = buf*momentum + g; x = bf16(nesterov ? g + momentum*buf : buf); sumsq += |x|^2   (:152-156, :47)
 *   cm3p_bf16_normalize: x /= (bf16(sqrt(sumsq)) + eps)                                                     (:48)
 *   cm3p_bf16_axpy     : out = bf16(bf16(a*x) + y) over [rows, cols] with pitches (y may be NULL)          (:51-53)
 *   cm3p_muon_apply    : param += alpha * bf16(x * post_scale)                                             (:164-167)
 *   cm3p_adamw_step    : the internal AdamW for embeddings / vectors (:179-203):
 *                        m1 = lerp(m1,g,1-b1); m2 = lerp(m2,g^2,1-b2); p = p*decay - step_size * m1/(eps+sqrt(m2)) */
int cm3p_muon_momentum(const float* grad, float* momentum_buffer, void* x_bf16, int64_t n, float momentum,
                       int nesterov, float* sumsq, void* stream);
int cm3p_bf16_normalize(void* x_bf16, int64_t n, const float* sumsq, float eps, void* stream);
int cm3p_bf16_axpy(void* out, int64_t ld_out, float a, const void* x, int64_t ld_x, const void* y, int64_t ld_y,
                   int64_t rows, int64_t cols, void* stream);
int cm3p_muon_apply(float* param, const void* x_bf16, int64_t n, float post_scale, float alpha, void* stream);
int cm3p_adamw_step(float* param, const float* grad, float* moment1, float* moment2, int64_t n, float beta1,
                    float beta2, float eps, float decay, float step_size, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CM3P_B200_H_ */
