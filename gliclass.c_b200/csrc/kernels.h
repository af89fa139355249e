// Launchers of the hand-written sm_100a kernels (K1..K5 of SURVEY.md §2.3).  All launch on
// `stream` and return the launch status; none synchronises.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace glc {

// Storage convention: activations and weights are fp16, every accumulation / statistic is fp32.
// fp16 rather than bf16 because bf16's 8-bit mantissa cannot meet the 2e-2 logit parity bar on
// this model family (DESIGN.md "Numerics").
//
// K2: C[M,N] = act(A[M,K] W[N,K]^T + bias); A, W fp16, fp32 accumulate; C fp16 or fp32.  act: 0 none, 1 erf-GELU, 2 ReLU,
// 3 SwiGLU (W rows interleaved in blocks of 32: gate rows, then the matching up rows; C is [M, N/2] = silu(gate) * up).
cudaError_t gemm_f16(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M,
                     int N, int K, int act, bool out_f32, int num_sms, cudaStream_t stream);

// K2 with the residual add fused into the epilogue: C = act(A W^T + bias) + resid (fp16 [M, ldr]); resid may be null.
cudaError_t gemm_f16_resid(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* resid,
                           int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int act, bool out_f32, int num_sms,
                           cudaStream_t stream);

// K2 with the rotary embedding fused into the epilogue (QKV projection of the decoder stack, head dim 128): columns
// [0, rope_cols) — the q and kv heads, 128 columns each — are rotated in pairs (p, p + 64) by the angle of the row's position
// (rope_tile_pos[row / 128] + row % 128 in the packed layout, else row % rope_S); rope_cs from rope_table; N % 128 == 0.
cudaError_t gemm_f16_rope(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N,
                          int K, const void* rope_cs_f32x2, int rope_S, const int32_t* rope_tile_pos, int rope_cols, int num_sms,
                          cudaStream_t stream);

// K2 on e4m3 operands (the opt-in FP8 FFN path): C = act((A8 W8^T) * a_scale[m] * a_const * w_scale[n] + bias).
// A8 [M,K], W8 [N,K] e4m3 bytes, K % 16 == 0; a_scale may be null (then a_const alone); C fp16 [M,ldc], or with
// out_e4m3 the saturating e4m3 of (act(...) * out_mult).  act: 0, or 1 (erf-GELU, e4m3 output only).
cudaError_t gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, const float* a_scale, float a_const,
                      const float* w_scale, const float* bias, void* C, int64_t ldc, int M, int N, int K, int act, bool out_e4m3,
                      float out_mult, int num_sms, cudaStream_t stream);
// per-row e4m3 quantiser: q[r, :] = e4m3(x[r, :] * 448 / amax_r), scale[r] = amax_r / 448 (1 for an all-zero row)
cudaError_t quantize_rows_e4m3(const void* x_f16, int64_t ldx, void* q8, int64_t ldq, float* scale, int M, int K, cudaStream_t stream);

// decoder backbone (decoder.cu): fp32 residual stream h, fp16 normalised activations
cudaError_t embed_rows_f32(const int64_t* ids, const void* emb_f16, float* h, int M, int H, int vocab, cudaStream_t stream);
// h += delta (fp16 [M,H], may be null); y = rmsnorm(h) * g
cudaError_t add_rmsnorm(float* h, const void* delta_f16, const float* g, float eps, void* y_f16, int M, int H, cudaStream_t stream,
                        void* y_e4m3 = nullptr, float* y_e4m3_scale = nullptr);   // optional: the row also as e4m3 + per-row scale
// (cos, sin) float2 table [S, head_dim / 2] from the rotary inverse frequencies
cudaError_t rope_table(const float* inv_freq, void* cs_f32x2, int S, int head_dim, cudaStream_t stream);
// rotary embedding in place on the first n_rot_heads heads (q heads, then kv heads) of qkv fp16 [M, ld]
cudaError_t rope_inplace(void* qkv_f16, int64_t ld, const void* cs_f32x2, int M, int S, int n_rot_heads, int head_dim,
                         cudaStream_t stream, const int32_t* tile_pos = nullptr);
// (tile_pos != nullptr: packed layout, position of row m = tile_pos[m / 128] + m % 128 instead of m % S)
// plain (bias-free) flash attention for head dim 128 with grouped-query heads, bidirectional, key padding mask:
// qkv fp16 [B*S, (heads + 2 kv_heads) * 128] (Q heads | K heads | V heads), ctx fp16 [B*S, heads * 128]
cudaError_t attention_flash128(const void* qkv_f16, const uint32_t* mask_bits, const int32_t* kv_len, void* ctx_f16, int B, int S,
                               int heads, int kv_heads, cudaStream_t stream);
// the same kernel on the packed (varlen) layout (conventions of attention_persist_packed)
cudaError_t attention_flash128_packed(const void* qkv, const uint32_t* row_bits, const int32_t* kv_len, const int32_t* text_row,
                                      const int32_t* tile_info, void* ctx, int B, int rows, int n_tiles, int heads, int kv_heads,
                                      cudaStream_t stream);

// K1: y[m,:] = LN(word_emb[ids[m],:]) * gamma + beta, times mask[m]   (T:520-564)
cudaError_t embed_ln(const int64_t* ids, const int64_t* mask, const void* word_emb_f16, const float* gamma,
                     const float* beta, float eps, void* y_f16, int M, int H, int vocab, cudaStream_t stream);

// K4: y = LN(x + r) * gamma + beta   (T:49-53, T:408-412); r may be null (plain LN).
// overflow_flag (device int, may be null) is set to 1 when a pre-LN sum reaches the fp16 saturation value (the GEMM
// epilogue clamps at +-65504): the engine turns that into a loud Run failure.  x_is_f32: x is the fp32 output of the
// GEMM's fp32 epilogue (robust mode, GLC_PRELN_F32=1) instead of fp16.
cudaError_t residual_ln(const void* x, const void* r_f16, const float* gamma, const float* beta, float eps,
                        void* y_f16, int M, int H, cudaStream_t stream, int* overflow_flag = nullptr, bool x_is_f32 = false,
                        void* y_e4m3 = nullptr, float* y_e4m3_scale = nullptr);
// (y_e4m3 / y_e4m3_scale: also write each row as e4m3 under its dynamic scale amax/448 — the FP8 FFN1 operand)
// plain LN on fp32 rows -> fp16 (load-time LN of rel_embeddings, T:597-601)
cudaError_t ln_f32_to_f16(const float* x, const float* gamma, const float* beta, float eps, void* y_f16, int M, int H,
                           cudaStream_t stream);

// key-validity words for attention: bits[b][w] bit j = mask[b][32w+j] != 0 ; kv_len[b] = 1 + last valid key
cudaError_t mask_prep(const int64_t* mask, uint32_t* bits, int32_t* kv_len, int B, int S, cudaStream_t stream);

// K3: fused disentangled attention (T:229-345).  qkv fp16 [B*S,3H] (Q | K | V, head-major inside each third);
// ctx fp16 [B*S,H].  The kernels read the per-layer position projections expanded at load to one row per relative
// distance (expand_pos_table): exp_k [expanded_pos_rows()][ld_exp], row rho = posK[idx(2047 - rho)] (index from
// expanded_pos_index), exp_qr in the OPPOSITE order, row sigma = posQ[idx(sigma - 2047)] (expanded_pos_index_rev).
int expanded_pos_rows();
void expanded_pos_index(int buckets, int max_pos, int32_t* out /* host, [expanded_pos_rows()] */);
void expanded_pos_index_rev(int buckets, int max_pos, int32_t* out /* host, [expanded_pos_rows()] */);
cudaError_t expand_pos_table(const void* pos_f16, int64_t ld_src, const int32_t* d_exp_index, void* out_f16, int64_t ld_dst,
                             int cols, cudaStream_t stream);
// production kernel (attention_persist.cu): both biases un-skewed in registers, a softmax thread owns a whole query row of
// a 64-key tile, three warpgroups rotate over the key tiles; one CTA per SM for the whole launch, work items dealt out in
// contiguous chunks ordered (head, query tile, batch row); for S <= 512 the expanded table windows of the CTA's
// (head, query tile) stay resident in shared memory, so a tile costs 16 KB of L2 traffic instead of 48 KB.  (The two
// generations it supersedes are experiments/attention_generations/attention_{shift,rows}.cu.)
cudaError_t attention_persist(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                              const uint32_t* mask_bits, const int32_t* kv_len, void* ctx, int B, int S, int heads,
                              int num_sms, cudaStream_t stream);
// The same kernel on the PACKED (varlen) layout: text b owns rows [text_row[b], text_row[b+1]) of qkv / ctx (multiples of
// 128), row_bits holds one validity bit per packed row, tile_info[t] = (query tile index within its text) << 24 | text for
// the rows / 128 query tiles ordered (query tile index, text); max_text_rows = the longest text's row count.
cudaError_t attention_persist_packed(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp,
                                     const uint32_t* row_bits, const int32_t* kv_len, const int32_t* text_row,
                                     const int32_t* tile_info, void* ctx, int B, int rows, int max_text_rows, int n_tiles,
                                     int heads, int num_sms, cudaStream_t stream);
// slow CUDA-core restatement of the same op on the UNEXPANDED tables (pos_k / pos_q fp16 [2*buckets][ld_pos], rel_idx
// int32 [2*Spad-1] with Spad = S rounded up to 128), used only by tests to localise bugs on the GPU
cudaError_t attention_naive(const void* qkv, const void* pos_k, const void* pos_q, int64_t ld_pos, const int32_t* rel_idx,
                            const uint32_t* mask_bits, void* ctx, int B, int S, int heads, int buckets,
                            cudaStream_t stream);

// context of padded QUERY rows as the traced graph computes it (uniform softmax over all S keys = mean of V);
// only needed when something reads padded positions (pooling_strategy='last' on right-padded batches)
cudaError_t pad_rows_mean_v(const void* qkv_f16, const int64_t* mask, void* ctx_f16, int B, int S, int H, cudaStream_t stream);

// K5a: <<LABEL>>-token pooling.  pooled[b,:] = h[b,0,:]; cls[b,c,:] = h[b,pos_c(b),:] or 0  
cudaError_t head_gather(const void* h_f16, const int64_t* ids, int64_t class_token, void* pooled_f16, void* cls_f16,
                        int B, int S, int H, int C, cudaStream_t stream);
// K5a with the pooling strategies of the gliclass package: pool_mode 0 first token, 1 last token (h[b,S-1,:]),
// 2 masked mean, 3 masked max over the sequence (mask required for 2/3)
// class_pos_offset = 1 reads each class row one position after its <<LABEL>> token (gliclass embed_class_token=false)
cudaError_t head_gather_pool(const void* h_f16, const int64_t* ids, const int64_t* mask, int64_t class_token, int pool_mode,
                             void* pooled_f16, void* cls_f16, int B, int S, int H, int C, cudaStream_t stream,
                             int class_pos_offset = 0, const int32_t* text_row = nullptr);
// (text_row != nullptr: packed layout, text b = rows [text_row[b], text_row[b+1]) of the flat h / ids / mask arrays)
// K5b generalised: logits[b,c] = scale * <t[b*t_stride..], k[b,c,:]> / ((|t|+eps)(|k|+eps) if normalize) + bias, then the
// sigmoid / strict-threshold epilogue.  t_stride = 0: shared weight row (last Linear(K->1) of the MLP scorers).
cudaError_t head_score_ex(const float* t, int64_t t_stride, const float* k, float* logits, float* probs, uint8_t* decisions,
                          float threshold, int B, int C, int K, bool normalize, float eps, float scale, float bias,
                          cudaStream_t stream);
// fp32 rows (row r reads src row r / rep) -> fp16 at dst[r*ld_dst + col0 ..], optional x / (|x| + eps)
cudaError_t head_rows16(const float* src, int K, int rep, void* dst_f16, int64_t ld_dst, int col0, int rows, bool normalize,
                        float eps, cudaStream_t stream);
// weighted-dot scorer: cat[row] = [pt0_b | pl0_row | pt1_b * pl1_row] from (d, half)-interleaved projections
cudaError_t head_wdot_combine(const float* pt, const float* pl, void* cat_f16, int B, int C, int Hh, cudaStream_t stream);
// K5b: logits[b,c] = <t[b,:], k[b,c,:]> (+ optional sigmoid / threshold decisions)
cudaError_t head_score(const float* t, const float* k, float* logits, float* probs, uint8_t* decisions, float threshold,
                       int B, int C, int Hh, cudaStream_t stream);

}  // namespace glc
