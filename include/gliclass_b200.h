/* gliclass_b200.h — native C ABI of the B200 GLiClass engine (libgliclass_b200.so).
 *
 * This is the drop-in boundary for the one hot path GLiClass.c hands to ONNX Runtime:
 *   reference src/model.c:217-281  create_ort_session -> g_ort->CreateSession   => glc_load
 *   reference src/model.c:122-207  run_inference      -> g_ort->Run             => glc_run
 *   reference main.c:186           g_ort->ReleaseSession                        => glc_free
 * The ORT-named entry points the unchanged reference sources link against live in
 * include/onnxruntime_c_api.h (same library); they are thin wrappers over the functions below.
 *
 * Plain pointers and sizes only.  All functions are thread-safe unless noted.  On failure they
 * return NULL / a negative code and glc_last_error() (thread-local) describes why.  There is no
 * CPU fallback: glc_load fails when no sm_100 device is usable.
 */
#ifndef GLICLASS_B200_H
#define GLICLASS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GLC_API __attribute__((visibility("default")))

typedef struct glc_model glc_model;   /* a loaded model replicated on 1..8 GPUs */
typedef struct glc_onnx glc_onnx;     /* host-only parse of a model.onnx (no GPU needed) */

enum { GLC_OK = 0, GLC_ERR = -1, GLC_ERR_ARG = -2, GLC_ERR_CUDA = -3, GLC_ERR_CAPACITY = -4 };
/* storage type of weights and activations (every accumulation / statistic is fp32).  FP16 is the default and the only
 * mode that meets the 2e-2 logit parity bar.  GLC_DTYPE_FP8_E4M3 is an OPT-IN throughput mode for both backbones
 * (the B200 analogue of the reference's int8 quantize_dynamic export, ONNX_CONVERTING/convert_to_onnx.py:81-89): the two
 * FFN / MLP GEMMs of every layer run on e4m3 operands (weights quantised at load, one scale per output channel; LN /
 * RMSNorm output under per-row dynamic scales; GELU or SwiGLU output under a static multiplier), everything else stays fp16.  Its measured
 * logit error is in DESIGN.md "FP8"; bf16 storage is rejected. */
enum { GLC_DTYPE_DEFAULT = 0, GLC_DTYPE_FP16 = 1, GLC_DTYPE_BF16 = 2 /* rejected */, GLC_DTYPE_FP8_E4M3 = 3 };

typedef struct glc_opts {
  uint32_t struct_size;      /* = sizeof(glc_opts) */
  int32_t num_devices;       /* 0 = GLC_DEVICES env or device 0 only */
  int32_t device_ids[8];
  int32_t weight_dtype;      /* GLC_DTYPE_DEFAULT, GLC_DTYPE_FP16 or GLC_DTYPE_FP8_E4M3 (opt-in, see above) */
  int32_t max_tokens;        /* micro-batch cap in tokens per device launch (0 = default 65536) */
  int32_t num_heads;         /* 0 = infer from graph; otherwise it must equal the graph's head count (glc_load fails if not) */
  int32_t preln_f32;         /* 1 = keep the pre-LayerNorm sums (out-proj / FFN2 outputs) in fp32 instead of fp16: for
                              * checkpoints whose dense outputs leave the fp16 range.  Default 0: such a checkpoint makes
                              * glc_run FAIL with an overflow error instead of returning logits from clamped values.
                              * Env GLC_PRELN_F32=1 does the same. */
  int32_t reserved[7];
} glc_opts;

typedef struct glc_info {
  int32_t vocab, hidden, layers, heads, inter, head_hidden, buckets, max_rel_pos;
  float ln_eps;
  int64_t class_token;
  int32_t num_devices;
  int32_t weight_dtype;
  /* head variant detected from the graph (gliclass config: pooling_strategy, scorer_type, normalize_features) */
  int32_t pooling;             /* 0 first, 1 last, 2 avg (masked mean), 3 max (masked) */
  int32_t scorer;              /* 0 simple (dot), 1 mlp, 2 weighted-dot */
  int32_t normalize_features;  /* features x / (|x| + eps), logits * logit_scale */
  float logit_scale;
  int32_t projector_act;       /* projector_hidden_act: 1 erf-GELU, 2 ReLU */
  int32_t class_pos_offset;    /* 0: class rows at the <<LABEL>> positions (embed_class_token=true); 1: one position later */
  int32_t backbone;            /* 0 DeBERTa-v2/v3 encoder, 1 Qwen2-style decoder stack used bidirectionally (Readme.md:91-94) */
  int32_t kv_heads, head_dim;  /* decoder backbones: grouped-query kv heads, head dim (128) */
} glc_info;

GLC_API const char* glc_last_error(void);
GLC_API int glc_device_count(void);          /* usable sm_100 devices; 0 when none (never fails) */

/* ---- the hot path ------------------------------------------------------------------------ */

/* Parse model.onnx, upload weights as fp16 (position-projection tables precomputed per layer). */
GLC_API glc_model* glc_load(const char* onnx_path, const glc_opts* opts /* may be NULL */);
GLC_API void glc_free(glc_model* m);
GLC_API int glc_model_info(const glc_model* m, glc_info* out);

/* C = max over rows of the number of <<LABEL>> tokens (output width of the reference graph). */
GLC_API int glc_num_classes(const glc_model* m, const int64_t* input_ids, int B, int S);

/* One forward.  HOST buffers: input_ids / attention_mask int64 [B,S] row-major (the layout
 * reference flatten_int_array, src/model.c:17-29, produces); logits_out fp32 [B,C] row-major,
 * C returned through C_out.  logits_capacity is in floats; GLC_ERR_CAPACITY if B*C exceeds it.
 * Rows are sharded across the model's devices; the call returns when logits_out is complete. */
GLC_API int glc_run(glc_model* m, const int64_t* input_ids, const int64_t* attention_mask, int B, int S,
                    float* logits_out, size_t logits_capacity, int* C_out);

/* glc_run with the decision epilogue of reference src/postprocessor.c:14-16,93-95 fused into the
 * scorer kernel on the GPU: probs_out [B,C] = 1/(1+expf(-logit)), decisions_out [B,C] (uint8) =
 * prob > threshold (strict).  Any of the three outputs may be NULL; `capacity` is in elements and
 * applies to each non-NULL buffer.  Zero-padded classes (c >= the row's label count) are scored
 * like the reference scores them (postprocessor.c prints them as "[Unknown]"). */
GLC_API int glc_run_decisions(glc_model* m, const int64_t* input_ids, const int64_t* attention_mask, int B, int S,
                              float threshold, float* logits_out /*nullable*/, float* probs_out /*nullable*/,
                              uint8_t* decisions_out /*nullable*/, size_t capacity, int* C_out);

/* ---- asynchronous submit / collect and request coalescing --------------------
 * The reference runs three barriered phases (main.c:116 -> 141 -> 153) and calls Run with BATCH_SIZE=8
 * batches from its OpenMP workers; a B200 needs ~32k tokens per launch.  Two remedies, both behind the
 * same results:
 *  (1) coalescing, on by default (GLC_COALESCE=0 disables): concurrent glc_run / Run callers whose request
 *      is <= GLC_COALESCE_TOKENS (default max_tokens/2) tokens are merged into one padded forward per device
 *      and scattered back; a lone caller pays nothing.  glc_coalesce_stats counts merged launches/requests.
 *  (2) glc_submit returns at once (C is known from the ids, so the caller can size / slice its buffer);
 *      the inputs and logits_out must stay valid until glc_collect(ticket), which waits, frees the ticket
 *      and returns the status.  glc_poll: 1 = finished, 0 = still running. */
typedef struct glc_ticket glc_ticket;
GLC_API glc_ticket* glc_submit(glc_model* m, const int64_t* input_ids, const int64_t* attention_mask, int B, int S,
                               float* logits_out, size_t logits_capacity, int* C_out);
GLC_API int glc_poll(glc_ticket* t);
GLC_API int glc_collect(glc_ticket* t);
GLC_API int glc_coalesce_stats(const glc_model* m, uint64_t* merged_launches, uint64_t* merged_requests);
/* Varlen packing (replaces the pad-to-longest batches of reference src/tokenizer.c:44-54 on the device side): a host
 * request whose attention masks leave >= 10 % of the B*S positions as trailing padding is compacted before the forward —
 * every text keeps positions [0, 1 + last unmasked) rounded up to 128 rows — so embedding, GEMM, LayerNorm and attention
 * rows scale with the real tokens.  Logits are those of the padded layout (padded keys are masked either way, padded
 * query rows are never read).  Applies to glc_run / glc_submit / the ORT-named Run on both backbones for requests of
 * >= 8192 positions; GLC_VARLEN=0 disables it.  glc_packed_stats: packed launches so far, rows they computed, rows the
 * padded layout would have computed. */
GLC_API int glc_packed_stats(const glc_model* m, uint64_t* launches, uint64_t* rows, uint64_t* rows_padded);
/* The host-side plan of that compaction, without a model or a GPU (what the CPU tests check): per text the key length
 * kv_len[b] = 1 + last unmasked position and the rows it would own, text_rows[b] = max(128, kv_len rounded up to 128);
 * launch_of[b] = index of the device launch (at most max_rows_per_launch packed rows each) the text falls into.  Returns
 * the total packed rows, 0 when the request would stay in the padded layout (a class token — or with class_pos_offset = 1
 * the token after it — outside the kept positions, or less than 10 % of the B*S positions saved), -1 on a bad argument.
 * Any output pointer may be NULL. */
GLC_API int64_t glc_pack_plan(const int64_t* input_ids, const int64_t* attention_mask, int B, int S, int64_t class_token,
                              int class_pos_offset, int max_rows_per_launch, int32_t* kv_len, int32_t* text_rows,
                              int32_t* launch_of);

/* Same forward with inputs/outputs already resident on `device` (kernel-only timing, parity
 * tests).  d_logits fp32 [B,C] device memory with C = num_classes (caller computes it with
 * glc_num_classes on the host copy).  Runs on the engine's stream for that device and
 * synchronises it before returning unless `async` is non-zero. */
GLC_API int glc_run_device(glc_model* m, int device_slot, const int64_t* d_input_ids,
                           const int64_t* d_attention_mask, int B, int S, int C, float* d_logits, int async);
GLC_API int glc_sync(glc_model* m, int device_slot);
/* CUDA stream (cudaStream_t) the engine launches on for a device slot, for event timing. */
GLC_API void* glc_stream(glc_model* m, int device_slot);
/* kernels launched by this model since load (all devices) — bench.py's gpu_launches */
GLC_API uint64_t glc_launch_count(const glc_model* m);
/* In-stream profiler: when enabled, every kernel launch of glc_run / glc_run_device on that slot is
 * bracketed by CUDA events on the engine stream.  glc_profile_collect synchronises the stream and
 * ADDS the elapsed milliseconds / launch counts per category into ms[] / n[] (capacity >= 9),
 * returning the number of categories:
 *   0 embed+mask  1 QKV GEMM  2 attention  3 out-proj GEMM  4 residual+LN  5 FFN1 GEMM(+GELU)
 *   6 FFN2 GEMM   7 head projector GEMMs   8 head gather/score */
GLC_API int glc_profile_enable(glc_model* m, int device_slot, int on);
GLC_API int glc_profile_collect(glc_model* m, int device_slot, double* ms, uint64_t* n, int capacity);
/* copy a named intermediate of the last forward on device_slot to host as fp32.
 * names: "emb", "qkv0", "ctx0", "h<l>" .  Returns element count or <0. */
GLC_API int64_t glc_debug_fetch(glc_model* m, int device_slot, const char* name, float* out, size_t capacity);

/* ---- fused decision epilogue (reference src/postprocessor.c:85-150) ------------------------ */
/* multi-label: out_mask[b*C+c] = sigmoid(logit) > threshold (strict); single-label: out_argmax[b]
 * = argmax_c sigmoid(logit) with the reference's initial max_prob = 0 / max_idx = -1. */
GLC_API int glc_decide(const float* logits, int B, int C, float threshold, uint8_t* out_mask /*nullable*/,
                       int32_t* out_argmax /*nullable*/, float* out_prob /*nullable [B,C]*/);

/* ---- host-only model inspection (tests; no GPU needed) ------------------------------------- */
GLC_API glc_onnx* glc_onnx_open(const char* onnx_path);
GLC_API void glc_onnx_close(glc_onnx* h);
GLC_API int glc_onnx_info(const glc_onnx* h, glc_info* out);
/* role-named tensor ("layer.3.q.w", "emb.word", ... see csrc/model_weights.cc); Linear weights
 * are [out,in].  Returns ndim (<=4) or <0; *data stays valid until glc_onnx_close. */
GLC_API int glc_onnx_tensor(const glc_onnx* h, const char* role, const float** data, int64_t dims[4]);
GLC_API int glc_onnx_num_roles(const glc_onnx* h);
GLC_API const char* glc_onnx_role_name(const glc_onnx* h, int i);
/* idx[delta+S-1] = clamp(bucket(delta)+buckets, 0, 2*buckets-1): the single c2p/p2c index table */
GLC_API int glc_rel_index_table(int S, int buckets, int max_pos, int32_t* out /* 2S-1 */);

/* ---- single-kernel entry points on device pointers (parity tests, ncu captures) ------------- */
/* All take a cudaStream_t as void* (NULL = default stream), launch on the CURRENT device and
 * return after launch (no sync).  These are the K1..K5 kernels of the forward (csrc/kernels.h). */

/* K2: C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]); A, W fp16, fp32 accumulate in TMEM (tcgen05 +
 * TMA).  act: 0 none, 1 erf-GELU, 2 ReLU, 3 SwiGLU (W rows interleaved in blocks of 32: gate rows, then the matching up
 * rows; C is [M, N/2] = silu(gate) * up).  out_f32: C is fp32 instead of fp16.  ld* in elements. */
GLC_API int glc_op_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                        int M, int N, int K, int act, int out_f32, void* stream);
/* K2 with the residual add of the layer fused into the epilogue (the "dense(ctx) + x" / "W2.gelu(..) + a" sums of
 * T:49-53, T:408-412): C = act(A W^T + bias) + resid, resid fp16 [M, ldr] or NULL. */
GLC_API int glc_op_gemm_resid(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* resid_f16,
                              int64_t ldr, void* C, int64_t ldc, int M, int N, int K, int act, int out_f32, void* stream);
/* K2 on e4m3 operands (tcgen05 kind::f8f6f4, the GLC_DTYPE_FP8_E4M3 FFN path):
 *   C = act((A8 W8^T) * a_scale[m] * a_const * w_scale[n] + bias);  A8 [M,K], W8 [N,K] e4m3 bytes, K % 16 == 0;
 *   a_scale may be NULL; C fp16, or with out_e4m3 the saturating e4m3 of act(..) * out_mult (act 0 or 1 then).
 * glc_op_quantize_rows_e4m3: q[r,:] = e4m3(x[r,:] * 448 / amax_r), scale[r] = amax_r / 448 (the load-time weight quantiser).
 * glc_op_residual_ln_e4m3: K4 that also writes y as e4m3 under per-row scales (the FFN1 operand). */
GLC_API int glc_op_gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, const float* a_scale, float a_const,
                             const float* w_scale, const float* bias, void* C, int64_t ldc, int M, int N, int K, int act,
                             int out_e4m3, float out_mult, void* stream);
GLC_API int glc_op_quantize_rows_e4m3(const void* x_f16, int64_t ldx, void* q8, int64_t ldq, float* scale, int M, int K,
                                      void* stream);
GLC_API int glc_op_residual_ln_e4m3(const void* x_f16, const void* r_f16, const float* gamma, const float* beta, float eps,
                                    void* y_f16, void* y_e4m3, float* y_scale, int M, int H, void* stream);
/* K1: y[m,:] = (LN(word_emb[ids[m],:]) * gamma + beta) * (mask[m] != 0) */
GLC_API int glc_op_embed_ln(const int64_t* ids, const int64_t* mask, const void* word_emb_f16, const float* gamma,
                            const float* beta, float eps, void* y_f16, int M, int H, int vocab, void* stream);
/* K4: y = LN(x + r) * gamma + beta (fp16 in/out, fp32 statistics); r may be NULL */
GLC_API int glc_op_residual_ln(const void* x_f16, const void* r_f16, const float* gamma, const float* beta,
                               float eps, void* y_f16, int M, int H, void* stream);
/* attention-mask packing: bits[b][w] bit j = mask[b][32w+j] != 0; kv_len[b] = 1 + last valid key */
GLC_API int glc_op_mask_prep(const int64_t* mask, uint32_t* bits, int32_t* kv_len, int B, int S, void* stream);
/* K3: fused disentangled attention for one layer (replaces the attention sub-graph of the ORT session Run, reference
 * src/model.c:173-182; arithmetic T:229-345).  qkv fp16 [B*S,3H] (Q|K|V); mask_bits / kv_len from glc_op_mask_prep; ctx
 * fp16 [B*S,H].  The kernels read the position projections expanded to one row per relative distance:
 *   glc_op_expand_pos      out[rho][0:cols)   = pos[idx(2047 - rho)][0:cols)    (posK half)
 *   glc_op_expand_pos_rev  out[sigma][0:cols) = pos[idx(sigma - 2047)][0:cols)  (posQ half)
 * for rho, sigma in [0, glc_expanded_pos_rows()); idx = glc_rel_index_table, the last row is zero; both synchronise
 * `stream`.  exp_k / exp_qr: fp16 [glc_expanded_pos_rows()][ld_exp], head h at columns h*64...
 *   glc_op_attention_persist  production kernel (csrc/attention_persist.cu): register skews of both biases, one softmax
 *                           thread per query row of a 64-key tile, three warpgroups rotating over the key tiles, one
 *                           persistent CTA per SM, position tables resident in shared memory for S <= 512.
 *   glc_op_attention_naive  slow CUDA-core restatement on the UNEXPANDED tables (pos_k / pos_q fp16 [2*buckets][ld_pos],
 *                           rel_idx int32 [2*Spad-1] from glc_rel_index_table(Spad), Spad = S rounded up to 128): the
 *                           on-GPU debugging oracle of the tests. */
GLC_API int glc_expanded_pos_rows(void);
GLC_API int glc_op_expand_pos(const void* pos_f16, int64_t ld_src, int buckets, int max_pos, void* out_f16, int64_t ld_dst,
                              int cols, void* stream);
GLC_API int glc_op_expand_pos_rev(const void* pos_f16, int64_t ld_src, int buckets, int max_pos, void* out_f16, int64_t ld_dst,
                                  int cols, void* stream);
GLC_API int glc_op_attention_persist(const void* qkv_f16, const void* exp_k_f16, const void* exp_qr_f16, int64_t ld_exp,
                                     const uint32_t* mask_bits, const int32_t* kv_len, void* ctx_f16, int B, int S, int heads,
                                     void* stream);
GLC_API int glc_op_attention_naive(const void* qkv_f16, const void* pos_k_f16, const void* pos_q_f16, int64_t ld_pos,
                                   const int32_t* rel_idx, const uint32_t* mask_bits, void* ctx_f16, int B, int S, int heads,
                                   int buckets, void* stream);
/* ---- decoder-backbone kernels (Qwen2-style stack: reference Readme.md:91-94, BASELINE.json configs[4]) ----
 * glc_op_add_rmsnorm: h fp32 [M,H] += delta (fp16 [M,H], NULL = none); y fp16 = h * rsqrt(mean(h^2) + eps) * g
 * glc_op_rope: rotary embedding in place on the first n_rot_heads heads of qkv fp16 [M, ld] (head j = columns j*head_dim..),
 *              position = row % S, pairs (p, p + head_dim/2), angle = position * inv_freq[p]; synchronises `stream`
 * glc_op_attention_flash128: bidirectional grouped-query flash attention, head dim 128, key padding mask:
 *              qkv fp16 [B*S, (heads + 2 kv_heads) * 128] = Q heads | K heads | V heads; ctx fp16 [B*S, heads * 128]
 * The SwiGLU MLP is glc_op_gemm with act = 3 (W rows interleaved in blocks of 32: gate rows then the matching up rows;
 * C is [M, N/2] = silu(gate) * up). */
GLC_API int glc_op_add_rmsnorm(float* h_f32, const void* delta_f16, const float* g, float eps, void* y_f16, int M, int H,
                               void* stream);
GLC_API int glc_op_rope(void* qkv_f16, int64_t ld, const float* inv_freq, int M, int S, int n_rot_heads, int head_dim,
                        void* stream);
/* glc_op_gemm (act 0) with glc_op_rope applied to columns [0, rope_cols) in the epilogue, head dim 128 (what the engine
 * runs for the decoder's QKV projection): position = row % S; N and rope_cols multiples of 128; synchronises `stream` */
GLC_API int glc_op_gemm_rope(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc,
                             int M, int N, int K, const float* inv_freq, int S, int rope_cols, void* stream);
GLC_API int glc_op_attention_flash128(const void* qkv_f16, const uint32_t* mask_bits, const int32_t* kv_len, void* ctx_f16,
                                      int B, int S, int heads, int kv_heads, void* stream);
/* K5a: pooled[b,:] = h[b,0,:]; cls[b,c,:] = h[b,pos_c(b),:] for the c-th <<LABEL>> token, else 0 */
GLC_API int glc_op_head_gather(const void* h_f16, const int64_t* ids, int64_t class_token, void* pooled_f16,
                               void* cls_f16, int B, int S, int H, int C, void* stream);
/* K5b: logits[b,c] = <t[b,:], k[b,c,:]>; optional probs = sigmoid(logit), decisions = probs > threshold */
GLC_API int glc_op_head_score(const float* t, const float* k, float* logits, float* probs, uint8_t* decisions,
                              float threshold, int B, int C, int Hh, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GLICLASS_B200_H */
