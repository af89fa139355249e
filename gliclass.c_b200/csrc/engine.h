// Engine: per-GPU replica of the model (fp16 weights + precomputed position projections), a
// workspace, one stream, and the forward pass as a fixed sequence of the K1..K5 kernels.  A
// Model owns one DeviceModel per GPU and shards the rows of a batch across them (SURVEY.md §8e:
// batch rows are independent, no collective; the only cross-GPU traffic is the host gather of
// [rows, C] fp32 logits).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <deque>
#include <exception>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "model_weights.h"

namespace glc {

struct DeviceLayer {
  void* wqkv = nullptr;   // fp16 [3H,H]  (Wq | Wk | Wv rows)
  float* bqkv = nullptr;  // [3H]
  void* wo = nullptr;     // fp16 [H,H]
  float* bo = nullptr;
  float *ln1g = nullptr, *ln1b = nullptr;
  void* w1 = nullptr;     // fp16 [I,H]
  float* b1 = nullptr;
  void* w2 = nullptr;     // fp16 [H,I]
  float* b2 = nullptr;
  float *ln2g = nullptr, *ln2b = nullptr;
  // opt-in FP8 FFN (glc_opts.weight_dtype = GLC_DTYPE_FP8_E4M3): e4m3 copies of w1 / w2 with one scale per output channel
  void *w1_8 = nullptr, *w2_8 = nullptr;
  float *w1_s = nullptr, *w2_s = nullptr;
  void* pos_qk = nullptr; // fp16 [2*buckets, 2H]: cols [0,H) = query_proj(rel), [H,2H) = key_proj(rel)
  void* pos_exp = nullptr; // fp16 [expanded_pos_rows(), 2H]: pos_qk expanded to one row per delta (row rho = pos_qk[idx(2047 - rho)]; in shift mode the posQ half [0,H) is stored in the opposite order, row sigma = idx(sigma - 2047))
};

struct DebugBuf { void* ptr = nullptr; size_t count = 0; };

// optional fused decision epilogue (reference src/postprocessor.c:85-150): host destinations, any may be null
struct DecisionOut {
  float* probs = nullptr;        // [B,C] sigmoid(logit)
  uint8_t* decisions = nullptr;  // [B,C] prob > threshold (strict)
  float threshold = 0.5f;
};

// kernel categories for the in-stream profiler (CUDA events around every launch)
enum KernelCat { KC_EMBED = 0, KC_GEMM_QKV, KC_ATTN, KC_GEMM_OUT, KC_LN, KC_GEMM_FFN1, KC_GEMM_FFN2, KC_HEAD_GEMM, KC_HEAD_MISC, KC_COUNT };

// Packed (varlen) layout of one host request (see DeviceModel::run_host): per text the key length (1 + last unmasked
// position) and the rows it owns (rounded up to whole 128-row query tiles), grouped into micro-batches of at most
// `max_rows` packed rows.  Pure host logic, also exported as glc_pack_plan for the CPU tests.
struct PackPlan {
  std::vector<int> len, prow;           // per text: kv length, packed rows (multiple of 128)
  struct MB { int b0, b1, rows, max_rows; };
  std::vector<MB> mbs;                  // micro-batches of at most max_rows packed rows
  int max_mb_rows = 0, max_mb_texts = 0;
  int64_t total_rows = 0;
};
// false when the request does not qualify: a class token (or the neighbour embed_class_token=false reads) outside the kept
// positions, or less than 10 % of the B*S positions saved
bool make_pack_plan(const int64_t* ids, const int64_t* mask, int B, int S, int64_t class_token, int class_pos_offset, int max_rows,
                    PackPlan& pl);

class DeviceModel {
 public:
  DeviceModel(int device, const ModelWeights& w, int max_tokens, bool preln_f32 = false, bool fp8_ffn = false);
  ~DeviceModel();
  DeviceModel(const DeviceModel&) = delete;

  // device-resident inputs; logits fp32 [B,C] on device.  Enqueues on stream(); no sync.
  void forward(const int64_t* d_ids, const int64_t* d_mask, int B, int S, int C, float* d_logits, float* d_probs = nullptr,
               uint8_t* d_decisions = nullptr, float threshold = 0.5f, bool cacheable = true);
  // host buffers: rows [0,B) of ids/mask, writes logits rows [0,B) (width C); micro-batches by
  // max_tokens; synchronises before returning.  Serialised per device by `mu`.
  void run_host(const int64_t* ids, const int64_t* mask, int B, int S, int C, float* logits,
                const DecisionOut* dec = nullptr, bool cacheable = true);

  // Request coalescing: the reference calls Run from its OpenMP workers with BATCH_SIZE=8
  // batches (main.c:141-150), each far too small to fill a B200.  Concurrent callers queue here; whichever
  // caller finds the device idle becomes the leader, takes every queued request that fits max_tokens, pads
  // them to the longest sequence of the group (pad id 0 / mask 0, exactly what tokenizer.c:78-82 does inside
  // a batch), runs ONE forward and scatters each request's rows [B_i, C_i] back.  A lone caller runs directly
  // from its own buffers (no copy, no thread hand-off).
  struct HostReq {
    const int64_t* ids = nullptr;
    const int64_t* mask = nullptr;
    int B = 0, S = 0, C = 0;
    float* logits = nullptr;
    const DecisionOut* dec = nullptr;
    bool done = false;
    std::exception_ptr err;
  };
  void run_host_coalesced(HostReq& r);
  uint64_t merged_groups() const { return merged_groups_.load(); }
  uint64_t merged_requests() const { return merged_requests_.load(); }
  // packed (varlen) launches so far: count, rows computed, rows the padded [B,S] layout would have computed
  void packed_stats(uint64_t* runs, uint64_t* rows, uint64_t* rows_padded) const {
    *runs = packed_runs_.load(); *rows = packed_rows_.load(); *rows_padded = packed_rows_padded_.load();
  }

  int device() const { return device_; }
  cudaStream_t stream() const { return stream_; }
  // throws when the last forwards on this device saturated an fp16 pre-LN sum (see kernels.h residual_ln); clears the flag
  void check_overflow_sync();
  uint64_t launches() const { return launches_.load(); }
  int64_t debug_fetch(const std::string& name, float* out, size_t capacity);
  // profiler: when enabled every launch is bracketed by CUDA events on stream(); collect() syncs
  // the stream, adds the elapsed times per category into ms[KC_COUNT] / n[KC_COUNT] and resets.
  void profile_enable(bool on);
  void profile_collect(double* ms, uint64_t* n);
  std::mutex mu;
  std::vector<cudaEvent_t> free_events_;   // completion events of run_host (guarded by mu)

 private:
  void init_deberta(const ModelWeights& w);
  void init_qwen2(const ModelWeights& w);
  void init_head(const ModelWeights& w);
  const void* rope_table_for(int S);
  void ensure_workspace(int tokens, int B, int C);
  void run_group(std::vector<HostReq*>& group);
  // ---- packed (varlen) layout of one host request: padding rows are dropped before the forward --------------------------
  // A text keeps positions [0, len) (len = 1 + last non-zero mask entry) rounded up to whole 128-row query tiles; GEMM, LN
  // and embedding rows then scale with the real tokens instead of B * S (the reference pads to the longest text of the
  // batch, tokenizer.c:44-54, and ORT computes every padded row).  Nothing a valid row reads changes: padded keys are
  // masked in both layouts and padded query rows are never read by the head.  Used by run_host when it removes >= 10 % of
  // the rows; needs a pooling other than 'last' (which reads padded position S-1) and
  // every class token inside the kept rows.  GLC_VARLEN=0 turns it off.
  struct PackedCtx {                      // device-side description of the micro-batch being run
    int rows = 0, max_rows = 0, n_tiles = 0;
    const int32_t *text_row = nullptr, *kv_len = nullptr, *tile_info = nullptr, *tile_pos = nullptr;
  };
  bool plan_pack(const int64_t* ids, const int64_t* mask, int B, int S, PackPlan& pl) const;
  bool varlen_ = true;
  const PackedCtx* pk_ = nullptr;         // set (under mu) around forward_eager for a packed micro-batch
  int32_t* pk_ints_ = nullptr;            // device: text_row[nb+1] | kv_len[nb] | tile_info[n_tiles] | tile_pos[n_tiles]
  int32_t* pk_scratch_ = nullptr;
  void forward_eager(const int64_t* d_ids, const int64_t* d_mask, int B, int S, int C, float* d_logits, float* d_probs,
                     uint8_t* d_decisions, float threshold);
  const int32_t* rel_table(int S);
  void* dalloc(size_t bytes);
  void upload_f32(float** dst, const HostTensor& t);
  void upload_w16(void** dst, const float* src, size_t n);
  void keep(const char* name, const void* src_f16, size_t count);

  int device_ = 0;
  int num_sms_ = 148;
  cudaStream_t stream_ = nullptr;
  ModelConfig cfg_;
  int max_tokens_ = 65536;
  bool debug_keep_ = false;
  bool fuse_resid_ = false;    // GLC_FUSE_RESID=1: residual add in the out-proj / FFN2 GEMM epilogues instead of the LN kernel
  bool fp8_ffn_ = false;       // FFN1 / FFN2 on e4m3 operands (kind::f8f6f4), everything else fp16; off by default (DESIGN.md "FP8")
  float fp8_mult_ = 4.0f;      // static multiplier the e4m3 GELU output is stored with (saturates at 448 / mult; GLC_FP8_MULT)
  void* x1_8_ = nullptr;       // e4m3 [tokens, H]: LN1 output under per-row scales x1_s_
  float* x1_s_ = nullptr;
  bool preln_f32_ = false;     // GLC_PRELN_F32=1 / glc_opts.preln_f32: out-proj and FFN2 outputs (the pre-LN sums) stay fp32
  int* d_overflow_ = nullptr;  // set by the LN kernels when a pre-LN sum hit the fp16 saturation value
  int* h_overflow_ = nullptr;  // pinned ring of per-request copies of the flag
  uint32_t overflow_seq_ = 0;
  // pinned staging for pageable caller buffers (the reference's malloc'd tensors, src/model.c:17-29): two input slots
  // (ids | mask of one micro-batch) used alternately, each guarded by the event of its last H2D copy; output blocks are
  // taken from a free list per request and copied to the caller after the request's completion event
  struct PinnedBlock { void* p = nullptr; size_t bytes = 0; };
  PinnedBlock h_in_[2];
  cudaEvent_t h_in_ev_[2] = {nullptr, nullptr};
  int h_in_next_ = 0;
  std::vector<PinnedBlock> out_pool_;
  PinnedBlock take_out_block(size_t bytes);
  std::atomic<uint64_t> launches_{0};

  // weights
  void* word_emb_ = nullptr;
  float *emb_g_ = nullptr, *emb_b_ = nullptr;
  std::vector<DeviceLayer> layers_;
  void *t1w_ = nullptr, *t2w_ = nullptr, *c1w_ = nullptr, *c2w_ = nullptr;
  float *t1b_ = nullptr, *t2b_ = nullptr, *c1b_ = nullptr, *c2b_ = nullptr;
  // scorer variants: MLP (m1 -> m2 -> m3) or weighted dot (pt, pl projections; o1 -> o2)
  void *m1w_ = nullptr, *m2w_ = nullptr, *ptw_ = nullptr, *plw_ = nullptr, *o1w_ = nullptr;
  float *m1b_ = nullptr, *m2b_ = nullptr, *ptb_ = nullptr, *plb_ = nullptr, *o1b_ = nullptr;
  float* last_w_ = nullptr;     // fp32 weight row of the final Linear(K -> 1)
  float last_b_ = 0.f;
  std::map<int, int32_t*> rel_tables_;   // keyed by Spad
  // decoder backbone (Qwen2): final RMSNorm weight, rotary inverse frequencies, (cos, sin) tables per sequence length
  float *norm_g_ = nullptr, *rope_inv_freq_ = nullptr;
  std::map<int, void*> rope_tables_;
  float* h32_ = nullptr;                 // fp32 residual stream [M,H]

  // workspace
  int ws_tokens_ = 0, ws_B_ = 0, ws_rows_ = 0;
  int64_t *ids_ = nullptr, *mask_ = nullptr;
  void *x_ = nullptr, *x1_ = nullptr, *qkv_ = nullptr, *ctx_ = nullptr, *tmp_ = nullptr, *ffn_ = nullptr;
  uint32_t* mask_bits_ = nullptr;
  int32_t* kv_len_ = nullptr;
  void *pooled_ = nullptr, *cls_ = nullptr, *tmid_ = nullptr, *cmid_ = nullptr;
  float *tvec_ = nullptr, *kvec_ = nullptr, *logits_ = nullptr, *probs_ = nullptr;
  uint8_t* decisions_ = nullptr;
  void *cat16_ = nullptr, *s1_ = nullptr, *t16_ = nullptr, *k16_ = nullptr;   // scorer-variant scratch
  float *s2_ = nullptr, *pt_ = nullptr, *pl_ = nullptr;
  std::vector<void*> ws_allocs_, perm_allocs_;
  std::map<std::string, DebugBuf> debug_;
  // CUDA graphs of the forward, one per (shape, buffer set): the kernel sequence of a forward is fixed,
  // so replaying it removes ~100 launches of CPU work per Run (what a batch-8 Run is bound by)
  struct GraphKey {
    int B, S, C;
    const void *ids, *mask, *logits, *probs, *dec;
    uint32_t thr;
    bool operator<(const GraphKey& o) const {
      return std::tie(B, S, C, ids, mask, logits, probs, dec, thr) <
             std::tie(o.B, o.S, o.C, o.ids, o.mask, o.logits, o.probs, o.dec, o.thr);
    }
  };
  struct GraphEntry { int seen = 0; cudaGraphExec_t exec = nullptr; uint64_t launches = 0; uint64_t last_use = 0; };
  std::map<GraphKey, GraphEntry> graphs_;
  uint64_t graph_clock_ = 0;
  int graphs_live_ = 0;        // instantiated graphs (placeholders of shapes seen once do not count)
  static constexpr int kMaxGraphs = 48, kMaxGraphKeys = 512;
  bool graphs_on_ = true;
  void drop_graphs();
  // coalescing queue + pinned staging of merged groups (touched only by the current leader)
  std::mutex qmu_;
  std::condition_variable qcv_;
  std::deque<HostReq*> queue_;
  bool leader_ = false;
  int64_t *h_ids_ = nullptr, *h_mask_ = nullptr;
  float *h_logits_ = nullptr, *h_probs_ = nullptr;
  uint8_t* h_dec_ = nullptr;
  size_t h_tok_ = 0, h_rows_ = 0;
  std::atomic<uint64_t> packed_runs_{0}, packed_rows_{0}, packed_rows_padded_{0};
  std::atomic<uint64_t> merged_groups_{0}, merged_requests_{0};
  int last_group_size_ = 1;    // under qmu_: size of the previous coalesced group (sizes the leader's batching window)

  // profiler state
  struct ProfRec { int cat; cudaEvent_t a, b; };
  bool prof_on_ = false;
  std::vector<ProfRec> prof_recs_;
  std::vector<cudaEvent_t> prof_pool_;
  cudaEvent_t prof_event();
  friend struct ProfScope;
};

// one persistent host thread per device of a multi-GPU model: runs that device's row shard of a large Run
// (replaces a std::thread spawned per call) and the asynchronous glc_submit requests
class TaskQueue {
 public:
  explicit TaskQueue(int threads);
  ~TaskQueue();
  void post(std::function<void()> fn);

 private:
  void loop();
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::function<void()>> q_;
  std::vector<std::thread> th_;
  bool stop_ = false;
};

class Model {
 public:
  Model(const std::string& onnx_path, const std::vector<int>& devices, int max_tokens, bool preln_f32 = false, bool fp8_ffn = false);
  ~Model();
  TaskQueue& submit_queue();
  int num_classes(const int64_t* ids, int B, int S) const;
  void run(const int64_t* ids, const int64_t* mask, int B, int S, int C, float* logits, const DecisionOut* dec = nullptr);
  const ModelConfig& cfg() const { return cfg_; }
  int num_devices() const { return (int)devs_.size(); }
  void coalesce_stats(uint64_t* groups, uint64_t* requests) const;
  DeviceModel& dev(int slot) { return *devs_[slot]; }
  uint64_t launches() const;

 private:
  ModelConfig cfg_;
  std::vector<std::unique_ptr<DeviceModel>> devs_;
  std::atomic<uint32_t> rr_{0};
  std::vector<std::unique_ptr<TaskQueue>> workers_;   // one single-thread queue per device (multi-GPU models only)
  std::unique_ptr<TaskQueue> submit_q_;               // glc_submit requests (created on first use)
  std::mutex submit_mu_;
  bool coalesce_ = true;        // GLC_COALESCE=0 disables
  int coalesce_tokens_ = 0;     // requests up to this many tokens go through the coalescing queue
};

int usable_device_count();   // sm_100 devices; 0 when there is no driver / GPU

}  // namespace glc
