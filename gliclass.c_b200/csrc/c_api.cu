// Native C ABI (include/gliclass_b200.h).  Every entry point catches C++ exceptions and turns
// them into an error code + thread-local message, as a C caller (the reference's model.c) expects.
#include <cuda_fp16.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <sstream>
#include <string>
#include <vector>

#include "engine.h"
#include "gliclass_b200.h"
#include "kernels.h"
#include "model_weights.h"

struct glc_model {
  glc::Model* m;
  int weight_dtype;
};
struct glc_ticket {
  std::promise<int> done;
  std::future<int> fut;
  std::string err;
  int C = 0;
};
struct glc_onnx {
  glc::ModelWeights w;
  std::vector<std::string> roles;
};

namespace {
thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
void fill_info(const glc::ModelConfig& c, glc_info* o) {
  memset(o, 0, sizeof(*o));
  o->vocab = c.vocab; o->hidden = c.hidden; o->layers = c.layers; o->heads = c.heads; o->inter = c.inter;
  o->head_hidden = c.head_hidden; o->buckets = c.buckets; o->max_rel_pos = c.max_rel_pos; o->ln_eps = c.ln_eps;
  o->class_token = c.class_token;
  o->pooling = c.pooling; o->scorer = c.scorer; o->normalize_features = c.normalize ? 1 : 0; o->logit_scale = c.logit_scale;
  o->projector_act = c.proj_act; o->class_pos_offset = c.class_pos_offset;
  o->backbone = c.backbone; o->kv_heads = c.kv_heads; o->head_dim = c.backbone == glc::BACKBONE_QWEN2 ? c.head_dim : (c.heads ? c.hidden / c.heads : 0);
}
}  // namespace

extern "C" {

const char* glc_last_error(void) { return g_err.c_str(); }

int glc_device_count(void) { return glc::usable_device_count(); }

glc_model* glc_load(const char* onnx_path, const glc_opts* opts) {
  try {
    if (!onnx_path) { fail(GLC_ERR_ARG, "glc_load: null path"); return nullptr; }
    std::vector<int> devices;
    int max_tokens = 0;
    int dtype = GLC_DTYPE_FP16;
    int want_heads = 0;
    bool preln_f32 = false;

    if (opts) {
      if (opts->struct_size != sizeof(glc_opts)) {
        fail(GLC_ERR_ARG, "glc_load: glc_opts.struct_size does not match this library's sizeof(glc_opts) (header / library mismatch)");
        return nullptr;
      }
      if (opts->num_devices < 0 || opts->num_devices > 8 || opts->max_tokens < 0 || opts->num_heads < 0) {
        fail(GLC_ERR_ARG, "glc_load: negative or out-of-range field in glc_opts");
        return nullptr;
      }
      for (int i = 0; i < opts->num_devices; ++i) devices.push_back(opts->device_ids[i]);
      max_tokens = opts->max_tokens;
      want_heads = opts->num_heads;
      preln_f32 = opts->preln_f32 != 0;
      if (opts->weight_dtype != GLC_DTYPE_DEFAULT) dtype = opts->weight_dtype;
    }
    if (dtype != GLC_DTYPE_FP16 && dtype != GLC_DTYPE_FP8_E4M3) {
      fail(GLC_ERR_ARG,
           "glc_load: weight_dtype must be GLC_DTYPE_FP16 (default) or GLC_DTYPE_FP8_E4M3 (opt-in: FFN weights and FFN "
           "activations in e4m3, everything else fp16); bf16 storage is not implemented (it cannot meet the 2e-2 parity bar "
           "and tcgen05 kind::f16 rejects fp16 x bf16 operands)");
      return nullptr;
    }
    if (devices.empty()) {
      // GLC_DEVICES="0,1,2" or "all"; default: device 0
      const char* env = getenv("GLC_DEVICES");
      if (env && *env) {
        if (!strcmp(env, "all")) {
          int n = 0;
          cudaGetDeviceCount(&n);
          for (int d = 0; d < n; ++d) devices.push_back(d);
        } else {
          std::stringstream ss(env);
          std::string tok;
          while (std::getline(ss, tok, ',')) {
            if (tok.empty()) continue;
            char* endp = nullptr;
            const long v = strtol(tok.c_str(), &endp, 10);
            if (*endp != '\0' || v < 0 || v > 1023) {
              fail(GLC_ERR_ARG, "glc_load: GLC_DEVICES must be 'all' or a comma-separated list of device ordinals (got '" + std::string(env) + "')");
              return nullptr;
            }
            devices.push_back((int)v);
          }
        }
      }
      if (devices.empty()) devices.push_back(0);
    }
    if (glc::usable_device_count() == 0) {
      fail(GLC_ERR_CUDA, "glc_load: no usable sm_100 CUDA device (this engine has no CPU fallback)");
      return nullptr;
    }
    {
      int ndev = 0;
      cudaGetDeviceCount(&ndev);
      for (size_t i = 0; i < devices.size(); ++i) {
        if (devices[i] < 0 || devices[i] >= ndev) {
          fail(GLC_ERR_ARG, "glc_load: device ordinal " + std::to_string(devices[i]) + " out of range (" + std::to_string(ndev) + " visible)");
          return nullptr;
        }
        for (size_t j = 0; j < i; ++j)
          if (devices[j] == devices[i]) { fail(GLC_ERR_ARG, "glc_load: device listed twice"); return nullptr; }
      }
    }
    if (const char* mt = getenv("GLC_MAX_TOKENS")) if (max_tokens <= 0) max_tokens = atoi(mt);
    glc_model* h = new glc_model;
    h->m = nullptr;
    try {
      h->m = new glc::Model(onnx_path, devices, max_tokens, preln_f32, dtype == GLC_DTYPE_FP8_E4M3);
      if (want_heads > 0 && want_heads != h->m->cfg().heads)
        throw std::runtime_error("glc_opts.num_heads = " + std::to_string(want_heads) + " but the graph has " +
                                 std::to_string(h->m->cfg().heads) + " attention heads");
    } catch (...) {
      delete h->m;
      delete h;
      throw;
    }
    h->weight_dtype = dtype;
    return h;
  } catch (const std::exception& e) {
    fail(GLC_ERR, std::string("glc_load: ") + e.what());
    return nullptr;
  }
}

void glc_free(glc_model* m) {
  if (!m) return;
  try { delete m->m; } catch (...) {}
  delete m;
}

int glc_model_info(const glc_model* m, glc_info* out) {
  if (!m || !out) return fail(GLC_ERR_ARG, "glc_model_info: null argument");
  fill_info(m->m->cfg(), out);
  out->num_devices = m->m->num_devices();
  out->weight_dtype = m->weight_dtype;
  return GLC_OK;
}

int glc_num_classes(const glc_model* m, const int64_t* input_ids, int B, int S) {
  if (!m || (!input_ids && B * S > 0) || B < 0 || S < 0) return fail(GLC_ERR_ARG, "glc_num_classes: bad argument");
  return m->m->num_classes(input_ids, B, S);
}

int glc_run(glc_model* m, const int64_t* input_ids, const int64_t* attention_mask, int B, int S, float* logits_out,
            size_t logits_capacity, int* C_out) {
  try {
    if (!m || B < 0 || S < 0) return fail(GLC_ERR_ARG, "glc_run: bad argument");
    if (B * S > 0 && (!input_ids || !attention_mask)) return fail(GLC_ERR_ARG, "glc_run: null input");
    const int C = m->m->num_classes(input_ids, B, S);
    if (C_out) *C_out = C;
    if ((size_t)B * C > logits_capacity) return fail(GLC_ERR_CAPACITY, "glc_run: logits buffer too small");
    if (B == 0 || S == 0 || C == 0) return GLC_OK;
    if (!logits_out) return fail(GLC_ERR_ARG, "glc_run: null output");
    m->m->run(input_ids, attention_mask, B, S, C, logits_out);
    return GLC_OK;
  } catch (const std::invalid_argument& e) {
    return fail(GLC_ERR_ARG, std::string("glc_run: ") + e.what());
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_run: ") + e.what());
  }
}

int glc_run_decisions(glc_model* m, const int64_t* input_ids, const int64_t* attention_mask, int B, int S, float threshold,
                      float* logits_out, float* probs_out, uint8_t* decisions_out, size_t capacity, int* C_out) {
  try {
    if (!m || B < 0 || S < 0) return fail(GLC_ERR_ARG, "glc_run_decisions: bad argument");
    if (B * S > 0 && (!input_ids || !attention_mask)) return fail(GLC_ERR_ARG, "glc_run_decisions: null input");
    const int C = m->m->num_classes(input_ids, B, S);
    if (C_out) *C_out = C;
    if ((size_t)B * C > capacity) return fail(GLC_ERR_CAPACITY, "glc_run_decisions: output buffers too small");
    if (B == 0 || S == 0 || C == 0) return GLC_OK;
    if (!logits_out && !probs_out && !decisions_out) return fail(GLC_ERR_ARG, "glc_run_decisions: no output requested");
    glc::DecisionOut d;
    d.probs = probs_out;
    d.decisions = decisions_out;
    d.threshold = threshold;
    m->m->run(input_ids, attention_mask, B, S, C, logits_out, &d);
    return GLC_OK;
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_run_decisions: ") + e.what());
  }
}

int glc_run_device(glc_model* m, int slot, const int64_t* d_ids, const int64_t* d_mask, int B, int S, int C,
                   float* d_logits, int async) {
  try {
    if (!m || slot < 0 || slot >= m->m->num_devices() || B < 0 || S < 0 || C < 0)
      return fail(GLC_ERR_ARG, "glc_run_device: bad argument");
    glc::DeviceModel& d = m->m->dev(slot);
    std::lock_guard<std::mutex> lk(d.mu);
    cudaSetDevice(d.device());
    d.forward(d_ids, d_mask, B, S, C, d_logits);
    if (!async) {
      cudaError_t e = cudaStreamSynchronize(d.stream());
      if (e != cudaSuccess) return fail(GLC_ERR_CUDA, std::string("glc_run_device: ") + cudaGetErrorString(e));
    }
    return GLC_OK;
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_run_device: ") + e.what());
  }
}

int glc_sync(glc_model* m, int slot) {
  if (!m || slot < 0 || slot >= m->m->num_devices()) return fail(GLC_ERR_ARG, "glc_sync: bad argument");
  try {
    m->m->dev(slot).check_overflow_sync();   // takes the device lock (a stream capture may be in progress on another thread)
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_sync: ") + e.what());
  }
  return GLC_OK;
}

void* glc_stream(glc_model* m, int slot) {
  if (!m || slot < 0 || slot >= m->m->num_devices()) return nullptr;
  return (void*)m->m->dev(slot).stream();
}

uint64_t glc_launch_count(const glc_model* m) { return m ? m->m->launches() : 0; }

glc_ticket* glc_submit(glc_model* m, const int64_t* input_ids, const int64_t* attention_mask, int B, int S, float* logits_out,
                       size_t logits_capacity, int* C_out) {
  try {
    if (!m || B < 0 || S < 0 || (B * S > 0 && (!input_ids || !attention_mask))) {
      fail(GLC_ERR_ARG, "glc_submit: bad argument");
      return nullptr;
    }
    const int C = m->m->num_classes(input_ids, B, S);
    if (C_out) *C_out = C;
    if ((size_t)B * C > logits_capacity || ((size_t)B * C > 0 && !logits_out)) {
      fail(GLC_ERR_CAPACITY, "glc_submit: logits buffer too small");
      return nullptr;
    }
    glc_ticket* t = new glc_ticket;
    t->C = C;
    glc::Model* mm = m->m;
    if (B == 0 || S == 0 || C == 0) {
      t->fut = t->done.get_future();
      t->done.set_value(GLC_OK);
      return t;
    }
    // a fixed pool of host threads (GLC_SUBMIT_THREADS, default 8) serves the in-flight requests: each blocks in the
    // coalescing queue / on the device like a reference OpenMP worker would block in g_ort->Run, while the
    // submitting thread goes on tokenising
    t->fut = t->done.get_future();
    mm->submit_queue().post([mm, t, input_ids, attention_mask, B, S, C, logits_out]() {
      int rc = GLC_OK;
      try {
        mm->run(input_ids, attention_mask, B, S, C, logits_out);
      } catch (const std::exception& e) {
        t->err = std::string("glc_submit: ") + e.what();
        rc = GLC_ERR_CUDA;
      }
      t->done.set_value(rc);
    });
    return t;
  } catch (const std::exception& e) {
    fail(GLC_ERR, std::string("glc_submit: ") + e.what());
    return nullptr;
  }
}

int glc_poll(glc_ticket* t) {
  if (!t) return fail(GLC_ERR_ARG, "glc_poll: null ticket");
  return t->fut.wait_for(std::chrono::seconds(0)) == std::future_status::ready ? 1 : 0;
}

int glc_collect(glc_ticket* t) {
  if (!t) return fail(GLC_ERR_ARG, "glc_collect: null ticket");
  int rc = GLC_ERR;
  try {
    rc = t->fut.get();
  } catch (const std::exception& e) {
    t->err = std::string("glc_collect: ") + e.what();
  }
  if (rc != GLC_OK) g_err = t->err;
  delete t;
  return rc;
}

int glc_coalesce_stats(const glc_model* m, uint64_t* groups, uint64_t* requests) {
  if (!m) return fail(GLC_ERR_ARG, "glc_coalesce_stats: null model");
  m->m->coalesce_stats(groups, requests);
  return GLC_OK;
}

int64_t glc_pack_plan(const int64_t* input_ids, const int64_t* attention_mask, int B, int S, int64_t class_token, int class_pos_offset,
                      int max_rows_per_launch, int32_t* kv_len, int32_t* text_rows, int32_t* launch_of) {
  if (!input_ids || !attention_mask || B <= 0 || S <= 0 || class_pos_offset < 0 || max_rows_per_launch <= 0) {
    fail(GLC_ERR_ARG, "glc_pack_plan: bad argument");
    return -1;
  }
  glc::PackPlan pl;
  const bool ok = glc::make_pack_plan(input_ids, attention_mask, B, S, class_token, class_pos_offset, max_rows_per_launch, pl);
  if ((int)pl.len.size() == B) {   // (a class token in the padded tail stops the scan early: nothing to report then)
    for (int b = 0; b < B; ++b) {
      if (kv_len) kv_len[b] = pl.len[b];
      if (text_rows) text_rows[b] = pl.prow[b];
    }
  }
  if (!ok) return 0;
  if (launch_of)
    for (size_t i = 0; i < pl.mbs.size(); ++i)
      for (int b = pl.mbs[i].b0; b < pl.mbs[i].b1; ++b) launch_of[b] = (int)i;
  return pl.total_rows;
}

int glc_packed_stats(const glc_model* m, uint64_t* launches, uint64_t* rows, uint64_t* rows_padded) {
  if (!m) return fail(GLC_ERR_ARG, "glc_packed_stats: null model");
  uint64_t a = 0, b = 0, c = 0;
  for (int i = 0; i < m->m->num_devices(); ++i) {
    uint64_t x, y, z;
    m->m->dev(i).packed_stats(&x, &y, &z);
    a += x; b += y; c += z;
  }
  if (launches) *launches = a;
  if (rows) *rows = b;
  if (rows_padded) *rows_padded = c;
  return GLC_OK;
}

int glc_profile_enable(glc_model* m, int slot, int on) {
  if (!m || slot < 0 || slot >= m->m->num_devices()) return fail(GLC_ERR_ARG, "glc_profile_enable: bad argument");
  m->m->dev(slot).profile_enable(on != 0);
  return GLC_OK;
}

int glc_profile_collect(glc_model* m, int slot, double* ms, uint64_t* n, int capacity) {
  try {
    if (!m || !ms || !n || slot < 0 || slot >= m->m->num_devices() || capacity < (int)glc::KC_COUNT)
      return fail(GLC_ERR_ARG, "glc_profile_collect: bad argument");
    m->m->dev(slot).profile_collect(ms, n);
    return (int)glc::KC_COUNT;
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_profile_collect: ") + e.what());
  }
}

int64_t glc_debug_fetch(glc_model* m, int slot, const char* name, float* out, size_t capacity) {
  try {
    if (!m || !name || slot < 0 || slot >= m->m->num_devices()) return fail(GLC_ERR_ARG, "glc_debug_fetch: bad argument");
    return m->m->dev(slot).debug_fetch(name, out, capacity);
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_debug_fetch: ") + e.what());
  }
}

// reference src/postprocessor.c:14-16 (sigmoid), :93-95 (strict >), :119-128 (argmax from 0 / -1)
int glc_decide(const float* logits, int B, int C, float threshold, uint8_t* out_mask, int32_t* out_argmax, float* out_prob) {
  if (B < 0 || C < 0 || (!logits && B * C > 0)) return fail(GLC_ERR_ARG, "glc_decide: bad argument");
  for (int i = 0; i < B; ++i) {
    float max_prob = 0.0f;
    int max_idx = -1;
    for (int j = 0; j < C; ++j) {
      const float prob = 1.0f / (1.0f + expf(-logits[(size_t)i * C + j]));
      if (out_prob) out_prob[(size_t)i * C + j] = prob;
      if (out_mask) out_mask[(size_t)i * C + j] = prob > threshold ? 1 : 0;
      if (prob > max_prob) { max_prob = prob; max_idx = j; }
    }
    if (out_argmax) out_argmax[i] = max_idx;
  }
  return GLC_OK;
}

// ---- host-only inspection --------------------------------------------------------------------

glc_onnx* glc_onnx_open(const char* path) {
  try {
    if (!path) { fail(GLC_ERR_ARG, "glc_onnx_open: null path"); return nullptr; }
    glc_onnx* h = new glc_onnx;
    try {
      glc::load_model_weights(path, &h->w);
    } catch (...) {
      delete h;
      throw;
    }
    for (auto& kv : h->w.t) h->roles.push_back(kv.first);
    return h;
  } catch (const std::exception& e) {
    fail(GLC_ERR, std::string("glc_onnx_open: ") + e.what());
    return nullptr;
  }
}
void glc_onnx_close(glc_onnx* h) { delete h; }
int glc_onnx_info(const glc_onnx* h, glc_info* out) {
  if (!h || !out) return fail(GLC_ERR_ARG, "glc_onnx_info: null argument");
  fill_info(h->w.cfg, out);
  return GLC_OK;
}
int glc_onnx_tensor(const glc_onnx* h, const char* role, const float** data, int64_t dims[4]) {
  if (!h || !role || !data || !dims) return fail(GLC_ERR_ARG, "glc_onnx_tensor: null argument");
  auto it = h->w.t.find(role);
  if (it == h->w.t.end()) return fail(GLC_ERR_ARG, std::string("glc_onnx_tensor: unknown role ") + role);
  const glc::HostTensor& t = it->second;
  if (t.dims.size() > 4) return fail(GLC_ERR, "glc_onnx_tensor: rank > 4");
  for (size_t i = 0; i < t.dims.size(); ++i) dims[i] = t.dims[i];
  *data = t.data.data();
  return (int)t.dims.size();
}
int glc_onnx_num_roles(const glc_onnx* h) { return h ? (int)h->roles.size() : 0; }
const char* glc_onnx_role_name(const glc_onnx* h, int i) {
  if (!h || i < 0 || i >= (int)h->roles.size()) return nullptr;
  return h->roles[i].c_str();
}
int glc_rel_index_table(int S, int buckets, int max_pos, int32_t* out) {
  if (S <= 0 || buckets <= 1 || max_pos <= 1 || !out) return fail(GLC_ERR_ARG, "glc_rel_index_table: bad argument");
  glc::rel_index_table(S, buckets, max_pos, out);
  return GLC_OK;
}

// ---- single-kernel entry points --------------------------------------------------------------

static int num_sms_current() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
static int wrap(const char* what, cudaError_t e) {
  if (e == cudaSuccess) return GLC_OK;
  return fail(GLC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define GLC_TRY(tag, expr)                                        \
  try {                                                           \
    return wrap(tag, (expr));                                     \
  } catch (const std::exception& e) {                             \
    return fail(GLC_ERR_CUDA, std::string(tag) + ": " + e.what()); \
  }

int glc_op_gemm(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N,
                int K, int act, int out_f32, void* stream) {
  GLC_TRY("glc_op_gemm", glc::gemm_f16(A, lda, W, ldw, bias, C, ldc, M, N, K, act, out_f32 != 0, num_sms_current(),
                                       (cudaStream_t)stream));
}
int glc_op_gemm_resid(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, const void* resid, int64_t ldr,
                      void* C, int64_t ldc, int M, int N, int K, int act, int out_f32, void* stream) {
  GLC_TRY("glc_op_gemm_resid", glc::gemm_f16_resid(A, lda, W, ldw, bias, resid, ldr, C, ldc, M, N, K, act, out_f32 != 0,
                                                   num_sms_current(), (cudaStream_t)stream));
}
int glc_op_gemm_e4m3(const void* A8, int64_t lda, const void* W8, int64_t ldw, const float* a_scale, float a_const,
                     const float* w_scale, const float* bias, void* C, int64_t ldc, int M, int N, int K, int act, int out_e4m3,
                     float out_mult, void* stream) {
  GLC_TRY("glc_op_gemm_e4m3", glc::gemm_e4m3(A8, lda, W8, ldw, a_scale, a_const, w_scale, bias, C, ldc, M, N, K, act, out_e4m3 != 0,
                                             out_mult, num_sms_current(), (cudaStream_t)stream));
}
int glc_op_quantize_rows_e4m3(const void* x_f16, int64_t ldx, void* q8, int64_t ldq, float* scale, int M, int K, void* stream) {
  GLC_TRY("glc_op_quantize_rows_e4m3", glc::quantize_rows_e4m3(x_f16, ldx, q8, ldq, scale, M, K, (cudaStream_t)stream));
}
int glc_op_residual_ln_e4m3(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y, void* y8,
                            float* y8_scale, int M, int H, void* stream) {
  GLC_TRY("glc_op_residual_ln_e4m3", glc::residual_ln(x, r, gamma, beta, eps, y, M, H, (cudaStream_t)stream, nullptr, false, y8, y8_scale));
}
int glc_op_embed_ln(const int64_t* ids, const int64_t* mask, const void* emb, const float* gamma, const float* beta, float eps,
                    void* y, int M, int H, int vocab, void* stream) {
  GLC_TRY("glc_op_embed_ln", glc::embed_ln(ids, mask, emb, gamma, beta, eps, y, M, H, vocab, (cudaStream_t)stream));
}
int glc_op_residual_ln(const void* x, const void* r, const float* gamma, const float* beta, float eps, void* y, int M, int H,
                       void* stream) {
  GLC_TRY("glc_op_residual_ln", glc::residual_ln(x, r, gamma, beta, eps, y, M, H, (cudaStream_t)stream));
}
int glc_op_mask_prep(const int64_t* mask, uint32_t* bits, int32_t* kv_len, int B, int S, void* stream) {
  GLC_TRY("glc_op_mask_prep", glc::mask_prep(mask, bits, kv_len, B, S, (cudaStream_t)stream));
}
int glc_op_attention_naive(const void* qkv, const void* pos_k, const void* pos_q, int64_t ld_pos, const int32_t* rel_idx,
                           const uint32_t* mask_bits, void* ctx, int B, int S, int heads, int buckets, void* stream) {
  GLC_TRY("glc_op_attention_naive", glc::attention_naive(qkv, pos_k, pos_q, ld_pos, rel_idx, mask_bits, ctx, B, S, heads,
                                                          buckets, (cudaStream_t)stream));
}
int glc_op_expand_pos(const void* pos_f16, int64_t ld_src, int buckets, int max_pos, void* out_f16, int64_t ld_dst, int cols,
                      void* stream) {
  try {
    const int ER = glc::expanded_pos_rows();
    std::vector<int32_t> h(ER);
    glc::expanded_pos_index(buckets, max_pos, h.data());
    int32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, (size_t)ER * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), (size_t)ER * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = glc::expand_pos_table(pos_f16, ld_src, d, out_f16, ld_dst, cols, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (d) cudaFree(d);
    if (e != cudaSuccess) return fail(GLC_ERR_CUDA, std::string("glc_op_expand_pos: ") + cudaGetErrorString(e));
    return GLC_OK;
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_op_expand_pos: ") + e.what());
  }
}
int glc_op_expand_pos_rev(const void* pos_f16, int64_t ld_src, int buckets, int max_pos, void* out_f16, int64_t ld_dst, int cols,
                          void* stream) {
  try {
    const int ER = glc::expanded_pos_rows();
    std::vector<int32_t> h(ER);
    glc::expanded_pos_index_rev(buckets, max_pos, h.data());
    int32_t* d = nullptr;
    cudaError_t e = cudaMalloc(&d, (size_t)ER * 4);
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), (size_t)ER * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = glc::expand_pos_table(pos_f16, ld_src, d, out_f16, ld_dst, cols, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (d) cudaFree(d);
    if (e != cudaSuccess) return fail(GLC_ERR_CUDA, std::string("glc_op_expand_pos_rev: ") + cudaGetErrorString(e));
    return GLC_OK;
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_op_expand_pos_rev: ") + e.what());
  }
}
int glc_op_attention_persist(const void* qkv, const void* exp_k, const void* exp_qr, int64_t ld_exp, const uint32_t* mask_bits,
                             const int32_t* kv_len, void* ctx, int B, int S, int heads, void* stream) {
  GLC_TRY("glc_op_attention_persist", glc::attention_persist(qkv, exp_k, exp_qr, ld_exp, mask_bits, kv_len, ctx, B, S, heads,
                                                             num_sms_current(), (cudaStream_t)stream));
}
int glc_expanded_pos_rows(void) { return glc::expanded_pos_rows(); }
int glc_op_add_rmsnorm(float* h, const void* delta_f16, const float* g, float eps, void* y_f16, int M, int H, void* stream) {
  GLC_TRY("glc_op_add_rmsnorm", glc::add_rmsnorm(h, delta_f16, g, eps, y_f16, M, H, (cudaStream_t)stream));
}
int glc_op_rope(void* qkv_f16, int64_t ld, const float* inv_freq, int M, int S, int n_rot_heads, int head_dim, void* stream) {
  try {
    void* cs = nullptr;
    cudaError_t e = cudaMalloc(&cs, (size_t)S * (head_dim / 2) * 8);
    if (e == cudaSuccess) e = glc::rope_table(inv_freq, cs, S, head_dim, (cudaStream_t)stream);
    if (e == cudaSuccess) e = glc::rope_inplace(qkv_f16, ld, cs, M, S, n_rot_heads, head_dim, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (cs) cudaFree(cs);
    return wrap("glc_op_rope", e);
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_op_rope: ") + e.what());
  }
}
int glc_op_gemm_rope(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, void* C, int64_t ldc, int M, int N, int K,
                     const float* inv_freq, int S, int rope_cols, void* stream) {
  try {
    void* cs = nullptr;
    cudaError_t e = cudaMalloc(&cs, (size_t)S * 64 * 8);
    if (e == cudaSuccess) e = glc::rope_table(inv_freq, cs, S, 128, (cudaStream_t)stream);
    if (e == cudaSuccess) e = glc::gemm_f16_rope(A, lda, W, ldw, bias, C, ldc, M, N, K, cs, S, nullptr, rope_cols, num_sms_current(), (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (cs) cudaFree(cs);
    return wrap("glc_op_gemm_rope", e);
  } catch (const std::exception& e) {
    return fail(GLC_ERR_CUDA, std::string("glc_op_gemm_rope: ") + e.what());
  }
}
int glc_op_attention_flash128(const void* qkv_f16, const uint32_t* mask_bits, const int32_t* kv_len, void* ctx_f16, int B, int S,
                              int heads, int kv_heads, void* stream) {
  GLC_TRY("glc_op_attention_flash128",
          glc::attention_flash128(qkv_f16, mask_bits, kv_len, ctx_f16, B, S, heads, kv_heads, (cudaStream_t)stream));
}
int glc_op_head_gather(const void* h, const int64_t* ids, int64_t class_token, void* pooled, void* cls, int B, int S, int H,
                       int C, void* stream) {
  GLC_TRY("glc_op_head_gather", glc::head_gather(h, ids, class_token, pooled, cls, B, S, H, C, (cudaStream_t)stream));
}
int glc_op_head_score(const float* t, const float* k, float* logits, float* probs, uint8_t* decisions, float threshold, int B,
                      int C, int Hh, void* stream) {
  GLC_TRY("glc_op_head_score", glc::head_score(t, k, logits, probs, decisions, threshold, B, C, Hh, (cudaStream_t)stream));
}

}  // extern "C"
