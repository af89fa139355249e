// See engine.h.  Forward = what ORT executes per Run for the reference (src/model.c:173-182),
// as a fixed kernel sequence:
//   K1 embed+LN+mask -> L x { K2 QKV GEMM -> K3 fused disentangled attention -> K2 out-proj ->
//   K4 +res LN -> K2 FFN1(+GELU) -> K2 FFN2 -> K4 +res LN } -> K5a label pooling -> K2 projectors ->
//   K5b dot scorer.
#include "engine.h"

#include <cuda_fp16.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <thread>

#include "kernels.h"

namespace glc {

#define GLC_CUDA(expr)                                                                                  \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess)                                                                              \
      throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr);    \
  } while (0)

namespace {

const char* const kOverflowMsg =
    "fp16 activation overflow: a dense output before a LayerNorm reached +-65504 and was clamped, so the logits would be "
    "wrong; reload the model with GLC_PRELN_F32=1 (glc_opts.preln_f32 = 1) to keep the pre-LayerNorm sums in fp32";

// fp32 -> fp16, round to nearest even, saturating to +-65504 (weights of this family are O(1))
inline uint16_t f32_to_f16(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x > 0x7f800000u) return (uint16_t)(sign | 0x7e00u);          // NaN
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7bffu);         // >= 65520 (or inf) -> 65504
  if (x < 0x33000001u) return (uint16_t)sign;                      // < 2^-25 -> 0
  int e = (int)(x >> 23) - 127;
  uint32_t m = (x & 0x7fffffu) | 0x800000u;
  int shift = (e < -14) ? (13 + (-14 - e)) : 13;
  uint32_t half_m = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_m & 1u))) ++half_m;
  uint32_t h = (e < -14) ? half_m : (((uint32_t)(e + 15) << 10) + (half_m - 0x400u));
  return (uint16_t)(sign | h);
}

inline float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, u;
  if (e == 0) {
    if (m == 0) u = sign;
    else {
      int k = 0;
      while (!(m & 0x400u)) { m <<= 1; ++k; }
      u = sign | ((uint32_t)(127 - 15 - k + 1) << 23) | ((m & 0x3ffu) << 13);
    }
  } else if (e == 31) u = sign | 0x7f800000u | (m << 13);
  else u = sign | ((e + 112u) << 23) | (m << 13);
  float f;
  memcpy(&f, &u, 4);
  return f;
}

inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

}  // namespace

int usable_device_count() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

void* DeviceModel::dalloc(size_t bytes) {
  void* p = nullptr;
  GLC_CUDA(cudaMalloc(&p, bytes ? bytes : 16));
  return p;
}

void DeviceModel::upload_f32(float** dst, const HostTensor& t) {
  *dst = (float*)dalloc(t.data.size() * 4);
  perm_allocs_.push_back(*dst);
  GLC_CUDA(cudaMemcpyAsync(*dst, t.data.data(), t.data.size() * 4, cudaMemcpyHostToDevice, stream_));
}

void DeviceModel::upload_w16(void** dst, const float* src, size_t n) {
  std::vector<uint16_t> h(n);
  for (size_t i = 0; i < n; ++i) h[i] = f32_to_f16(src[i]);
  *dst = dalloc(n * 2);
  perm_allocs_.push_back(*dst);
  GLC_CUDA(cudaMemcpy(*dst, h.data(), n * 2, cudaMemcpyHostToDevice));
}

DeviceModel::DeviceModel(int device, const ModelWeights& w, int max_tokens, bool preln_f32, bool fp8_ffn) : device_(device), cfg_(w.cfg) {
  GLC_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  GLC_CUDA(cudaGetDeviceProperties(&prop, device_));
  if (prop.major != 10)
    throw std::runtime_error("device " + std::to_string(device_) + " (" + prop.name + ", sm_" + std::to_string(prop.major) +
                             std::to_string(prop.minor) + ") is not sm_100: this engine has no fallback path");
  num_sms_ = prop.multiProcessorCount;
  if (cfg_.backbone == BACKBONE_DEBERTA && cfg_.hidden / cfg_.heads != 64)
    throw std::runtime_error("attention kernel requires head dim 64 (got " + std::to_string(cfg_.hidden / cfg_.heads) + ")");
  if (cfg_.backbone == BACKBONE_QWEN2 && cfg_.head_dim != 128)
    throw std::runtime_error("decoder-backbone attention kernel requires head dim 128 (got " + std::to_string(cfg_.head_dim) + ")");
  if (cfg_.backbone == BACKBONE_QWEN2 && cfg_.pooling == POOL_LAST)
    throw std::runtime_error("pooling_strategy 'last' is not supported with a decoder backbone (padded rows are not computed)");
  if (max_tokens > 0) max_tokens_ = max_tokens;
  const char* dk = getenv("GLC_DEBUG_KEEP");
  debug_keep_ = dk && dk[0] == '1';
  graphs_on_ = getenv("GLC_NO_GRAPHS") == nullptr;
  // attention = attention_persist.cu (persistent CTAs, row-owner warpgroups rotating over key tiles, position tables
  // resident in shared memory for S <= 512); the generations it supersedes live under experiments/attention_generations/
  {
    const char* fr = getenv("GLC_FUSE_RESID");
    fuse_resid_ = fr && fr[0] == '1';   // measured: LN -0.1 ms, but the out-proj / FFN2 epilogues +0.28 ms per step -> off
  }
  {
    const char* pf = getenv("GLC_PRELN_F32");
    preln_f32_ = preln_f32 || (pf && pf[0] == '1');
    if (preln_f32_) fuse_resid_ = false;
  }
  {
    const char* vl = getenv("GLC_VARLEN");
    varlen_ = !(vl && vl[0] == '0');
  }
  {
    const char* f8 = getenv("GLC_FP8_FFN");
    fp8_ffn_ = fp8_ffn || (f8 && f8[0] == '1');
    if (const char* fm = getenv("GLC_FP8_MULT")) fp8_mult_ = (float)atof(fm);
    if (fp8_ffn_) {
      if (preln_f32_) throw std::runtime_error("FP8 FFN weights and preln_f32 cannot be combined");
      if (cfg_.hidden % 16 || cfg_.inter % 16) throw std::runtime_error("FP8 FFN weights need hidden and intermediate sizes divisible by 16");
      if (!(fp8_mult_ > 0.f)) throw std::runtime_error("GLC_FP8_MULT must be positive");
      fuse_resid_ = false;
    }
  }
  GLC_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
  d_overflow_ = (int*)dalloc(sizeof(int));
  perm_allocs_.push_back(d_overflow_);
  GLC_CUDA(cudaMemset(d_overflow_, 0, sizeof(int)));
  GLC_CUDA(cudaMallocHost((void**)&h_overflow_, 64 * sizeof(int)));
  memset(h_overflow_, 0, 64 * sizeof(int));

  if (cfg_.backbone == BACKBONE_QWEN2) init_qwen2(w);
  else init_deberta(w);
  init_head(w);
  GLC_CUDA(cudaStreamSynchronize(stream_));
}

// decoder backbone (Qwen2): fused QKV (q heads | kv heads K | kv heads V), gate / up rows interleaved in blocks of 32 for
// the SwiGLU GEMM epilogue, RMSNorm weights, rotary inverse frequencies
void DeviceModel::init_qwen2(const ModelWeights& w) {
  const int H = cfg_.hidden, I = cfg_.inter, d = cfg_.head_dim;
  const int Wq = cfg_.heads * d, Wkv = cfg_.kv_heads * d, Wqkv = Wq + 2 * Wkv;
  if (I % 32) throw std::runtime_error("decoder backbone: intermediate size must be a multiple of 32");
  const HostTensor& we = w.at("emb.word");
  upload_w16(&word_emb_, we.data.data(), we.data.size());
  upload_f32(&norm_g_, w.at("norm.g"));
  upload_f32(&rope_inv_freq_, w.at("rope.inv_freq"));
  layers_.resize(cfg_.layers);
  std::vector<float> cat((size_t)Wqkv * H), bcat((size_t)Wqkv), gu((size_t)2 * I * H);
  for (int l = 0; l < cfg_.layers; ++l) {
    DeviceLayer& dl = layers_[l];
    const std::string r = "layer." + std::to_string(l);
    memcpy(cat.data(), w.at(r + ".q.w").data.data(), (size_t)Wq * H * 4);
    memcpy(cat.data() + (size_t)Wq * H, w.at(r + ".k.w").data.data(), (size_t)Wkv * H * 4);
    memcpy(cat.data() + (size_t)(Wq + Wkv) * H, w.at(r + ".v.w").data.data(), (size_t)Wkv * H * 4);
    memcpy(bcat.data(), w.at(r + ".q.b").data.data(), (size_t)Wq * 4);
    memcpy(bcat.data() + Wq, w.at(r + ".k.b").data.data(), (size_t)Wkv * 4);
    memcpy(bcat.data() + Wq + Wkv, w.at(r + ".v.b").data.data(), (size_t)Wkv * 4);
    upload_w16(&dl.wqkv, cat.data(), cat.size());
    HostTensor hb;
    hb.data = bcat;
    upload_f32(&dl.bqkv, hb);
    GLC_CUDA(cudaStreamSynchronize(stream_));   // hb is a temporary
    upload_w16(&dl.wo, w.at(r + ".o.w").data.data(), (size_t)H * Wq);
    const float* gw = w.at(r + ".gate.w").data.data();
    const float* uw = w.at(r + ".up.w").data.data();
    for (int j = 0; j < I / 32; ++j) {
      memcpy(gu.data() + (size_t)(64 * j) * H, gw + (size_t)(32 * j) * H, (size_t)32 * H * 4);
      memcpy(gu.data() + (size_t)(64 * j + 32) * H, uw + (size_t)(32 * j) * H, (size_t)32 * H * 4);
    }
    upload_w16(&dl.w1, gu.data(), gu.size());
    upload_w16(&dl.w2, w.at(r + ".down.w").data.data(), (size_t)H * I);
    if (fp8_ffn_) {
      // load-time quantiser (see init_deberta): the interleaved gate|up matrix and the down projection, one scale per row
      dl.w1_8 = dalloc((size_t)2 * I * H); perm_allocs_.push_back(dl.w1_8);
      dl.w2_8 = dalloc((size_t)H * I); perm_allocs_.push_back(dl.w2_8);
      dl.w1_s = (float*)dalloc((size_t)2 * I * 4); perm_allocs_.push_back(dl.w1_s);
      dl.w2_s = (float*)dalloc((size_t)H * 4); perm_allocs_.push_back(dl.w2_s);
      GLC_CUDA(quantize_rows_e4m3(dl.w1, H, dl.w1_8, H, dl.w1_s, 2 * I, H, stream_));
      GLC_CUDA(quantize_rows_e4m3(dl.w2, I, dl.w2_8, I, dl.w2_s, H, I, stream_));
    }
    upload_f32(&dl.ln1g, w.at(r + ".ln1.g"));
    upload_f32(&dl.ln2g, w.at(r + ".ln2.g"));
  }
}

void DeviceModel::init_deberta(const ModelWeights& w) {
  const int H = cfg_.hidden, I = cfg_.inter, R = 2 * cfg_.buckets;
  const HostTensor& we = w.at("emb.word");
  upload_w16(&word_emb_, we.data.data(), we.data.size());
  upload_f32(&emb_g_, w.at("emb.ln.g"));
  upload_f32(&emb_b_, w.at("emb.ln.b"));

  // LN(rel_embeddings) once (T:597-601: input independent)
  float *rel_f32 = nullptr, *rel_g = nullptr, *rel_b = nullptr;
  upload_f32(&rel_f32, w.at("rel.emb"));
  upload_f32(&rel_g, w.at("rel.ln.g"));
  upload_f32(&rel_b, w.at("rel.ln.b"));
  void* rel_ln = dalloc((size_t)R * H * 2);
  perm_allocs_.push_back(rel_ln);
  GLC_CUDA(ln_f32_to_f16(rel_f32, rel_g, rel_b, cfg_.ln_eps, rel_ln, R, H, stream_));
  ++launches_;

  // index of the delta-expanded position tables: posK half row rho <- pos_qk[idx(2047 - rho)], posQ half row sigma <- idx(sigma - 2047)
  const int ER = expanded_pos_rows();
  int32_t *d_exp_idx = nullptr, *d_exp_idx_rev = nullptr;
  {
    std::vector<int32_t> h(ER);
    expanded_pos_index(cfg_.buckets, cfg_.max_rel_pos, h.data());
    d_exp_idx = (int32_t*)dalloc((size_t)ER * 4);
    perm_allocs_.push_back(d_exp_idx);
    GLC_CUDA(cudaMemcpy(d_exp_idx, h.data(), (size_t)ER * 4, cudaMemcpyHostToDevice));
    expanded_pos_index_rev(cfg_.buckets, cfg_.max_rel_pos, h.data());
    d_exp_idx_rev = (int32_t*)dalloc((size_t)ER * 4);
    perm_allocs_.push_back(d_exp_idx_rev);
    GLC_CUDA(cudaMemcpy(d_exp_idx_rev, h.data(), (size_t)ER * 4, cudaMemcpyHostToDevice));
  }

  layers_.resize(cfg_.layers);
  std::vector<float> cat((size_t)3 * H * H), bcat((size_t)3 * H);
  for (int l = 0; l < cfg_.layers; ++l) {
    DeviceLayer& d = layers_[l];
    const std::string r = "layer." + std::to_string(l);
    const char* names[3] = {".q", ".k", ".v"};
    for (int j = 0; j < 3; ++j) {
      memcpy(cat.data() + (size_t)j * H * H, w.at(r + names[j] + ".w").data.data(), (size_t)H * H * 4);
      memcpy(bcat.data() + (size_t)j * H, w.at(r + names[j] + ".b").data.data(), (size_t)H * 4);
    }
    upload_w16(&d.wqkv, cat.data(), cat.size());
    HostTensor hb;
    hb.data = bcat;
    upload_f32(&d.bqkv, hb);
    GLC_CUDA(cudaStreamSynchronize(stream_));   // hb is a temporary
    upload_w16(&d.wo, w.at(r + ".o.w").data.data(), (size_t)H * H);
    upload_f32(&d.bo, w.at(r + ".o.b"));
    upload_f32(&d.ln1g, w.at(r + ".ln1.g"));
    upload_f32(&d.ln1b, w.at(r + ".ln1.b"));
    upload_w16(&d.w1, w.at(r + ".ffn1.w").data.data(), (size_t)I * H);
    upload_f32(&d.b1, w.at(r + ".ffn1.b"));
    upload_w16(&d.w2, w.at(r + ".ffn2.w").data.data(), (size_t)H * I);
    upload_f32(&d.b2, w.at(r + ".ffn2.b"));
    if (fp8_ffn_) {
      // load-time quantiser: per-output-channel (row of W) symmetric e4m3, scale = amax / 448
      d.w1_8 = dalloc((size_t)I * H); perm_allocs_.push_back(d.w1_8);
      d.w2_8 = dalloc((size_t)H * I); perm_allocs_.push_back(d.w2_8);
      d.w1_s = (float*)dalloc((size_t)I * 4); perm_allocs_.push_back(d.w1_s);
      d.w2_s = (float*)dalloc((size_t)H * 4); perm_allocs_.push_back(d.w2_s);
      GLC_CUDA(quantize_rows_e4m3(d.w1, H, d.w1_8, H, d.w1_s, I, H, stream_));
      GLC_CUDA(quantize_rows_e4m3(d.w2, I, d.w2_8, I, d.w2_s, H, I, stream_));
    }
    upload_f32(&d.ln2g, w.at(r + ".ln2.g"));
    upload_f32(&d.ln2b, w.at(r + ".ln2.b"));
    // position projections with the shared content weights (T:296-302, share_att_key), hoisted
    // out of the per-Run graph: pos_qk[:, 0:H] = query_proj(rel), pos_qk[:, H:2H] = key_proj(rel)
    d.pos_qk = dalloc((size_t)R * 2 * H * 2);
    perm_allocs_.push_back(d.pos_qk);
    GLC_CUDA(gemm_f16(rel_ln, H, d.wqkv, H, d.bqkv, d.pos_qk, 2 * H, R, 2 * H, H, 0, false, num_sms_, stream_));
    ++launches_;
    // posQ half in sigma order, posK half in rho order (one row per relative distance)
    d.pos_exp = dalloc((size_t)ER * 2 * H * 2);
    perm_allocs_.push_back(d.pos_exp);
    GLC_CUDA(expand_pos_table(d.pos_qk, 2 * H, d_exp_idx_rev, d.pos_exp, 2 * H, H, stream_));
    GLC_CUDA(expand_pos_table((const __half*)d.pos_qk + H, 2 * H, d_exp_idx, (__half*)d.pos_exp + H, 2 * H, H, stream_));
    launches_ += 2;
  }
}

void DeviceModel::init_head(const ModelWeights& w) {
  const int H = cfg_.hidden, Hh = cfg_.head_hidden;
  upload_w16(&t1w_, w.at("text.1.w").data.data(), (size_t)Hh * H);
  upload_f32(&t1b_, w.at("text.1.b"));
  upload_w16(&t2w_, w.at("text.2.w").data.data(), (size_t)Hh * Hh);
  upload_f32(&t2b_, w.at("text.2.b"));
  upload_w16(&c1w_, w.at("cls.1.w").data.data(), (size_t)Hh * H);
  upload_f32(&c1b_, w.at("cls.1.b"));
  upload_w16(&c2w_, w.at("cls.2.w").data.data(), (size_t)Hh * Hh);
  upload_f32(&c2b_, w.at("cls.2.b"));
  if (cfg_.scorer == SCORER_MLP) {
    upload_w16(&m1w_, w.at("scorer.mlp.0.w").data.data(), w.at("scorer.mlp.0.w").data.size());
    upload_f32(&m1b_, w.at("scorer.mlp.0.b"));
    upload_w16(&m2w_, w.at("scorer.mlp.2.w").data.data(), w.at("scorer.mlp.2.w").data.size());
    upload_f32(&m2b_, w.at("scorer.mlp.2.b"));
    upload_f32(&last_w_, w.at("scorer.mlp.4.w"));
    last_b_ = w.at("scorer.mlp.4.b").data.at(0);
  } else if (cfg_.scorer == SCORER_WEIGHTED_DOT) {
    upload_w16(&ptw_, w.at("scorer.pt.w").data.data(), w.at("scorer.pt.w").data.size());
    upload_f32(&ptb_, w.at("scorer.pt.b"));
    upload_w16(&plw_, w.at("scorer.pl.w").data.data(), w.at("scorer.pl.w").data.size());
    upload_f32(&plb_, w.at("scorer.pl.b"));
    upload_w16(&o1w_, w.at("scorer.o1.w").data.data(), w.at("scorer.o1.w").data.size());
    upload_f32(&o1b_, w.at("scorer.o1.b"));
    upload_f32(&last_w_, w.at("scorer.o2.w"));
    last_b_ = w.at("scorer.o2.b").data.at(0);
  }
}

void DeviceModel::drop_graphs() {
  for (auto& kv : graphs_)
    if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
  graphs_live_ = 0;
}

DeviceModel::~DeviceModel() {
  cudaSetDevice(device_);
  if (stream_) cudaStreamSynchronize(stream_);
  drop_graphs();
  for (void* p : ws_allocs_) cudaFree(p);
  for (void* p : perm_allocs_) cudaFree(p);
  if (h_ids_) { cudaFreeHost(h_ids_); cudaFreeHost(h_mask_); }
  if (h_logits_) { cudaFreeHost(h_logits_); cudaFreeHost(h_probs_); cudaFreeHost(h_dec_); }
  if (h_overflow_) cudaFreeHost(h_overflow_);
  for (auto& b : h_in_) if (b.p) cudaFreeHost(b.p);
  for (auto e : h_in_ev_) if (e) cudaEventDestroy(e);
  for (auto& b : out_pool_) if (b.p) cudaFreeHost(b.p);
  for (auto& kv : rel_tables_) cudaFree(kv.second);
  for (auto& kv : rope_tables_) cudaFree(kv.second);
  for (auto& kv : debug_) cudaFree(kv.second.ptr);
  for (auto& r : prof_recs_) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto e : prof_pool_) cudaEventDestroy(e);
  for (auto e : free_events_) cudaEventDestroy(e);
  if (stream_) cudaStreamDestroy(stream_);
}

const int32_t* DeviceModel::rel_table(int S) {
  const int Spad = round_up(S, 128);
  auto it = rel_tables_.find(Spad);
  if (it != rel_tables_.end()) return it->second;
  std::vector<int32_t> h((size_t)2 * Spad - 1);
  rel_index_table(Spad, cfg_.buckets, cfg_.max_rel_pos, h.data());
  int32_t* d = (int32_t*)dalloc(h.size() * 4);
  GLC_CUDA(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  rel_tables_[Spad] = d;
  return d;
}

const void* DeviceModel::rope_table_for(int S) {
  auto it = rope_tables_.find(S);
  if (it != rope_tables_.end()) return it->second;
  void* d = dalloc((size_t)S * (cfg_.head_dim / 2) * 8);
  GLC_CUDA(rope_table(rope_inv_freq_, d, S, cfg_.head_dim, stream_));
  GLC_CUDA(cudaStreamSynchronize(stream_));   // (first sight of a shape runs eagerly, never inside a graph capture)
  rope_tables_[S] = d;
  return d;
}

void DeviceModel::ensure_workspace(int tokens, int B, int C) {
  const int rows = B * (C > 0 ? C : 1);
  if (tokens <= ws_tokens_ && B <= ws_B_ && rows <= ws_rows_) return;
  GLC_CUDA(cudaStreamSynchronize(stream_));
  drop_graphs();   // they captured the old workspace pointers
  for (void* p : ws_allocs_) cudaFree(p);
  ws_allocs_.clear();
  ws_tokens_ = tokens > ws_tokens_ ? tokens : ws_tokens_;
  ws_B_ = B > ws_B_ ? B : ws_B_;
  ws_rows_ = rows > ws_rows_ ? rows : ws_rows_;
  const size_t M = (size_t)ws_tokens_, H = cfg_.hidden, I = cfg_.inter, Hh = cfg_.head_hidden;
  const bool dec = cfg_.backbone == BACKBONE_QWEN2;
  const size_t Wq = dec ? (size_t)cfg_.heads * cfg_.head_dim : H, Wqkv = dec ? Wq + 2 * (size_t)cfg_.kv_heads * cfg_.head_dim : 3 * H;
  auto A = [&](size_t bytes) { void* p = dalloc(bytes); ws_allocs_.push_back(p); return p; };
  ids_ = (int64_t*)A(M * 8);
  mask_ = (int64_t*)A(M * 8);
  x_ = A(M * H * 2);
  x1_ = dec ? nullptr : A(M * H * 2);
  h32_ = dec ? (float*)A(M * H * 4) : nullptr;   // decoder backbone: fp32 residual stream
  qkv_ = A(M * Wqkv * 2);
  ctx_ = A(M * Wq * 2);
  tmp_ = A(M * H * (preln_f32_ ? 4 : 2));
  ffn_ = A(M * I * 2);   // fp16 [M, I]; the FP8 path stores e4m3 [M, I] in its first half
  if (fp8_ffn_) {
    x1_8_ = A(M * H);
    x1_s_ = (float*)A(M * 4);
  }
  mask_bits_ = (uint32_t*)A(((M + 31) / 32 + (size_t)ws_B_) * 4);
  kv_len_ = (int32_t*)A((size_t)ws_B_ * 4);
  pk_ints_ = (int32_t*)A(((size_t)2 * ws_B_ + 2 + 2 * (M / 128 + 1)) * 4);
  pk_scratch_ = (int32_t*)A((M / 128 + 1) * 4);
  pooled_ = A((size_t)ws_B_ * H * 2);
  cls_ = A((size_t)ws_rows_ * H * 2);
  tmid_ = A((size_t)ws_B_ * Hh * 2);
  cmid_ = A((size_t)ws_rows_ * Hh * 2);
  tvec_ = (float*)A((size_t)ws_B_ * Hh * 4);
  kvec_ = (float*)A((size_t)ws_rows_ * Hh * 4);
  logits_ = (float*)A((size_t)ws_rows_ * 4);
  probs_ = (float*)A((size_t)ws_rows_ * 4);
  decisions_ = (uint8_t*)A((size_t)ws_rows_);
  if (cfg_.scorer == SCORER_MLP) {
    cat16_ = A((size_t)ws_rows_ * 2 * Hh * 2);
    s1_ = A((size_t)ws_rows_ * cfg_.mlp1 * 2);
    s2_ = (float*)A((size_t)ws_rows_ * cfg_.mlp2 * 4);
  } else if (cfg_.scorer == SCORER_WEIGHTED_DOT) {
    t16_ = A((size_t)ws_B_ * Hh * 2);
    k16_ = A((size_t)ws_rows_ * Hh * 2);
    pt_ = (float*)A((size_t)ws_B_ * 2 * Hh * 4);
    pl_ = (float*)A((size_t)ws_rows_ * 2 * Hh * 4);
    cat16_ = A((size_t)ws_rows_ * 3 * Hh * 2);
    s2_ = (float*)A((size_t)ws_rows_ * 4 * Hh * 4);
  }
}

void DeviceModel::keep(const char* name, const void* src, size_t count) {
  if (!debug_keep_) return;
  DebugBuf& d = debug_[name];
  if (d.count < count) {
    if (d.ptr) cudaFree(d.ptr);
    d.ptr = dalloc(count * 2);
  }
  d.count = count;
  GLC_CUDA(cudaMemcpyAsync(d.ptr, src, count * 2, cudaMemcpyDeviceToDevice, stream_));
}

int64_t DeviceModel::debug_fetch(const std::string& name, float* out, size_t capacity) {
  std::lock_guard<std::mutex> lk(mu);
  auto it = debug_.find(name);
  if (it == debug_.end()) return -1;
  const size_t n = it->second.count;
  if (n > capacity) return -(int64_t)n;
  cudaSetDevice(device_);
  std::vector<uint16_t> h(n);
  GLC_CUDA(cudaStreamSynchronize(stream_));
  GLC_CUDA(cudaMemcpy(h.data(), it->second.ptr, n * 2, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < n; ++i) {
    out[i] = f16_to_f32(h[i]);
  }
  return (int64_t)n;
}

cudaEvent_t DeviceModel::prof_event() {
  if (!prof_pool_.empty()) {
    cudaEvent_t e = prof_pool_.back();
    prof_pool_.pop_back();
    return e;
  }
  cudaEvent_t e;
  GLC_CUDA(cudaEventCreate(&e));
  return e;
}

void DeviceModel::profile_enable(bool on) {
  std::lock_guard<std::mutex> lk(mu);
  prof_on_ = on;
}

void DeviceModel::profile_collect(double* ms, uint64_t* n) {
  std::lock_guard<std::mutex> lk(mu);
  cudaSetDevice(device_);
  GLC_CUDA(cudaStreamSynchronize(stream_));
  for (auto& r : prof_recs_) {
    float t = 0.f;
    GLC_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms[r.cat] += t;
    n[r.cat] += 1;
    prof_pool_.push_back(r.a);
    prof_pool_.push_back(r.b);
  }
  prof_recs_.clear();
}

// brackets one launch with events when profiling is on
struct ProfScope {
  DeviceModel* d;
  int cat;
  cudaEvent_t a = nullptr;
  ProfScope(DeviceModel* dm, int c) : d(dm), cat(c) {
    if (d->prof_on_) {
      a = d->prof_event();
      cudaEventRecord(a, d->stream_);
    }
  }
  ~ProfScope() {
    if (a) {
      cudaEvent_t b = d->prof_event();
      cudaEventRecord(b, d->stream_);
      d->prof_recs_.push_back({cat, a, b});
    }
  }
};
#define GLC_LAUNCH(cat, expr)      \
  do {                             \
    ProfScope _ps(this, cat);      \
    GLC_CUDA(expr);                \
    ++n;                           \
  } while (0)

void DeviceModel::forward(const int64_t* d_ids, const int64_t* d_mask, int B, int S, int C, float* d_logits, float* d_probs,
                          uint8_t* d_decisions, float threshold, bool cacheable) {
  if (B * S <= 0) return;
  if (d_ids != ids_) ensure_workspace(B * S, B, C);   // run_host already sized it
  if (!graphs_on_ || prof_on_ || debug_keep_ || !cacheable) {
    forward_eager(d_ids, d_mask, B, S, C, d_logits, d_probs, d_decisions, threshold);
    return;
  }
  uint32_t thr;
  memcpy(&thr, &threshold, 4);
  const GraphKey key{B, S, C, d_ids, d_mask, d_logits, d_probs, d_decisions, thr};
  if (graphs_.size() >= (size_t)kMaxGraphKeys && !graphs_.count(key)) {
    // unbounded variety of shapes (pad-to-longest batches): forget the placeholders of shapes seen only once
    for (auto it = graphs_.begin(); it != graphs_.end();)
      it = it->second.exec ? std::next(it) : graphs_.erase(it);
  }
  GraphEntry& e = graphs_[key];
  e.last_use = ++graph_clock_;
  if (e.exec) {
    GLC_CUDA(cudaGraphLaunch(e.exec, stream_));
    launches_ += e.launches;
    return;
  }
  if (e.seen++ == 0) {   // first sight of this shape: run eagerly (one-time attribute / table set-up is not capturable)
    forward_eager(d_ids, d_mask, B, S, C, d_logits, d_probs, d_decisions, threshold);
    return;
  }
  if (graphs_live_ >= kMaxGraphs) {   // evict the least recently replayed graph
    auto victim = graphs_.end();
    for (auto it = graphs_.begin(); it != graphs_.end(); ++it)
      if (it->second.exec && (victim == graphs_.end() || it->second.last_use < victim->second.last_use)) victim = it;
    if (victim != graphs_.end()) {
      GLC_CUDA(cudaStreamSynchronize(stream_));   // the victim may still be executing
      cudaGraphExecDestroy(victim->second.exec);
      graphs_.erase(victim);
      --graphs_live_;
    }
  }
  const uint64_t before = launches_.load();
  GLC_CUDA(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
  cudaGraph_t g = nullptr;
  try {
    forward_eager(d_ids, d_mask, B, S, C, d_logits, d_probs, d_decisions, threshold);
  } catch (...) {
    cudaStreamEndCapture(stream_, &g);
    if (g) cudaGraphDestroy(g);
    throw;
  }
  GLC_CUDA(cudaStreamEndCapture(stream_, &g));
  cudaGraphExec_t exec = nullptr;
  cudaError_t ie = cudaGraphInstantiate(&exec, g, 0);
  cudaGraphDestroy(g);
  if (ie != cudaSuccess) throw std::runtime_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ie));
  GraphEntry& e2 = graphs_[key];   // (the eviction above may have rebalanced the map: look the entry up again)
  e2.launches = launches_.load() - before;
  e2.exec = exec;
  e2.last_use = graph_clock_;
  ++graphs_live_;
  GLC_CUDA(cudaGraphLaunch(e2.exec, stream_));
}

void DeviceModel::forward_eager(const int64_t* d_ids, const int64_t* d_mask, int B, int S, int C, float* d_logits,
                                float* d_probs, uint8_t* d_decisions, float threshold) {
  const int H = cfg_.hidden, I = cfg_.inter, Hh = cfg_.head_hidden;
  const PackedCtx* pk = pk_;   // packed (varlen) micro-batch: d_ids / d_mask are flat [rows], see engine.h
  const int M = pk ? pk->rows : B * S;
  cudaStream_t st = stream_;
  uint64_t n = 0;

  if (pk) {
    // one validity bit per packed row (rows / 128 pseudo-rows of 128); the per-text key lengths came from the host
    GLC_LAUNCH(KC_EMBED, mask_prep(d_mask, mask_bits_, pk_scratch_, pk->n_tiles, 128, st));
  } else {
    GLC_LAUNCH(KC_EMBED, mask_prep(d_mask, mask_bits_, kv_len_, B, S, st));
  }
  if (cfg_.backbone == BACKBONE_QWEN2) {
    // decoder backbone (transformers modeling_qwen2.py Q:353-410, bidirectional): pre-norm residual stream in fp32
    const int d = cfg_.head_dim, nh = cfg_.heads, nkv = cfg_.kv_heads;
    const int Wq = nh * d, Wqkv = Wq + 2 * nkv * d;
    const void* cs = rope_table_for(pk ? round_up(S, 128) : S);   // packed texts own whole 128-row tiles
    GLC_LAUNCH(KC_EMBED, embed_rows_f32(d_ids, word_emb_, h32_, M, H, cfg_.vocab, st));
    for (int l = 0; l < cfg_.layers; ++l) {
      const DeviceLayer& dl = layers_[l];
      GLC_LAUNCH(KC_LN, add_rmsnorm(h32_, l == 0 ? nullptr : tmp_, dl.ln1g, cfg_.rms_eps, x_, M, H, st));
      if (d == 128) {
        // rotary embedding in the GEMM epilogue (fp32 accumulators, one rounding): no separate pass over the q | k slab
        GLC_LAUNCH(KC_GEMM_QKV, gemm_f16_rope(x_, H, dl.wqkv, H, dl.bqkv, qkv_, Wqkv, M, Wqkv, H, cs, pk ? round_up(S, 128) : S,
                                              pk ? pk->tile_pos : nullptr, (nh + nkv) * d, num_sms_, st));
      } else {
        GLC_LAUNCH(KC_GEMM_QKV, gemm_f16(x_, H, dl.wqkv, H, dl.bqkv, qkv_, Wqkv, M, Wqkv, H, 0, false, num_sms_, st));
        GLC_LAUNCH(KC_EMBED, rope_inplace(qkv_, Wqkv, cs, M, S, nh + nkv, d, st, pk ? pk->tile_pos : nullptr));
      }
      if (pk) {
        GLC_LAUNCH(KC_ATTN, attention_flash128_packed(qkv_, mask_bits_, pk->kv_len, pk->text_row, pk->tile_info, ctx_, B, pk->rows,
                                                      pk->n_tiles, nh, nkv, st));
      } else {
        GLC_LAUNCH(KC_ATTN, attention_flash128(qkv_, mask_bits_, kv_len_, ctx_, B, S, nh, nkv, st));
      }
      if (l == 0) keep("ctx0", ctx_, (size_t)M * Wq);
      GLC_LAUNCH(KC_GEMM_OUT, gemm_f16(ctx_, Wq, dl.wo, Wq, nullptr, tmp_, H, M, H, Wq, 0, false, num_sms_, st));
      GLC_LAUNCH(KC_LN, add_rmsnorm(h32_, tmp_, dl.ln2g, cfg_.rms_eps, x_, M, H, st, fp8_ffn_ ? x1_8_ : nullptr,
                                    fp8_ffn_ ? x1_s_ : nullptr));
      if (fp8_ffn_) {
        // e4m3 MLP: the post-attention RMSNorm wrote the rows as e4m3 under per-row scales; silu(gate) * up stays e4m3
        GLC_LAUNCH(KC_GEMM_FFN1, gemm_e4m3(x1_8_, H, dl.w1_8, H, x1_s_, 1.0f, dl.w1_s, nullptr, ffn_, I, M, 2 * I, H, 3, true,
                                           fp8_mult_, num_sms_, st));
        GLC_LAUNCH(KC_GEMM_FFN2, gemm_e4m3(ffn_, I, dl.w2_8, I, nullptr, 1.0f / fp8_mult_, dl.w2_s, nullptr, tmp_, H, M, H, I, 0,
                                           false, 1.0f, num_sms_, st));
        continue;
      }
      GLC_LAUNCH(KC_GEMM_FFN1, gemm_f16(x_, H, dl.w1, H, nullptr, ffn_, I, M, 2 * I, H, 3, false, num_sms_, st));
      GLC_LAUNCH(KC_GEMM_FFN2, gemm_f16(ffn_, I, dl.w2, I, nullptr, tmp_, H, M, H, I, 0, false, num_sms_, st));
    }
    GLC_LAUNCH(KC_LN, add_rmsnorm(h32_, tmp_, norm_g_, cfg_.rms_eps, x_, M, H, st));
    keep("final", x_, (size_t)M * H);
  } else {
  GLC_LAUNCH(KC_EMBED, embed_ln(d_ids, d_mask, word_emb_, emb_g_, emb_b_, cfg_.ln_eps, x_, M, H, cfg_.vocab, st));
  keep("emb", x_, (size_t)M * H);
  for (int l = 0; l < cfg_.layers; ++l) {
    const DeviceLayer& d = layers_[l];
    GLC_LAUNCH(KC_GEMM_QKV, gemm_f16(x_, H, d.wqkv, H, d.bqkv, qkv_, 3 * H, M, 3 * H, H, 0, false, num_sms_, st));
    if (l == 0) keep("qkv0", qkv_, (size_t)M * 3 * H);
    const __half* pe = (const __half*)d.pos_exp;
    if (pk) {
      GLC_LAUNCH(KC_ATTN, attention_persist_packed(qkv_, pe + H, pe, 2 * H, mask_bits_, pk->kv_len, pk->text_row, pk->tile_info, ctx_,
                                                   B, pk->rows, pk->max_rows, pk->n_tiles, cfg_.heads, num_sms_, st));
    } else {
      GLC_LAUNCH(KC_ATTN, attention_persist(qkv_, pe + H, pe, 2 * H, mask_bits_, kv_len_, ctx_, B, S, cfg_.heads, num_sms_, st));
    }
    if (cfg_.pooling == POOL_LAST) GLC_LAUNCH(KC_ATTN, pad_rows_mean_v(qkv_, d_mask, ctx_, B, S, H, st));
    if (l == 0) keep("ctx0", ctx_, (size_t)M * H);
    // GLC_FUSE_RESID=1: the residual add rides in the GEMM epilogue and LN reads one tensor (slower in total, see engine ctor)
    if (fuse_resid_) {
      GLC_LAUNCH(KC_GEMM_OUT, gemm_f16_resid(ctx_, H, d.wo, H, d.bo, x_, H, tmp_, H, M, H, H, 0, false, num_sms_, st));
      GLC_LAUNCH(KC_LN, residual_ln(tmp_, nullptr, d.ln1g, d.ln1b, cfg_.ln_eps, x1_, M, H, st, d_overflow_));
    } else {
      GLC_LAUNCH(KC_GEMM_OUT, gemm_f16(ctx_, H, d.wo, H, d.bo, tmp_, H, M, H, H, 0, preln_f32_, num_sms_, st));
      GLC_LAUNCH(KC_LN, residual_ln(tmp_, x_, d.ln1g, d.ln1b, cfg_.ln_eps, x1_, M, H, st, d_overflow_, preln_f32_,
                                    fp8_ffn_ ? x1_8_ : nullptr, fp8_ffn_ ? x1_s_ : nullptr));
    }
    if (fp8_ffn_) {
      // e4m3 FFN: LN1 wrote the rows as e4m3 under per-row scales; GELU output stays e4m3 (x fp8_mult_) for FFN2
      GLC_LAUNCH(KC_GEMM_FFN1, gemm_e4m3(x1_8_, H, d.w1_8, H, x1_s_, 1.0f, d.w1_s, d.b1, ffn_, I, M, I, H, 1, true, fp8_mult_,
                                         num_sms_, st));
      GLC_LAUNCH(KC_GEMM_FFN2, gemm_e4m3(ffn_, I, d.w2_8, I, nullptr, 1.0f / fp8_mult_, d.w2_s, d.b2, tmp_, H, M, H, I, 0, false,
                                         1.0f, num_sms_, st));
      GLC_LAUNCH(KC_LN, residual_ln(tmp_, x1_, d.ln2g, d.ln2b, cfg_.ln_eps, x_, M, H, st, d_overflow_));
      if (debug_keep_) keep(("h" + std::to_string(l)).c_str(), x_, (size_t)M * H);
      continue;
    }
    GLC_LAUNCH(KC_GEMM_FFN1, gemm_f16(x1_, H, d.w1, H, d.b1, ffn_, I, M, I, H, 1, false, num_sms_, st));
    if (fuse_resid_) {
      GLC_LAUNCH(KC_GEMM_FFN2, gemm_f16_resid(ffn_, I, d.w2, I, d.b2, x1_, H, tmp_, H, M, H, I, 0, false, num_sms_, st));
      GLC_LAUNCH(KC_LN, residual_ln(tmp_, nullptr, d.ln2g, d.ln2b, cfg_.ln_eps, x_, M, H, st, d_overflow_));
    } else {
      GLC_LAUNCH(KC_GEMM_FFN2, gemm_f16(ffn_, I, d.w2, I, d.b2, tmp_, H, M, H, I, 0, preln_f32_, num_sms_, st));
      GLC_LAUNCH(KC_LN, residual_ln(tmp_, x1_, d.ln2g, d.ln2b, cfg_.ln_eps, x_, M, H, st, d_overflow_, preln_f32_));
    }
    if (debug_keep_) keep(("h" + std::to_string(l)).c_str(), x_, (size_t)M * H);
  }
  }
  if (C > 0) {
    const int pa = cfg_.proj_act;
    GLC_LAUNCH(KC_HEAD_MISC, head_gather_pool(x_, d_ids, d_mask, cfg_.class_token, cfg_.pooling, pooled_, cls_, B, S, H, C, st,
                                              cfg_.class_pos_offset, pk ? pk->text_row : nullptr));
    GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(pooled_, H, t1w_, H, t1b_, tmid_, Hh, B, Hh, H, pa, false, num_sms_, st));
    GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(tmid_, Hh, t2w_, Hh, t2b_, tvec_, Hh, B, Hh, Hh, 0, true, num_sms_, st));
    GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(cls_, H, c1w_, H, c1b_, cmid_, Hh, B * C, Hh, H, pa, false, num_sms_, st));
    GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(cmid_, Hh, c2w_, Hh, c2b_, kvec_, Hh, B * C, Hh, Hh, 0, true, num_sms_, st));
    const bool nrm = cfg_.normalize;
    const float eps = cfg_.norm_eps, ls = nrm ? cfg_.logit_scale : 1.0f;
    const int R = B * C;
    if (cfg_.scorer == SCORER_DOT) {
      GLC_LAUNCH(KC_HEAD_MISC, head_score_ex(tvec_, Hh, kvec_, d_logits, d_probs, d_decisions, threshold, B, C, Hh, nrm, eps, ls,
                                             0.f, st));
    } else if (cfg_.scorer == SCORER_MLP) {
      // cat[t_b | k_bc] -> Linear-ReLU -> Linear-ReLU -> Linear(1)
      const int m1 = cfg_.mlp1, m2 = cfg_.mlp2;
      GLC_LAUNCH(KC_HEAD_MISC, head_rows16(tvec_, Hh, C, cat16_, 2 * Hh, 0, R, nrm, eps, st));
      GLC_LAUNCH(KC_HEAD_MISC, head_rows16(kvec_, Hh, 1, cat16_, 2 * Hh, Hh, R, nrm, eps, st));
      GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(cat16_, 2 * Hh, m1w_, 2 * Hh, m1b_, s1_, m1, R, m1, 2 * Hh, 2, false, num_sms_, st));
      GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(s1_, m1, m2w_, m1, m2b_, s2_, m2, R, m2, m1, 2, true, num_sms_, st));
      GLC_LAUNCH(KC_HEAD_MISC, head_score_ex(last_w_, 0, s2_, d_logits, d_probs, d_decisions, threshold, B, C, m2, false, 0.f, ls,
                                             last_b_ * ls, st));
    } else {
      // proj_text / proj_label -> (d, half) -> cat[t0 | l0 | t1*l1] -> Linear-ReLU -> Linear(1)
      GLC_LAUNCH(KC_HEAD_MISC, head_rows16(tvec_, Hh, 1, t16_, Hh, 0, B, nrm, eps, st));
      GLC_LAUNCH(KC_HEAD_MISC, head_rows16(kvec_, Hh, 1, k16_, Hh, 0, R, nrm, eps, st));
      GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(t16_, Hh, ptw_, Hh, ptb_, pt_, 2 * Hh, B, 2 * Hh, Hh, 0, true, num_sms_, st));
      GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(k16_, Hh, plw_, Hh, plb_, pl_, 2 * Hh, R, 2 * Hh, Hh, 0, true, num_sms_, st));
      GLC_LAUNCH(KC_HEAD_MISC, head_wdot_combine(pt_, pl_, cat16_, B, C, Hh, st));
      GLC_LAUNCH(KC_HEAD_GEMM, gemm_f16(cat16_, 3 * Hh, o1w_, 3 * Hh, o1b_, s2_, 4 * Hh, R, 4 * Hh, 3 * Hh, 2, true, num_sms_, st));
      GLC_LAUNCH(KC_HEAD_MISC, head_score_ex(last_w_, 0, s2_, d_logits, d_probs, d_decisions, threshold, B, C, 4 * Hh, false, 0.f,
                                             ls, last_b_ * ls, st));
    }
  }
  launches_ += n;
}

namespace {
bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}
}  // namespace

DeviceModel::PinnedBlock DeviceModel::take_out_block(size_t bytes) {   // caller holds mu
  for (size_t i = 0; i < out_pool_.size(); ++i)
    if (out_pool_[i].bytes >= bytes) {
      PinnedBlock b = out_pool_[i];
      out_pool_.erase(out_pool_.begin() + (long)i);
      return b;
    }
  PinnedBlock b;
  b.bytes = bytes < 4096 ? 4096 : bytes;
  GLC_CUDA(cudaMallocHost(&b.p, b.bytes));
  return b;
}

void DeviceModel::check_overflow_sync() {
  std::lock_guard<std::mutex> lk(mu);
  GLC_CUDA(cudaSetDevice(device_));
  int v = 0;
  GLC_CUDA(cudaMemcpyAsync(&h_overflow_[63], d_overflow_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
  GLC_CUDA(cudaMemsetAsync(d_overflow_, 0, sizeof(int), stream_));
  GLC_CUDA(cudaStreamSynchronize(stream_));
  v = h_overflow_[63];
  if (v) throw std::runtime_error(kOverflowMsg);
}

bool make_pack_plan(const int64_t* ids, const int64_t* mask, int B, int S, int64_t class_token, int class_pos_offset, int max_rows,
                    PackPlan& pl) {
  pl = PackPlan{};
  if (B <= 0 || S <= 0) return false;
  pl.len.resize(B);
  pl.prow.resize(B);
  int64_t total = 0;
  for (int b = 0; b < B; ++b) {
    const int64_t* m = mask + (size_t)b * S;
    int L = S;
    while (L > 0 && m[L - 1] == 0) --L;
    // every class token (and the neighbour embed_class_token=false reads) must lie inside the kept positions
    const int64_t* id = ids + (size_t)b * S;
    int j0 = L - class_pos_offset;
    for (int j = j0 < 0 ? 0 : j0; j < S; ++j)
      if (id[j] == class_token) return false;
    pl.len[b] = L;
    pl.prow[b] = L <= 128 ? 128 : round_up(L, 128);
    total += pl.prow[b];
  }
  pl.total_rows = total;
  if (total * 10 > (int64_t)B * S * 9) return false;   // under 10 % of the rows are padding: not worth leaving the graph path
  PackPlan::MB cur{0, 0, 0, 0};
  auto close = [&](int b1) {
    cur.b1 = b1;
    pl.mbs.push_back(cur);
    if (cur.rows > pl.max_mb_rows) pl.max_mb_rows = cur.rows;
    if (b1 - cur.b0 > pl.max_mb_texts) pl.max_mb_texts = b1 - cur.b0;
    cur = PackPlan::MB{b1, b1, 0, 0};
  };
  for (int b = 0; b < B; ++b) {
    if (cur.rows + pl.prow[b] > max_rows && b > cur.b0) close(b);
    cur.rows += pl.prow[b];
    if (pl.prow[b] > cur.max_rows) cur.max_rows = pl.prow[b];
  }
  close(B);
  return true;
}

bool DeviceModel::plan_pack(const int64_t* ids, const int64_t* mask, int B, int S, PackPlan& pl) const {
  if (!varlen_ || cfg_.pooling == POOL_LAST || S < 256 || S > 2048) return false;
  if ((int64_t)B * S < 8192) return false;   // small requests replay a captured graph of the [B,S] layout: latency first
  return make_pack_plan(ids, mask, B, S, cfg_.class_token, cfg_.class_pos_offset, max_tokens_, pl);
}

void DeviceModel::run_host(const int64_t* ids, const int64_t* mask, int B, int S, int C, float* logits,
                           const DecisionOut* dec, bool cacheable) {
  if (B <= 0 || S <= 0) return;
  PackPlan plan;
  const bool packed = plan_pack(ids, mask, B, S, plan);
  static const bool timing = getenv("GLC_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = timing ? now() : 0;
  // Enqueue under the device lock, wait OUTSIDE it on an event recorded behind this request's last copy: the next
  // caller's H2D copies and graph launch are queued behind ours while we still wait, so back-to-back requests (two
  // glc_submit tickets in flight, or two OpenMP workers) leave no bubble on the device.  One stream keeps the shared
  // workspace and the id / mask / logits staging buffers safe: request N+1's copies execute after request N's forward
  // and D2H in stream order.
  // Pageable caller buffers (what the reference passes: malloc'd int64 arrays, src/model.c:17-29, and ORT-owned output)
  // never meet cudaMemcpyAsync directly — that call degrades to a synchronous staged copy which would hold `mu` for the
  // whole forward.  Inputs are copied into one of two pinned slots, outputs land in a pinned block that is copied to the
  // caller after the completion event.
  const bool want_p = dec && dec->probs, want_d = dec && dec->decisions;
  const bool in_pinned = is_pinned(ids) && is_pinned(mask);
  const bool out_pinned = is_pinned(logits) && (!want_p || is_pinned(dec->probs)) && (!want_d || is_pinned(dec->decisions));
  const size_t nout = (size_t)B * (size_t)(C > 0 ? C : 0);
  PinnedBlock ob;
  float* o_logits = logits;
  float* o_probs = want_p ? dec->probs : nullptr;
  uint8_t* o_dec = want_d ? dec->decisions : nullptr;
  cudaEvent_t done = nullptr;
  double t1 = 0;
  uint32_t oslot = 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    GLC_CUDA(cudaSetDevice(device_));
    int rows_mb = max_tokens_ / S;
    if (rows_mb < 1) rows_mb = 1;
    if (rows_mb > B) rows_mb = B;
    if (packed) ensure_workspace(plan.max_mb_rows, plan.max_mb_texts, C);
    else ensure_workspace(rows_mb * S, rows_mb, C);
    if (!out_pinned && nout > 0) {
      ob = take_out_block(nout * 9);
      o_logits = logits ? (float*)ob.p : nullptr;
      o_probs = want_p ? (float*)ob.p + nout : nullptr;
      o_dec = want_d ? (uint8_t*)((float*)ob.p + 2 * nout) : nullptr;
    }
    for (size_t mi = 0; packed && mi < plan.mbs.size(); ++mi) {
      // ---- packed micro-batch: compact the texts into a pinned slot (ids | mask | text_row | kv_len | tile_info), one
      //      H2D each, then the same forward on rows = sum of the texts' 128-aligned lengths
      const PackPlan::MB& mb = plan.mbs[mi];
      const int nb = mb.b1 - mb.b0, rows = mb.rows, nt = rows / 128;
      const size_t idb = (size_t)rows * 8, nints = (size_t)2 * nb + 1 + 2 * nt;
      const int k = h_in_next_;
      h_in_next_ ^= 1;
      PinnedBlock& hb = h_in_[k];
      if (hb.bytes < 2 * idb + nints * 4) {
        if (hb.p) { GLC_CUDA(cudaStreamSynchronize(stream_)); cudaFreeHost(hb.p); hb.p = nullptr; }
        hb.bytes = 2 * (size_t)plan.max_mb_rows * 8 + ((size_t)2 * plan.max_mb_texts + 2 + 2 * (plan.max_mb_rows / 128)) * 4;
        if (hb.bytes < 2 * idb + nints * 4) hb.bytes = 2 * idb + nints * 4;
        GLC_CUDA(cudaMallocHost(&hb.p, hb.bytes));
      }
      if (!h_in_ev_[k]) GLC_CUDA(cudaEventCreateWithFlags(&h_in_ev_[k], cudaEventDisableTiming));
      else GLC_CUDA(cudaEventSynchronize(h_in_ev_[k]));
      int64_t* pi = (int64_t*)hb.p;
      int64_t* pm = pi + rows;
      int32_t* text_row = (int32_t*)(pm + rows);
      int32_t* kvl = text_row + nb + 1;
      int32_t* tinfo = kvl + nb;
      int32_t* tpos = tinfo + nt;   // position of each 128-row tile's first row within its text (rotary embedding)
      int row = 0;
      for (int i = 0; i < nb; ++i) {
        const int b = mb.b0 + i, L = plan.len[b], P = plan.prow[b];
        text_row[i] = row;
        kvl[i] = L;
        for (int q = 0; q * 128 < P; ++q) tpos[row / 128 + q] = q * 128;
        memcpy(pi + row, ids + (size_t)b * S, (size_t)L * 8);
        memcpy(pm + row, mask + (size_t)b * S, (size_t)L * 8);
        memset(pi + row + L, 0, (size_t)(P - L) * 8);   // id 0 / mask 0, what the reference pads with (tokenizer.c:78-82)
        memset(pm + row + L, 0, (size_t)(P - L) * 8);
        row += P;
      }
      text_row[nb] = row;
      int n_t = 0;
      for (int q = 0; q * 128 < mb.max_rows; ++q)
        for (int i = 0; i < nb; ++i)
          if (q * 128 < plan.prow[mb.b0 + i]) tinfo[n_t++] = (q << 24) | i;
      GLC_CUDA(cudaMemcpyAsync(ids_, pi, idb, cudaMemcpyHostToDevice, stream_));
      GLC_CUDA(cudaMemcpyAsync(mask_, pm, idb, cudaMemcpyHostToDevice, stream_));
      GLC_CUDA(cudaMemcpyAsync(pk_ints_, text_row, nints * 4, cudaMemcpyHostToDevice, stream_));
      GLC_CUDA(cudaEventRecord(h_in_ev_[k], stream_));
      PackedCtx pc;
      pc.rows = rows; pc.max_rows = mb.max_rows; pc.n_tiles = nt;
      pc.text_row = pk_ints_; pc.kv_len = pk_ints_ + nb + 1; pc.tile_info = pk_ints_ + 2 * nb + 1; pc.tile_pos = pc.tile_info + nt;
      pk_ = &pc;
      try {
        forward_eager(ids_, mask_, nb, S, C, logits_, want_p ? probs_ : nullptr, want_d ? decisions_ : nullptr,
                      dec ? dec->threshold : 0.5f);
      } catch (...) {
        pk_ = nullptr;
        throw;
      }
      pk_ = nullptr;
      packed_runs_ += 1;
      packed_rows_ += (uint64_t)rows;
      packed_rows_padded_ += (uint64_t)nb * S;
      if (C > 0) {
        const int r0 = mb.b0;
        if (o_logits)
          GLC_CUDA(cudaMemcpyAsync(o_logits + (size_t)r0 * C, logits_, (size_t)nb * C * 4, cudaMemcpyDeviceToHost, stream_));
        if (o_probs)
          GLC_CUDA(cudaMemcpyAsync(o_probs + (size_t)r0 * C, probs_, (size_t)nb * C * 4, cudaMemcpyDeviceToHost, stream_));
        if (o_dec)
          GLC_CUDA(cudaMemcpyAsync(o_dec + (size_t)r0 * C, decisions_, (size_t)nb * C, cudaMemcpyDeviceToHost, stream_));
      }
    }
    for (int r0 = 0; !packed && r0 < B; r0 += rows_mb) {
      const int nb = (B - r0 < rows_mb) ? (B - r0) : rows_mb;
      const size_t bytes = (size_t)nb * S * 8;
      const int64_t* src_i = ids + (size_t)r0 * S;
      const int64_t* src_m = mask + (size_t)r0 * S;
      if (!in_pinned) {
        const int k = h_in_next_;
        h_in_next_ ^= 1;
        PinnedBlock& hb = h_in_[k];
        if (hb.bytes < 2 * bytes) {
          if (hb.p) { GLC_CUDA(cudaStreamSynchronize(stream_)); cudaFreeHost(hb.p); hb.p = nullptr; }
          hb.bytes = 2 * (size_t)rows_mb * S * 8;
          GLC_CUDA(cudaMallocHost(&hb.p, hb.bytes));
        }
        if (!h_in_ev_[k]) GLC_CUDA(cudaEventCreateWithFlags(&h_in_ev_[k], cudaEventDisableTiming));
        else GLC_CUDA(cudaEventSynchronize(h_in_ev_[k]));   // the H2D that last read this slot has finished
        memcpy(hb.p, src_i, bytes);
        memcpy((uint8_t*)hb.p + bytes, src_m, bytes);
        src_i = (const int64_t*)hb.p;
        src_m = (const int64_t*)((uint8_t*)hb.p + bytes);
        GLC_CUDA(cudaMemcpyAsync(ids_, src_i, bytes, cudaMemcpyHostToDevice, stream_));
        GLC_CUDA(cudaMemcpyAsync(mask_, src_m, bytes, cudaMemcpyHostToDevice, stream_));
        GLC_CUDA(cudaEventRecord(h_in_ev_[k], stream_));
      } else {
        GLC_CUDA(cudaMemcpyAsync(ids_, src_i, bytes, cudaMemcpyHostToDevice, stream_));
        GLC_CUDA(cudaMemcpyAsync(mask_, src_m, bytes, cudaMemcpyHostToDevice, stream_));
      }
      forward(ids_, mask_, nb, S, C, logits_, want_p ? probs_ : nullptr, want_d ? decisions_ : nullptr,
              dec ? dec->threshold : 0.5f, cacheable);
      if (C > 0) {
        if (o_logits)
          GLC_CUDA(cudaMemcpyAsync(o_logits + (size_t)r0 * C, logits_, (size_t)nb * C * 4, cudaMemcpyDeviceToHost, stream_));
        if (o_probs)
          GLC_CUDA(cudaMemcpyAsync(o_probs + (size_t)r0 * C, probs_, (size_t)nb * C * 4, cudaMemcpyDeviceToHost, stream_));
        if (o_dec)
          GLC_CUDA(cudaMemcpyAsync(o_dec + (size_t)r0 * C, decisions_, (size_t)nb * C, cudaMemcpyDeviceToHost, stream_));
      }
    }
    // this request's copy of the fp16-saturation flag; the device flag is cleared behind it for the next request
    oslot = overflow_seq_++ % 63u;
    GLC_CUDA(cudaMemcpyAsync(&h_overflow_[oslot], d_overflow_, sizeof(int), cudaMemcpyDeviceToHost, stream_));
    GLC_CUDA(cudaMemsetAsync(d_overflow_, 0, sizeof(int), stream_));
    if (free_events_.empty()) {
      GLC_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
    } else {
      done = free_events_.back();
      free_events_.pop_back();
    }
    GLC_CUDA(cudaEventRecord(done, stream_));
    t1 = timing ? now() : 0;
  }
  const cudaError_t we = cudaEventSynchronize(done);
  const int overflowed = (we == cudaSuccess) ? h_overflow_[oslot] : 0;
  if (we == cudaSuccess && !overflowed && ob.p) {
    if (logits) memcpy(logits, o_logits, nout * 4);
    if (want_p) memcpy(dec->probs, o_probs, nout * 4);
    if (want_d) memcpy(dec->decisions, o_dec, nout);
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    free_events_.push_back(done);
    if (ob.p) out_pool_.push_back(ob);
  }
  if (we != cudaSuccess) throw std::runtime_error(std::string("cudaEventSynchronize: ") + cudaGetErrorString(we));
  if (overflowed) throw std::runtime_error(kOverflowMsg);
  if (timing) fprintf(stderr, "glc run_host B=%d S=%d: enqueue %.1f us, wait %.1f us\n", B, S, t1 - t0, now() - t1);
}

// ---------------------------------------------------------------------------------------------
// request coalescing (see engine.h)

void DeviceModel::run_host_coalesced(HostReq& r) {
  std::unique_lock<std::mutex> lk(qmu_);
  queue_.push_back(&r);
  if (leader_) qcv_.notify_all();   // a leader inside its batching window re-checks the queue length
  while (!r.done) {
    if (leader_) {
      qcv_.wait(lk);
      continue;
    }
    // device idle: lead.  Take requests from the front while the padded group fits one launch.
    leader_ = true;
    // The caller that led the previous group is usually back first with its next request and would lead ALONE (the
    // reference's OpenMP loop, main.c:141-150: launches of 15 requests alternated with launches of 1) — when the last
    // group was a merged one, give the other callers a moment to enqueue: at most 150 us, or until half of that group's
    // size has arrived.  A lone caller (last group of size 1) never waits.
    if (last_group_size_ > 1) {
      const size_t want = (size_t)(last_group_size_ + 1) / 2;
      qcv_.wait_for(lk, std::chrono::microseconds(150), [&] { return queue_.size() >= want; });
    }
    std::vector<HostReq*> group;
    int rows = 0, smax = 0;
    while (!queue_.empty()) {
      HostReq* q = queue_.front();
      const int ns = q->S > smax ? q->S : smax;
      if (!group.empty()) {
        if ((int64_t)(rows + q->B) * ns > max_tokens_) break;
        // one threshold per launch: group only requests that agree on the decision epilogue
        // 'last' pooling reads position S-1, padded or not: its value depends on the batch's padded length
        if (cfg_.pooling == POOL_LAST && q->S != smax) break;
        const DecisionOut *a = group[0]->dec, *b = q->dec;
        if ((a != nullptr) != (b != nullptr) || (a && b && a->threshold != b->threshold)) break;
      }
      group.push_back(q);
      queue_.pop_front();
      rows += q->B;
      smax = ns;
    }
    lk.unlock();
    std::exception_ptr err;
    try {
      run_group(group);
    } catch (...) {
      err = std::current_exception();
    }
    lk.lock();
    for (HostReq* q : group) { q->err = err; q->done = true; }
    last_group_size_ = (int)group.size();
    leader_ = false;
    qcv_.notify_all();
  }
  lk.unlock();
  if (r.err) std::rethrow_exception(r.err);
}

void DeviceModel::run_group(std::vector<HostReq*>& group) {
  if (group.size() == 1) {
    HostReq* q = group[0];
    run_host(q->ids, q->mask, q->B, q->S, q->C, q->logits, q->dec);
    return;
  }
  int rows = 0, S = 0, C = 0;
  bool want_l = false, want_p = false, want_d = false;
  for (HostReq* q : group) {
    rows += q->B;
    if (q->S > S) S = q->S;
    if (q->C > C) C = q->C;
    want_l |= q->logits != nullptr;
    want_p |= q->dec && q->dec->probs;
    want_d |= q->dec && q->dec->decisions;
  }
  GLC_CUDA(cudaSetDevice(device_));
  const size_t tok = (size_t)rows * S, nout = (size_t)rows * C;
  if (tok > h_tok_) {
    if (h_ids_) { cudaFreeHost(h_ids_); cudaFreeHost(h_mask_); }
    h_tok_ = tok > (size_t)max_tokens_ ? tok : (size_t)max_tokens_;
    GLC_CUDA(cudaMallocHost((void**)&h_ids_, h_tok_ * 8));
    GLC_CUDA(cudaMallocHost((void**)&h_mask_, h_tok_ * 8));
  }
  if (nout > h_rows_) {
    if (h_logits_) { cudaFreeHost(h_logits_); cudaFreeHost(h_probs_); cudaFreeHost(h_dec_); }
    h_rows_ = nout * 2;
    GLC_CUDA(cudaMallocHost((void**)&h_logits_, h_rows_ * 4));
    GLC_CUDA(cudaMallocHost((void**)&h_probs_, h_rows_ * 4));
    GLC_CUDA(cudaMallocHost((void**)&h_dec_, h_rows_));
  }
  // pad every request to the group's longest sequence: id 0 / mask 0 (tokenizer.c:78-82)
  int r0 = 0;
  for (HostReq* q : group) {
    for (int b = 0; b < q->B; ++b) {
      int64_t* di = h_ids_ + (size_t)(r0 + b) * S;
      int64_t* dm = h_mask_ + (size_t)(r0 + b) * S;
      memcpy(di, q->ids + (size_t)b * q->S, (size_t)q->S * 8);
      memcpy(dm, q->mask + (size_t)b * q->S, (size_t)q->S * 8);
      if (q->S < S) {
        memset(di + q->S, 0, (size_t)(S - q->S) * 8);
        memset(dm + q->S, 0, (size_t)(S - q->S) * 8);
      }
    }
    r0 += q->B;
  }
  DecisionOut d;
  d.threshold = group[0]->dec ? group[0]->dec->threshold : 0.5f;
  d.probs = want_p ? h_probs_ : nullptr;
  d.decisions = want_d ? h_dec_ : nullptr;
  run_host(h_ids_, h_mask_, rows, S, C, want_l ? h_logits_ : nullptr, (want_p || want_d) ? &d : nullptr, /*cacheable=*/false);
  // scatter: request i gets its rows, first C_i columns (columns >= C_i only exist because another request of the
  // group has more labels; columns in [count(b), C_i) are the zero-padded classes the reference also scores)
  r0 = 0;
  for (HostReq* q : group) {
    for (int b = 0; b < q->B; ++b) {
      const size_t src = (size_t)(r0 + b) * C, dst = (size_t)b * q->C;
      if (q->logits) memcpy(q->logits + dst, h_logits_ + src, (size_t)q->C * 4);
      if (q->dec && q->dec->probs) memcpy(q->dec->probs + dst, h_probs_ + src, (size_t)q->C * 4);
      if (q->dec && q->dec->decisions) memcpy(q->dec->decisions + dst, h_dec_ + src, (size_t)q->C);
    }
    r0 += q->B;
  }
  merged_groups_ += 1;
  merged_requests_ += group.size();
}

// ---------------------------------------------------------------------------------------------

TaskQueue::TaskQueue(int threads) {
  for (int i = 0; i < threads; ++i) th_.emplace_back([this] { loop(); });
}

TaskQueue::~TaskQueue() {
  {
    std::lock_guard<std::mutex> lk(mu_);
    stop_ = true;
  }
  cv_.notify_all();
  for (auto& t : th_) t.join();
}

void TaskQueue::post(std::function<void()> fn) {
  {
    std::lock_guard<std::mutex> lk(mu_);
    q_.push_back(std::move(fn));
  }
  cv_.notify_one();
}

void TaskQueue::loop() {
  for (;;) {
    std::function<void()> fn;
    {
      std::unique_lock<std::mutex> lk(mu_);
      cv_.wait(lk, [this] { return stop_ || !q_.empty(); });
      if (q_.empty()) return;   // stop requested and drained
      fn = std::move(q_.front());
      q_.pop_front();
    }
    fn();
  }
}

Model::Model(const std::string& onnx_path, const std::vector<int>& devices, int max_tokens, bool preln_f32, bool fp8_ffn) {
  ModelWeights w;
  load_model_weights(onnx_path, &w);
  cfg_ = w.cfg;
  if (devices.empty()) throw std::runtime_error("no CUDA device selected");
  for (int d : devices) devs_.emplace_back(new DeviceModel(d, w, max_tokens, preln_f32, fp8_ffn));
  if (devs_.size() > 1)
    for (size_t i = 0; i < devs_.size(); ++i) workers_.emplace_back(new TaskQueue(1));
  const char* co = getenv("GLC_COALESCE");
  coalesce_ = !(co && co[0] == '0');
  coalesce_tokens_ = (max_tokens > 0 ? max_tokens : 65536) / 2;
  if (const char* ct = getenv("GLC_COALESCE_TOKENS")) coalesce_tokens_ = atoi(ct);
}

Model::~Model() {
  submit_q_.reset();   // drains in-flight submits before the devices go away
  workers_.clear();
}

TaskQueue& Model::submit_queue() {
  std::lock_guard<std::mutex> lk(submit_mu_);
  if (!submit_q_) {
    int n = 8;   // in-flight glc_submit requests served concurrently (each blocks in the coalescing queue like an OpenMP worker)
    if (const char* e = getenv("GLC_SUBMIT_THREADS")) n = atoi(e) > 0 ? atoi(e) : n;
    submit_q_.reset(new TaskQueue(n));
  }
  return *submit_q_;
}

void Model::coalesce_stats(uint64_t* groups, uint64_t* requests) const {
  uint64_t g = 0, r = 0;
  for (auto& d : devs_) { g += d->merged_groups(); r += d->merged_requests(); }
  if (groups) *groups = g;
  if (requests) *requests = r;
}

static int count_classes_rows(const int64_t* ids, int B, int S, int64_t class_token) {
  int C = 0;
  for (int b = 0; b < B; ++b) {
    int n = 0;
    const int64_t* row = ids + (size_t)b * S;
    for (int j = 0; j < S; ++j) n += (row[j] == class_token);
    if (n > C) C = n;
  }
  return C;
}

int Model::num_classes(const int64_t* ids, int B, int S) const {
  const int G = (int)workers_.size();
  if (G < 2 || (int64_t)B * S < (1 << 20)) return count_classes_rows(ids, B, S, cfg_.class_token);
  // a reranker-sized batch (4096 x 1024 ids = 33 MB): the scan is the only serial host work in front of a sharded run,
  // so the per-device workers do it
  const int per = (B + G - 1) / G;
  std::vector<int> part(G, 0);
  std::mutex lmu;
  std::condition_variable lcv;
  int pending = 0;
  for (int g = 0; g < G; ++g) {
    const int r0 = g * per, nb = (r0 >= B) ? 0 : ((B - r0 < per) ? B - r0 : per);
    if (nb == 0) continue;
    {
      std::lock_guard<std::mutex> lk(lmu);
      ++pending;
    }
    workers_[g]->post([&, g, r0, nb]() {
      part[g] = count_classes_rows(ids + (size_t)r0 * S, nb, S, cfg_.class_token);
      std::lock_guard<std::mutex> lk(lmu);
      if (--pending == 0) lcv.notify_all();
    });
  }
  std::unique_lock<std::mutex> lk(lmu);
  lcv.wait(lk, [&] { return pending == 0; });
  int C = 0;
  for (int v : part) C = v > C ? v : C;
  return C;
}

uint64_t Model::launches() const {
  uint64_t n = 0;
  for (auto& d : devs_) n += d->launches();
  return n;
}

// ORT's Gather fails on an index outside the embedding table; so does this Run — checked before the request can be merged with
// other callers' (the kernel itself clamps, which only the device-pointer entry point relies on)
static void validate_ids(const int64_t* ids, int row0, int B, int S, int vocab) {
  const uint64_t V = (uint64_t)vocab;
  const size_t n = (size_t)B * S;
  for (size_t i = 0; i < n; ++i)
    if ((uint64_t)ids[i] >= V)
      throw std::invalid_argument("input_ids[" + std::to_string(row0 + (int)(i / S)) + "][" + std::to_string(i % S) + "] = " +
                                  std::to_string((long long)ids[i]) + " is outside the embedding table [0, " +
                                  std::to_string(vocab) + ")");
}

void Model::run(const int64_t* ids, const int64_t* mask, int B, int S, int C, float* logits, const DecisionOut* dec) {
  const int G = (int)devs_.size();
  if (G == 1 || B < 2 * G) {
    validate_ids(ids, 0, B, S, cfg_.vocab);
    // small call (the reference's BATCH_SIZE=8 Run): whole batch on one device, round robin
    // across concurrent callers (the OpenMP loop of main.c:141-150)
    const int slot = (int)(rr_.fetch_add(1) % (uint32_t)G);
    if (coalesce_ && (int64_t)B * S <= coalesce_tokens_) {
      DeviceModel::HostReq r;
      r.ids = ids; r.mask = mask; r.B = B; r.S = S; r.C = C; r.logits = logits; r.dec = dec;
      devs_[slot]->run_host_coalesced(r);
    } else {
      devs_[slot]->run_host(ids, mask, B, S, C, logits, dec);
    }
    return;
  }
  // large call: contiguous row shards, one PERSISTENT host thread per device (workers_), host gather into `logits`
  const int per = (B + G - 1) / G;
  std::vector<std::exception_ptr> err(G);
  std::mutex lmu;
  std::condition_variable lcv;
  int pending = 0;
  for (int g = 0; g < G; ++g) {
    const int r0 = g * per;
    const int nb = (r0 >= B) ? 0 : ((B - r0 < per) ? B - r0 : per);
    if (nb == 0) continue;
    {
      std::lock_guard<std::mutex> lk(lmu);
      ++pending;
    }
    workers_[g]->post([&, g, r0, nb]() {
      try {
        DecisionOut sub;
        if (dec) {
          sub.threshold = dec->threshold;
          sub.probs = dec->probs ? dec->probs + (size_t)r0 * C : nullptr;
          sub.decisions = dec->decisions ? dec->decisions + (size_t)r0 * C : nullptr;
        }
        validate_ids(ids + (size_t)r0 * S, r0, nb, S, cfg_.vocab);   // each worker checks its own shard
        devs_[g]->run_host(ids + (size_t)r0 * S, mask + (size_t)r0 * S, nb, S, C, logits ? logits + (size_t)r0 * C : nullptr,
                           dec ? &sub : nullptr);
      } catch (...) {
        err[g] = std::current_exception();
      }
      std::lock_guard<std::mutex> lk(lmu);
      if (--pending == 0) lcv.notify_all();
    });
  }
  {
    std::unique_lock<std::mutex> lk(lmu);
    lcv.wait(lk, [&] { return pending == 0; });
  }
  for (auto& e : err)
    if (e) std::rethrow_exception(e);
}

}  // namespace glc
