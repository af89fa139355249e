// ONNX graph -> role-named fp32 host tensors.  Roles:
//   emb.word [V,H]  emb.ln.g/b [H]  rel.emb [2*buckets,H]  rel.ln.g/b [H]
//   layer.<l>.{q,k,v,o}.w [H,H]  .b [H]   layer.<l>.ln1.g/b
//   layer.<l>.ffn1.w [I,H] .b [I]  layer.<l>.ffn2.w [H,I] .b [H]  layer.<l>.ln2.g/b
//   text.1.w [Hh,H] text.1.b  text.2.w [Hh,Hh] text.2.b   cls.1.* cls.2.*  (FeaturesProjector x2)
//   scorer.mlp.{0,2,4}.w/b (MLP scorer)   scorer.pt / scorer.pl [2Hh,Hh], scorer.o1 [4Hh,3Hh], scorer.o2 [1,4Hh]
//   (weighted-dot scorer); logit_scale is a config scalar
// Naming facts relied on (verified on a real torch export):
//   * 3-D-input nn.Linear -> MatMul(x, anonymous [in,out] initializer) + Add(named bias)
//   * 2-D-input nn.Linear -> Gemm(x, W[out,in], b) with transB=1
//   * node names carry module scope: .../encoder/layer.<l>/attention/self/query_proj/MatMul
//   * identical tensors are deduplicated into Identity aliases (resolved by OnnxGraph::resolve)
#include "model_weights.h"

#include <cmath>
#include <cstring>
#include <stdexcept>

#include "onnx_reader.h"

namespace glc {
namespace {

bool ends_with(const std::string& s, const std::string& suf) {
  return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}

const OnnxNode* find_node(const OnnxGraph& g, const std::string& suffix, const char* op = nullptr) {
  for (auto& n : g.nodes)
    if (ends_with(n.name, suffix) && (!op || n.op_type == op)) return &n;
  return nullptr;
}

HostTensor to_host(const OnnxTensor& t) {
  HostTensor h;
  h.dims = t.dims;
  if (t.numel() < 0) throw std::runtime_error("onnx: bad dimensions in " + t.name);
  h.data.resize((size_t)t.numel());
  tensor_to_float(t, h.data.data());
  return h;
}

HostTensor transpose2d(const HostTensor& a) {
  HostTensor o;
  int64_t r = a.dims[0], c = a.dims[1];
  o.dims = {c, r};
  o.data.resize(a.data.size());
  for (int64_t i = 0; i < r; ++i)
    for (int64_t j = 0; j < c; ++j) o.data[(size_t)(j * r + i)] = a.data[(size_t)(i * c + j)];
  return o;
}

// scope e.g. "/layer.0/attention/self/query_proj"
void load_linear(const OnnxGraph& g, const std::string& scope, const std::string& role, ModelWeights* out) {
  const OnnxNode* mm = find_node(g, scope + "/MatMul", "MatMul");
  if (mm) {
    if (mm->inputs.size() != 2) throw std::runtime_error("onnx: MatMul arity at " + mm->name);
    const OnnxTensor* w = g.resolve(mm->inputs[1]);
    if (!w || w->dims.size() != 2) throw std::runtime_error("onnx: no constant 2-D weight at " + mm->name);
    out->t[role + ".w"] = transpose2d(to_host(*w));   // [in,out] -> [out,in]
    const OnnxNode* add = find_node(g, scope + "/Add", "Add");
    if (!add) throw std::runtime_error("onnx: no bias Add after " + mm->name);
    const OnnxTensor* b = nullptr;
    for (auto& in : add->inputs) { b = g.resolve(in); if (b) break; }
    if (!b) throw std::runtime_error("onnx: no constant bias at " + add->name);
    out->t[role + ".b"] = to_host(*b);
    return;
  }
  const OnnxNode* gm = find_node(g, scope + "/Gemm", "Gemm");
  if (gm) {
    if (gm->inputs.size() < 2) throw std::runtime_error("onnx: Gemm arity at " + gm->name);
    const OnnxTensor* w = g.resolve(gm->inputs[1]);
    if (!w || w->dims.size() != 2) throw std::runtime_error("onnx: no constant 2-D weight at " + gm->name);
    const OnnxAttr* tb = gm->attr("transB");
    const OnnxAttr* ta = gm->attr("transA");
    const OnnxAttr* al = gm->attr("alpha");
    const OnnxAttr* be = gm->attr("beta");
    if ((ta && ta->i != 0) || (al && al->f != 1.0f) || (be && be->f != 1.0f))
      throw std::runtime_error("onnx: unsupported Gemm attributes at " + gm->name);
    HostTensor hw = to_host(*w);
    out->t[role + ".w"] = (tb && tb->i == 1) ? hw : transpose2d(hw);
    HostTensor hb;
    if (gm->inputs.size() >= 3) {
      const OnnxTensor* b = g.resolve(gm->inputs[2]);
      if (!b) throw std::runtime_error("onnx: non-constant Gemm bias at " + gm->name);
      hb = to_host(*b);
    } else {
      hb.dims = {out->t[role + ".w"].dims[0]};
      hb.data.assign((size_t)hb.dims[0], 0.f);
    }
    out->t[role + ".b"] = hb;
    return;
  }
  throw std::runtime_error("onnx: no MatMul/Gemm node for scope " + scope);
}

// decomposed opset-14 LayerNorm: .../LayerNorm/Mul(x, gamma), .../LayerNorm/Add_1(x, beta)
void load_layernorm(const OnnxGraph& g, const std::string& scope, const std::string& role, int64_t H,
                    ModelWeights* out) {
  auto pick = [&](const char* leaf, const char* op) -> HostTensor {
    const OnnxNode* n = find_node(g, scope + "/LayerNorm/" + leaf, op);
    if (!n) throw std::runtime_error("onnx: no " + scope + "/LayerNorm/" + leaf);
    for (auto& in : n->inputs) {
      const OnnxTensor* t = g.resolve(in);
      if (t && t->numel() == H) return to_host(*t);
    }
    throw std::runtime_error("onnx: no constant [H] operand at " + n->name);
  };
  out->t[role + ".g"] = pick("Mul", "Mul");
  out->t[role + ".b"] = pick("Add_1", "Add");
}

const OnnxTensor* find_named(const OnnxGraph& g, const std::string& suffix) {
  for (auto& t : g.initializers)
    if (ends_with(t.name, suffix)) return &t;
  for (auto& n : g.nodes)
    if (n.op_type == "Identity" && n.outputs.size() == 1 && ends_with(n.outputs[0], suffix))
      return g.resolve(n.outputs[0]);
  return nullptr;
}

}  // namespace

const HostTensor& ModelWeights::at(const std::string& role) const {
  auto it = t.find(role);
  if (it == t.end()) throw std::runtime_error("weights: missing role " + role);
  return it->second;
}

namespace {

// MatMul-only Linear (no bias Add after it): o_proj / gate_proj / up_proj / down_proj of the decoder backbones
void load_linear_nobias(const OnnxGraph& g, const std::string& scope, const std::string& role, ModelWeights* out) {
  const OnnxNode* mm = find_node(g, scope + "/MatMul", "MatMul");
  if (!mm || mm->inputs.size() != 2) throw std::runtime_error("onnx: no MatMul node for scope " + scope);
  const OnnxTensor* w = g.resolve(mm->inputs[1]);
  if (!w || w->dims.size() != 2) throw std::runtime_error("onnx: no constant 2-D weight at " + mm->name);
  out->t[role + ".w"] = transpose2d(to_host(*w));   // [in,out] -> [out,in]
}

// Qwen2-style decoder stack traced from transformers' Qwen2Model (modeling_qwen2.py): scopes /layers.<l>/self_attn/
// {q,k,v,o}_proj, /layers.<l>/mlp/{gate,up,down}_proj, RMSNorm weights as named initializers.
void load_qwen2_encoder(const OnnxGraph& g, ModelWeights* out) {
  ModelConfig& c = out->cfg;
  c.backbone = BACKBONE_QWEN2;
  const OnnxTensor* we = find_named(g, "embed_tokens.weight");
  if (!we || we->dims.size() != 2) throw std::runtime_error("onnx: embed_tokens.weight not found");
  c.vocab = (int)we->dims[0];
  c.hidden = (int)we->dims[1];
  out->t["emb.word"] = to_host(*we);
  for (auto& n : g.nodes)
    if (n.op_type == "Trilu")
      throw std::runtime_error("onnx: causal attention mask (Trilu at " + n.name + "): decoder backbones are supported as "
                               "bidirectional encoders only");
  // rotary inverse frequencies: the float constant expanded in /rotary_emb/Expand
  {
    const OnnxNode* ex = find_node(g, "/rotary_emb/Expand", "Expand");
    const OnnxTensor* f = ex && !ex->inputs.empty() ? g.resolve(ex->inputs[0]) : nullptr;
    if (!f || f->data_type != 1 || f->numel() < 8) throw std::runtime_error("onnx: rotary inv_freq constant not found");
    HostTensor h = to_host(*f);
    h.dims = {(int64_t)h.data.size()};
    c.head_dim = 2 * (int)h.data.size();
    out->t["rope.inv_freq"] = h;
  }
  int L = 0;
  while (find_node(g, "/layers." + std::to_string(L) + "/self_attn/q_proj/MatMul")) ++L;
  if (L == 0) throw std::runtime_error("onnx: no decoder layers found");
  c.layers = L;
  auto named_vec = [&](const std::string& name, const std::string& role) {
    const OnnxTensor* t = find_named(g, name);
    if (!t || t->numel() != c.hidden) throw std::runtime_error("onnx: " + name + " not found");
    out->t[role] = to_host(*t);
  };
  for (int l = 0; l < L; ++l) {
    const std::string s = "/layers." + std::to_string(l), r = "layer." + std::to_string(l), nm = "layers." + std::to_string(l) + ".";
    load_linear(g, s + "/self_attn/q_proj", r + ".q", out);
    load_linear(g, s + "/self_attn/k_proj", r + ".k", out);
    load_linear(g, s + "/self_attn/v_proj", r + ".v", out);
    load_linear_nobias(g, s + "/self_attn/o_proj", r + ".o", out);
    load_linear_nobias(g, s + "/mlp/gate_proj", r + ".gate", out);
    load_linear_nobias(g, s + "/mlp/up_proj", r + ".up", out);
    load_linear_nobias(g, s + "/mlp/down_proj", r + ".down", out);
    named_vec(nm + "input_layernorm.weight", r + ".ln1.g");
    named_vec(nm + "post_attention_layernorm.weight", r + ".ln2.g");
    // hidden_act must be SiLU (Sigmoid + Mul under /mlp)
    bool sig = false;
    for (auto& n : g.nodes)
      if (n.name.find(s + "/mlp/") != std::string::npos && n.op_type == "Sigmoid") sig = true;
    if (!sig) throw std::runtime_error("onnx: layer " + std::to_string(l) + " MLP activation is not SiLU");
  }
  named_vec("encoder_model.norm.weight", "norm.g");
  if (const OnnxNode* n = find_node(g, "/layers.0/input_layernorm/Add", "Add")) {
    float e;
    for (auto& in : n->inputs)
      if (g.scalar_float(in, &e) && e > 0.f && e < 1e-2f) c.rms_eps = e;
  }
  c.inter = (int)out->at("layer.0.gate.w").dims[0];
  const int64_t qo = out->at("layer.0.q.w").dims[0], ko = out->at("layer.0.k.w").dims[0];
  if (c.head_dim <= 0 || qo % c.head_dim || ko % c.head_dim) throw std::runtime_error("onnx: q/k projection widths do not match the rotary head dim");
  c.heads = (int)(qo / c.head_dim);
  c.kv_heads = (int)(ko / c.head_dim);
  if (c.kv_heads <= 0 || c.heads % c.kv_heads) throw std::runtime_error("onnx: heads not divisible by kv heads");
  auto expect = [&](const std::string& role, std::vector<int64_t> d) {
    if (out->at(role).dims != d) throw std::runtime_error("weights: unexpected shape for " + role);
  };
  const int64_t H = c.hidden, I = c.inter, Q = qo, K = ko;
  for (int l = 0; l < L; ++l) {
    const std::string r = "layer." + std::to_string(l);
    expect(r + ".q.w", {Q, H}); expect(r + ".q.b", {Q});
    expect(r + ".k.w", {K, H}); expect(r + ".k.b", {K});
    expect(r + ".v.w", {K, H}); expect(r + ".v.b", {K});
    expect(r + ".o.w", {H, Q});
    expect(r + ".gate.w", {I, H}); expect(r + ".up.w", {I, H}); expect(r + ".down.w", {H, I});
  }
}

void load_head(const OnnxGraph& g, ModelWeights* out);
void load_deberta_encoder(const OnnxGraph& g, ModelWeights* out);

}  // namespace

void load_model_weights(const std::string& path, ModelWeights* out) {
  OnnxGraph g;
  g.load(path);
  if (g.input_names.size() < 2 || g.input_names[0] != "input_ids" || g.input_names[1] != "attention_mask")
    throw std::runtime_error("onnx: expected graph inputs (input_ids, attention_mask)");
  if (find_named(g, "embed_tokens.weight")) load_qwen2_encoder(g, out);
  else load_deberta_encoder(g, out);
  load_head(g, out);
}

namespace {

void load_deberta_encoder(const OnnxGraph& g, ModelWeights* out) {
  ModelConfig& c = out->cfg;

  // ---- embeddings
  const OnnxTensor* we = find_named(g, "embeddings.word_embeddings.weight");
  if (!we || we->dims.size() != 2) throw std::runtime_error("onnx: word_embeddings.weight not found");
  c.vocab = (int)we->dims[0];
  c.hidden = (int)we->dims[1];
  out->t["emb.word"] = to_host(*we);
  load_layernorm(g, "/embeddings", "emb.ln", c.hidden, out);
  const OnnxTensor* re = find_named(g, "encoder.rel_embeddings.weight");
  if (!re || re->dims.size() != 2 || re->dims[1] != c.hidden)
    throw std::runtime_error("onnx: encoder.rel_embeddings.weight not found (relative_attention model expected)");
  c.buckets = (int)(re->dims[0] / 2);
  out->t["rel.emb"] = to_host(*re);
  load_layernorm(g, "/encoder", "rel.ln", c.hidden, out);

  // ---- layers
  int L = 0;
  while (find_node(g, "/layer." + std::to_string(L) + "/attention/self/query_proj/MatMul")) ++L;
  if (L == 0) throw std::runtime_error("onnx: no encoder layers found (node names without module scopes?)");
  c.layers = L;
  for (int l = 0; l < L; ++l) {
    std::string s = "/layer." + std::to_string(l), r = "layer." + std::to_string(l);
    load_linear(g, s + "/attention/self/query_proj", r + ".q", out);
    load_linear(g, s + "/attention/self/key_proj", r + ".k", out);
    load_linear(g, s + "/attention/self/value_proj", r + ".v", out);
    load_linear(g, s + "/attention/output/dense", r + ".o", out);
    load_layernorm(g, s + "/attention/output", r + ".ln1", c.hidden, out);
    load_linear(g, s + "/intermediate/dense", r + ".ffn1", out);
    load_linear(g, s + "/output/dense", r + ".ffn2", out);
    load_layernorm(g, s + "/output", r + ".ln2", c.hidden, out);
    // share_att_key: the position projections must reuse the content projections
    const OnnxNode* pq = find_node(g, s + "/attention/self/query_proj_1/MatMul", "MatMul");
    const OnnxNode* cq = find_node(g, s + "/attention/self/query_proj/MatMul", "MatMul");
    if (!pq) throw std::runtime_error("onnx: layer " + std::to_string(l) + " has no query_proj_1 (share_att_key p2c|c2p model expected)");
    // pos_att_type must be exactly {c2p, p2c}: the kernel's 1/sqrt(3d) scale and both bias terms are hard-wired (T:237-242)
    if (!find_node(g, s + "/attention/self/key_proj_1/MatMul", "MatMul"))
      throw std::runtime_error("onnx: layer " + std::to_string(l) + " has no key_proj_1: pos_att_type without c2p is not supported");
    if (pq->inputs.size() != 2 || cq->inputs.size() != 2) throw std::runtime_error("onnx: MatMul arity at " + pq->name);
    if (g.resolve(pq->inputs[1]) != g.resolve(cq->inputs[1])) {
      // not an alias: accept only if values are identical
      const OnnxTensor* a = g.resolve(pq->inputs[1]);
      const OnnxTensor* b = g.resolve(cq->inputs[1]);
      if (!a || !b || a->numel() != b->numel() || !a->raw || !b->raw || memcmp(a->raw, b->raw, a->raw_bytes) != 0)
        throw std::runtime_error("onnx: separate pos_query_proj weights are not supported (share_att_key=false)");
    }
  }
  c.inter = (int)out->at("layer.0.ffn1.w").dims[0];

  // ---- heads: Concat(batch, seq, Constant(heads), Constant(-1)) feeding the first Reshape
  c.heads = 0;
  for (auto& n : g.nodes) {
    if (n.op_type != "Concat" || n.inputs.size() != 4) continue;
    if (n.name.find("/layer.0/attention/self/") == std::string::npos) continue;
    int64_t hv = 0, last = 0;
    if (g.scalar_int(n.inputs[2], &hv) && g.scalar_int(n.inputs[3], &last) && last == -1 && hv > 0) {
      c.heads = (int)hv;
      break;
    }
  }
  if (c.heads == 0) c.heads = c.hidden / 64;   // d = 64 in every DeBERTa-v3 size
  if (c.hidden % c.heads != 0) throw std::runtime_error("onnx: hidden not divisible by heads");

  // ---- LayerNorm eps: Constant added to the variance in the embeddings LN
  if (const OnnxNode* n = find_node(g, "/embeddings/LayerNorm/Add", "Add")) {
    float e;
    for (auto& in : n->inputs)
      if (g.scalar_float(in, &e) && e > 0.f && e < 1e-2f) c.ln_eps = e;
  }

  // ---- log-bucket constants: Div(abs_pos, mid) -> Log -> Div(., log((max_pos-1)/mid))
  for (auto& n : g.nodes) {
    if (n.op_type != "Log" || n.inputs.empty() || n.outputs.empty()) continue;
    float mid = 0.f, lg = 0.f;
    auto pit = g.producer_of.find(n.inputs[0]);
    if (pit != g.producer_of.end() && g.nodes[pit->second].op_type == "Div" && g.nodes[pit->second].inputs.size() == 2 &&
        g.scalar_float(g.nodes[pit->second].inputs[1], &mid) && mid > 0.f) {
      for (auto& m : g.nodes)
        if (m.op_type == "Div" && m.inputs.size() == 2 && m.inputs[0] == n.outputs[0] &&
            g.scalar_float(m.inputs[1], &lg) && lg > 0.f) {
          int mp = (int)std::lround(std::exp((double)lg) * mid) + 1;
          if ((int)std::lround(mid) * 2 == c.buckets && mp > 1) c.max_rel_pos = mp;
        }
    }
    break;
  }

}

void load_head(const OnnxGraph& g, ModelWeights* out) {
  ModelConfig& c = out->cfg;
  // ---- class token id: Equal(input_ids, Constant)
  for (auto& n : g.nodes) {
    if (n.op_type != "Equal" || n.inputs.size() != 2) continue;
    int64_t v;
    if (n.inputs[0] == "input_ids" && g.scalar_int(n.inputs[1], &v)) { c.class_token = v; break; }
    if (n.inputs[1] == "input_ids" && g.scalar_int(n.inputs[0], &v)) { c.class_token = v; break; }
  }
  if (c.class_token < 0) throw std::runtime_error("onnx: Equal(input_ids, class_token_index) not found");

  // ---- encoder / head variants this engine does NOT implement: refuse them instead of returning wrong logits
  // (gliclass config: use_lstm, position_biased_input / token types / embed_proj of DeBERTa-v2, conv layer)
  for (auto& n : g.nodes) {
    if (n.op_type == "LSTM" || n.op_type == "GRU" || n.op_type == "RNN")
      throw std::runtime_error("onnx: recurrent layer (" + n.op_type + " at " + n.name + "): use_lstm models are not supported");
    if (n.op_type == "Conv" || n.op_type == "ConvTranspose")
      throw std::runtime_error("onnx: convolution (" + n.name + "): DeBERTa conv_kernel_size > 0 is not supported");
  }
  for (const char* bad : {"embeddings.position_embeddings.weight", "embeddings.token_type_embeddings.weight",
                          "embeddings.embed_proj.weight"})
    if (find_named(g, bad))
      throw std::runtime_error(std::string("onnx: initializer ") + bad + " present: position_biased_input / token types / "
                               "embedding_size != hidden_size are not supported");

  // ---- head
  load_linear(g, "/text_projector/linear_1", "text.1", out);
  load_linear(g, "/text_projector/linear_2", "text.2", out);
  load_linear(g, "/classes_projector/linear_1", "cls.1", out);
  load_linear(g, "/classes_projector/linear_2", "cls.2", out);
  c.head_hidden = (int)out->at("text.2.w").dims[0];
  int64_t Hh_ = c.head_hidden;
  // projector activation (gliclass config.projector_hidden_act): the nodes between linear_1 and linear_2 of a projector
  {
    auto act_of = [&](const std::string& scope) -> int {
      bool erf = false, relu = false, other = false;
      std::string what;
      for (auto& n : g.nodes) {
        if (n.name.find(scope) == std::string::npos) continue;
        if (n.name.find("/linear_1/") != std::string::npos || n.name.find("/linear_2/") != std::string::npos) continue;
        if (n.op_type == "Erf") erf = true;
        else if (n.op_type == "Relu") relu = true;
        else if (n.op_type == "Tanh" || n.op_type == "Sigmoid" || n.op_type == "Softplus" || n.op_type == "LeakyRelu" ||
                 n.op_type == "Elu" || n.op_type == "Selu" || n.op_type == "HardSigmoid" || n.op_type == "PRelu") {
          other = true;
          what = n.op_type;
        }
      }
      if (other || (erf && relu)) throw std::runtime_error("onnx: unsupported projector activation (" + what + ") under " + scope);
      if (erf) return 1;    // erf-GELU
      if (relu) return 2;   // ReLU
      throw std::runtime_error("onnx: no activation found between linear_1 and linear_2 of " + scope);
    };
    c.proj_act = act_of("/text_projector/");
    if (act_of("/classes_projector/") != c.proj_act)
      throw std::runtime_error("onnx: text and class projectors use different activations");
  }
  // embed_class_token = false: class rows are gathered one position AFTER each <<LABEL>> token.  In the traced graph that
  // is an Add(+1) on the NonZero-derived positions, outside the encoder scope.
  c.class_pos_offset = 0;
  for (auto& n : g.nodes) {
    if (n.op_type != "Add" || n.inputs.size() != 2 || n.name.find("/encoder_model/") != std::string::npos) continue;
    int64_t one = 0;
    int other = -1;
    if (g.scalar_int(n.inputs[1], &one) && one == 1) other = 0;
    else if (g.scalar_int(n.inputs[0], &one) && one == 1) other = 1;
    if (other < 0) continue;
    std::string v = n.inputs[other];
    for (int hop = 0; hop < 8; ++hop) {
      auto pit = g.producer_of.find(v);
      if (pit == g.producer_of.end()) break;
      const OnnxNode& pn = g.nodes[pit->second];
      if (pn.op_type == "NonZero") { c.class_pos_offset = 1; break; }
      if (pn.inputs.empty()) break;
      v = pn.inputs[0];
    }
    if (c.class_pos_offset) break;
  }
  // pooling strategy: whatever produces the input of text_projector.linear_1
  {
    const OnnxNode* l1 = find_node(g, "/text_projector/linear_1/Gemm", "Gemm");
    if (!l1) l1 = find_node(g, "/text_projector/linear_1/MatMul", "MatMul");
    const OnnxNode* src = nullptr;
    if (l1) {
      auto pit = g.producer_of.find(l1->inputs[0]);
      if (pit != g.producer_of.end()) src = &g.nodes[pit->second];
    }
    if (!src) throw std::runtime_error("onnx: cannot find what feeds text_projector.linear_1");
    int64_t gi = 0;
    if (src->op_type == "Gather" && src->inputs.size() == 2 && g.scalar_int(src->inputs[1], &gi) && (gi == 0 || gi == -1)) {
      c.pooling = gi == 0 ? POOL_FIRST : POOL_LAST;
    } else if (src->op_type == "ReduceMax") {
      c.pooling = POOL_MAX;
    } else if (src->op_type == "Div") {
      auto pit = g.producer_of.find(src->inputs[0]);
      if (pit == g.producer_of.end() || g.nodes[pit->second].op_type != "ReduceSum")
        throw std::runtime_error("onnx: unsupported pooling (Div not fed by ReduceSum) at " + src->name);
      c.pooling = POOL_AVG;
    } else {
      throw std::runtime_error("onnx: unsupported pooling strategy (" + src->op_type + " at " + src->name +
                               "); supported: first, last, avg, max");
    }
  }
  // feature normalisation: ReduceL2 -> Add(eps) -> Div on both features, logits * logit_scale
  {
    int n_l2 = 0;
    for (auto& n : g.nodes) {
      if (n.op_type != "ReduceL2") continue;
      ++n_l2;
      for (auto& m : g.nodes) {
        float e;
        if (m.op_type == "Add" && m.inputs.size() == 2 && m.inputs[0] == n.outputs[0] && g.scalar_float(m.inputs[1], &e) &&
            e >= 0.f && e < 1e-3f)
          c.norm_eps = e;
      }
    }
    const OnnxTensor* ls = find_named(g, "logit_scale");
    if (n_l2 == 2 && ls && ls->numel() == 1) {
      c.normalize = true;
      tensor_to_float(*ls, &c.logit_scale);
    } else if (n_l2 != 0 || ls) {
      throw std::runtime_error("onnx: unrecognised feature normalisation (ReduceL2 count " + std::to_string(n_l2) + ")");
    }
  }
  // scorer
  bool has_einsum = false;
  for (auto& n : g.nodes) if (n.op_type == "Einsum") has_einsum = true;
  if (find_node(g, "/scorer/mlp/mlp.0/MatMul") || find_node(g, "/scorer/mlp/mlp.0/Gemm")) {
    c.scorer = SCORER_MLP;
    load_linear(g, "/scorer/mlp/mlp.0", "scorer.mlp.0", out);
    load_linear(g, "/scorer/mlp/mlp.2", "scorer.mlp.2", out);
    load_linear(g, "/scorer/mlp/mlp.4", "scorer.mlp.4", out);
    c.mlp1 = (int)out->at("scorer.mlp.0.w").dims[0];
    c.mlp2 = (int)out->at("scorer.mlp.2.w").dims[0];
    if (out->at("scorer.mlp.0.w").dims[1] != 2 * Hh_ || out->at("scorer.mlp.2.w").dims[1] != c.mlp1 ||
        out->at("scorer.mlp.4.w").dims != std::vector<int64_t>{1, (int64_t)c.mlp2} || c.mlp1 % 8 || c.mlp2 % 8)
      throw std::runtime_error("onnx: unexpected MLP scorer shapes");
  } else if (find_node(g, "/scorer/proj_text/MatMul") || find_node(g, "/scorer/proj_text/Gemm")) {
    c.scorer = SCORER_WEIGHTED_DOT;
    load_linear(g, "/scorer/proj_text", "scorer.pt", out);
    load_linear(g, "/scorer/proj_label", "scorer.pl", out);
    load_linear(g, "/scorer/out_mlp/out_mlp.0", "scorer.o1", out);
    load_linear(g, "/scorer/out_mlp/out_mlp.3", "scorer.o2", out);
    if (out->at("scorer.pt.w").dims != std::vector<int64_t>{2 * Hh_, Hh_} ||
        out->at("scorer.pl.w").dims != std::vector<int64_t>{2 * Hh_, Hh_} ||
        out->at("scorer.o1.w").dims != std::vector<int64_t>{4 * Hh_, 3 * Hh_} ||
        out->at("scorer.o2.w").dims != std::vector<int64_t>{1, 4 * Hh_})
      throw std::runtime_error("onnx: unexpected weighted-dot scorer shapes");
  } else if (has_einsum) {
    c.scorer = SCORER_DOT;
  } else {
    throw std::runtime_error("onnx: scorer not recognised (expected Einsum dot, /scorer/mlp or /scorer/proj_text)");
  }

  // ---- shape checks
  auto expect = [&](const std::string& role, std::vector<int64_t> d) {
    if (out->at(role).dims != d) throw std::runtime_error("weights: unexpected shape for " + role);
  };
  int64_t H = c.hidden, I = c.inter, Hh = c.head_hidden;
  for (int l = 0; l < c.layers && c.backbone == BACKBONE_DEBERTA; ++l) {
    std::string r = "layer." + std::to_string(l);
    for (const char* p : {".q", ".k", ".v", ".o"}) { expect(r + p + ".w", {H, H}); expect(r + p + ".b", {H}); }
    expect(r + ".ffn1.w", {I, H}); expect(r + ".ffn1.b", {I});
    expect(r + ".ffn2.w", {H, I}); expect(r + ".ffn2.b", {H});
  }
  expect("text.1.w", {Hh, H}); expect("text.2.w", {Hh, Hh});
  expect("cls.1.w", {Hh, H}); expect("cls.2.w", {Hh, Hh});
  for (const char* r : {"text.1.b", "text.2.b", "cls.1.b", "cls.2.b"}) expect(r, {Hh});
  if (c.scorer == SCORER_MLP) {
    expect("scorer.mlp.0.b", {(int64_t)c.mlp1}); expect("scorer.mlp.2.b", {(int64_t)c.mlp2}); expect("scorer.mlp.4.b", {1});
  } else if (c.scorer == SCORER_WEIGHTED_DOT) {
    expect("scorer.pt.b", {2 * Hh}); expect("scorer.pl.b", {2 * Hh}); expect("scorer.o1.b", {4 * Hh}); expect("scorer.o2.b", {1});
  }
  if (c.backbone == BACKBONE_DEBERTA)
    for (const char* r : {"emb.ln.g", "emb.ln.b", "rel.ln.g", "rel.ln.b"}) expect(r, {H});
}

}  // namespace

void rel_index_table(int S, int buckets, int max_pos, int32_t* out) {
  // make_log_bucket_position in fp32, as traced (T:57-69): mid = buckets/2
  const int mid = buckets / 2;
  const float denom = logf((float)(max_pos - 1) / (float)mid);
  for (int delta = -(S - 1); delta <= S - 1; ++delta) {
    int sign = (delta > 0) - (delta < 0);
    int abs_pos = (delta < mid && delta > -mid) ? (mid - 1) : (delta < 0 ? -delta : delta);
    int bucket;
    if (abs_pos <= mid) {
      bucket = delta;
    } else {
      float lp = ceilf(logf((float)abs_pos / (float)mid) / denom * (float)(mid - 1)) + (float)mid;
      bucket = (int)lp * sign;
    }
    int idx = bucket + buckets;
    if (idx < 0) idx = 0;
    if (idx > 2 * buckets - 1) idx = 2 * buckets - 1;
    out[delta + S - 1] = idx;
  }
}

}  // namespace glc
