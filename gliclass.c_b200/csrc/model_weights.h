// Role mapping: ONNX graph -> named fp32 host tensors of the GLiClass uni-encoder
// (DeBERTa-v3 backbone + projector head with dot / MLP / weighted-dot scorer).  This is the "weights" half of what ORT's
// CreateSession does for the reference (src/model.c:269); the graph structure itself is not
// executed — the engine hard-codes the architecture.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <vector>

namespace glc {

struct HostTensor {
  std::vector<int64_t> dims;
  std::vector<float> data;
};

enum { BACKBONE_DEBERTA = 0, BACKBONE_QWEN2 = 1 };

struct ModelConfig {
  int backbone = BACKBONE_DEBERTA;   // which layer stack feeds the head (reference Readme.md:91-94 lists both families)
  int vocab = 0, hidden = 0, layers = 0, heads = 0, inter = 0;
  // decoder backbones (Qwen2 / Llama style): grouped-query attention, rotary embedding, RMSNorm, SwiGLU
  int kv_heads = 0, head_dim = 0;
  float rms_eps = 1e-6f;
  int head_hidden = 0;          // projector width
  int buckets = 256;            // position_buckets (= att_span); rel table has 2*buckets rows
  int max_rel_pos = 512;        // max_position in make_log_bucket_position
  float ln_eps = 1e-7f;
  int64_t class_token = -1;     // <<LABEL>> id: Constant feeding Equal(input_ids, .)
  // head variant, detected from the graph structure
  int pooling = 0;              // POOL_*: what feeds text_projector.linear_1
  int scorer = 0;               // SCORER_*
  bool normalize = false;       // ReduceL2 on both features + logits * logit_scale
  float norm_eps = 1e-8f;       // x / (|x| + eps)
  float logit_scale = 1.0f;
  int mlp1 = 0, mlp2 = 0;       // MLP scorer widths (cat[t,l] -> mlp1 -> mlp2 -> 1)
  int proj_act = 1;             // projector_hidden_act between linear_1 and linear_2: 1 erf-GELU, 2 ReLU (others are refused)
  int class_pos_offset = 0;     // embed_class_token=false: class rows are read one position after each <<LABEL>> token
};
enum { POOL_FIRST = 0, POOL_LAST = 1, POOL_AVG = 2, POOL_MAX = 3 };
enum { SCORER_DOT = 0, SCORER_MLP = 1, SCORER_WEIGHTED_DOT = 2 };

// Decoder-backbone roles (BACKBONE_QWEN2): emb.word [V,H]; layer.<l>.{q,k,v}.w [heads*d | kv*d, H] .b; layer.<l>.o.w
// [H, heads*d]; layer.<l>.{gate,up}.w [I,H]; layer.<l>.down.w [H,I]; layer.<l>.ln1.g / ln2.g [H] (RMSNorm); norm.g [H];
// rope.inv_freq [d/2].
// Linear weights are stored [out, in] row-major (torch Linear layout == the K-major "B"
// operand of the tcgen05 GEMM); ONNX MatMul initializers ([in, out]) are transposed on load.
struct ModelWeights {
  ModelConfig cfg;
  std::map<std::string, HostTensor> t;   // role -> tensor, roles listed in model_weights.cc
  const HostTensor& at(const std::string& role) const;
  bool has(const std::string& role) const { return t.count(role) != 0; }
};

// throws std::runtime_error with a precise message when a role cannot be found
void load_model_weights(const std::string& onnx_path, ModelWeights* out);

// idx[delta + S - 1] = clamp(bucket(delta) + buckets, 0, 2*buckets-1), delta in (-S, S)
// (HF make_log_bucket_position / build_relative_position)
void rel_index_table(int S, int buckets, int max_pos, int32_t* out /* 2S-1 */);

}  // namespace glc
