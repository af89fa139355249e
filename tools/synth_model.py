"""Synthetic workload builder: random-init GLiClass checkpoints of the named architectures exported
to model.onnx with the reference's own export call, and synthetic token batches.

This is the offline stand-in for the reference's model tooling (ONNX_CONVERTING/convert_to_onnx.py,
which downloads a trained checkpoint — impossible here: no network, no `gliclass` package) and for
its host-side tokenisation (src/preprocessor.c + src/tokenizer.c), producing the SAME file format
and tensor layouts the engine consumes.  It contains no forward arithmetic: the CPU oracle lives in
oracle/gliclass_oracle.py and imports the definitions below.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, asdict

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# architecture table (HF microsoft/deberta-v3-{small,base,large})
# --------------------------------------------------------------------------------------------


@dataclass
class ArchConfig:
    name: str
    vocab_size: int
    hidden_size: int
    num_layers: int
    num_heads: int
    intermediate_size: int
    position_buckets: int = 256
    max_relative_positions: int = 512      # config value -1 -> max_position_embeddings (T:158-160)
    layer_norm_eps: float = 1e-7
    class_token_index: int = 128001        # <<LABEL>>
    sep_token_index: int = 128002          # <<SEP>>
    head_hidden_size: int = 0              # GLiClass config.hidden_size; 0 -> same as encoder
    # head variants of the gliclass package (`M:` recalled, unpinned)
    pooling_strategy: str = "first"        # first | last | avg | max
    scorer_type: str = "simple"            # simple (dot) | weighted-dot | mlp
    normalize_features: bool = False       # L2-normalise text / class features, logits *= logit_scale
    mlp_hidden_size: int = 256             # MLPScorer: cat[t,l] -> 256 -> 128 -> 1
    embed_class_token: bool = True         # False: class rows are read one position AFTER each <<LABEL>> token
    projector_hidden_act: str = "gelu"     # gelu (erf) | relu
    # decoder backbones (reference Readme.md:91-94: gliclass-qwen-1.5B / gliclass-llama-1.3B; BASELINE.json configs[4])
    backbone: str = "deberta"              # deberta | qwen2
    num_kv_heads: int = 0                  # qwen2: grouped-query attention (0 -> num_heads)
    kv_head_dim: int = 0                   # qwen2: head dim when it is not hidden_size / num_heads
    rope_theta: float = 1.0e6
    rms_norm_eps: float = 1.0e-6

    def __post_init__(self):
        if self.head_hidden_size == 0:
            self.head_hidden_size = self.hidden_size
        assert self.pooling_strategy in ("first", "last", "avg", "max"), self.pooling_strategy
        assert self.scorer_type in ("simple", "weighted-dot", "mlp"), self.scorer_type
        assert self.projector_hidden_act in ("gelu", "relu", "tanh"), self.projector_hidden_act
        assert self.backbone in ("deberta", "qwen2"), self.backbone
        if self.num_kv_heads == 0:
            self.num_kv_heads = self.num_heads

    @property
    def head_dim(self) -> int:
        return self.kv_head_dim or self.hidden_size // self.num_heads


ARCHS = {
    # unit-test scale; two heads of d=64 so the real kernels (d=64 only) run it
    "tiny": dict(vocab_size=1027, hidden_size=128, num_layers=2, num_heads=2, intermediate_size=512,
                 class_token_index=1025, sep_token_index=1026),
    "mini": dict(vocab_size=2051, hidden_size=256, num_layers=3, num_heads=4, intermediate_size=1024,
                 class_token_index=2049, sep_token_index=2050),
    "small": dict(vocab_size=128003, hidden_size=768, num_layers=6, num_heads=12, intermediate_size=3072),
    "base": dict(vocab_size=128003, hidden_size=768, num_layers=12, num_heads=12, intermediate_size=3072),
    "large": dict(vocab_size=128003, hidden_size=1024, num_layers=24, num_heads=16, intermediate_size=4096),
    # decoder backbones: Qwen2 (RMSNorm, RoPE theta 1e6, GQA, SwiGLU, QKV bias), bidirectional attention (LLM2Vec style)
    "qwen-mini": dict(backbone="qwen2", vocab_size=2051, hidden_size=512, num_layers=3, num_heads=4, num_kv_heads=2,
                      kv_head_dim=128, intermediate_size=1536, class_token_index=2049, sep_token_index=2050),
    # the 1.5B layer geometry at 4 layers and a small vocabulary (< 2 GB: inline weights, quick to export) for the default suite
    "qwen1.5b-4l": dict(backbone="qwen2", vocab_size=32003, hidden_size=1536, num_layers=4, num_heads=12, num_kv_heads=2,
                        kv_head_dim=128, intermediate_size=8960, class_token_index=32001, sep_token_index=32002),
    "qwen1.5b": dict(backbone="qwen2", vocab_size=151938, hidden_size=1536, num_layers=28, num_heads=12, num_kv_heads=2,
                     kv_head_dim=128, intermediate_size=8960, class_token_index=151936, sep_token_index=151937),
}


def make_config(arch: str, **over) -> ArchConfig:
    d = dict(ARCHS[arch])
    d.update(over)
    return ArchConfig(name=arch, **d)


# --------------------------------------------------------------------------------------------
# deterministic non-degenerate init (every tensor random so that dedup in the
# exporter cannot merge them and bias / gamma bugs are observable)
# --------------------------------------------------------------------------------------------

ENC = "model.encoder_model."


def init_weights(cfg: ArchConfig, seed: int = 0) -> dict[str, torch.Tensor]:
    """Fan-in scaled Gaussian init, every tensor random.

    Scales are chosen so that the random model behaves like a trained one where it matters for
    parity: attention scores have std ~2 (peaked softmax, so rel-pos bias errors are visible), the
    attention / FFN branches are comparable to the residual, and logits are O(1) and straddle the
    sigmoid threshold (HF's default std=0.02 gives three identical logits).  Larger
    gains (q/k 1.8, v/o/ffn 1.4) put a random net in a chaotic regime where even rounding the
    WEIGHTS to 16 bits moves logits by several 1e-2 (scripts/emulate_precision.py, DESIGN.md
    "Numerics"); that says nothing about a kernel, so the fixtures stay out of it.
    """
    if cfg.backbone == "qwen2":
        return init_weights_qwen2(cfg, seed)
    g = torch.Generator().manual_seed(seed)
    H, I, Hh = cfg.hidden_size, cfg.intermediate_size, cfg.head_hidden_size

    def n(*shape, std=0.05, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).float()

    w: dict[str, torch.Tensor] = {}
    w[ENC + "embeddings.word_embeddings.weight"] = n(cfg.vocab_size, H, std=0.5)
    w[ENC + "embeddings.LayerNorm.weight"] = n(H, std=0.1, mean=1.0)
    w[ENC + "embeddings.LayerNorm.bias"] = n(H, std=0.02)
    w[ENC + "encoder.rel_embeddings.weight"] = n(2 * cfg.position_buckets, H, std=0.5)
    w[ENC + "encoder.LayerNorm.weight"] = n(H, std=0.1, mean=1.0)
    w[ENC + "encoder.LayerNorm.bias"] = n(H, std=0.02)
    for l in range(cfg.num_layers):
        p = f"{ENC}encoder.layer.{l}."
        for nm, s in (("query_proj", 1.4), ("key_proj", 1.4), ("value_proj", 1.0)):
            w[p + f"attention.self.{nm}.weight"] = n(H, H, std=s / math.sqrt(H))
            w[p + f"attention.self.{nm}.bias"] = n(H, std=0.02)
        w[p + "attention.output.dense.weight"] = n(H, H, std=1.0 / math.sqrt(H))
        w[p + "attention.output.dense.bias"] = n(H, std=0.02)
        w[p + "attention.output.LayerNorm.weight"] = n(H, std=0.1, mean=1.0)
        w[p + "attention.output.LayerNorm.bias"] = n(H, std=0.02)
        w[p + "intermediate.dense.weight"] = n(I, H, std=1.0 / math.sqrt(H))
        w[p + "intermediate.dense.bias"] = n(I, std=0.02)
        w[p + "output.dense.weight"] = n(H, I, std=0.7 / math.sqrt(I))
        w[p + "output.dense.bias"] = n(H, std=0.02)
        w[p + "output.LayerNorm.weight"] = n(H, std=0.1, mean=1.0)
        w[p + "output.LayerNorm.bias"] = n(H, std=0.02)
    # head: two FeaturesProjectors (Linear-GELU-Linear); linear_2 scaled for O(1) logits.
    s2 = 1.2 / math.sqrt(Hh) / (Hh ** 0.25)
    for pj in ("text_projector", "classes_projector"):
        w[f"model.{pj}.linear_1.weight"] = n(Hh, H, std=1.4 / math.sqrt(H))
        w[f"model.{pj}.linear_1.bias"] = n(Hh, std=0.02)
        w[f"model.{pj}.linear_2.weight"] = n(Hh, Hh, std=s2)
        w[f"model.{pj}.linear_2.bias"] = n(Hh, std=0.02)
    # head variants (drawn AFTER everything above so the default checkpoints are unchanged)
    # feature scale entering the scorer: ~1/sqrt(Hh) per element after normalisation, else s2*sqrt(Hh)
    fs = (1.0 / math.sqrt(Hh)) if cfg.normalize_features else (s2 * math.sqrt(Hh) * 0.8)
    if cfg.normalize_features:
        w["model.logit_scale"] = torch.tensor(2.6592) + n(1, std=0.05)[0]
    if cfg.scorer_type == "mlp":
        m1, m2 = cfg.mlp_hidden_size, cfg.mlp_hidden_size // 2
        w["model.scorer.mlp.0.weight"] = n(m1, 2 * Hh, std=1.0 / (fs * math.sqrt(2 * Hh)))
        w["model.scorer.mlp.0.bias"] = n(m1, std=0.05)
        w["model.scorer.mlp.2.weight"] = n(m2, m1, std=1.6 / math.sqrt(m1))
        w["model.scorer.mlp.2.bias"] = n(m2, std=0.05)
        w["model.scorer.mlp.4.weight"] = n(1, m2, std=2.0 / math.sqrt(m2))
        w["model.scorer.mlp.4.bias"] = n(1, std=0.05)
    elif cfg.scorer_type == "weighted-dot":
        for nm in ("proj_text", "proj_label"):
            w[f"model.scorer.{nm}.weight"] = n(2 * Hh, Hh, std=1.0 / (fs * math.sqrt(Hh)))
            w[f"model.scorer.{nm}.bias"] = n(2 * Hh, std=0.05)
        w["model.scorer.out_mlp.0.weight"] = n(4 * Hh, 3 * Hh, std=1.0 / math.sqrt(3 * Hh))
        w["model.scorer.out_mlp.0.bias"] = n(4 * Hh, std=0.05)
        w["model.scorer.out_mlp.3.weight"] = n(1, 4 * Hh, std=2.0 / math.sqrt(4 * Hh))
        w["model.scorer.out_mlp.3.bias"] = n(1, std=0.05)
    if cfg.num_layers >= DEEP_INIT_MIN_LAYERS:
        _trained_like_structure(w, cfg)
    return w


# Deep random post-LN stacks collapse: every layer adds a token-independent component (the mean of
# GELU through W2, near-uniform attention averages), so after 12-24 layers all positions hold almost
# the same vector, the <<LABEL>> rows (one token id!) become identical and a row's logits span 1e-2
# on one side of the threshold — decision parity is then vacuous and a label-position bug passes
# (SURVEY.md H2).  A trained GLiClass tells its labels apart because attention at a
# <<LABEL>> position reads the label-name tokens that follow it.  The three deterministic edits below
# give a random net that property without leaving the well-conditioned regime (no gain above 1.4):
#   1. residual branches at half gain (out-proj and FFN2 weights x 0.5): a layer perturbs a token, it
#      does not replace it, so token identity survives the depth;
#   2. FFN2 rows centred (zero sum over the intermediate units): the mean of GELU no longer adds the
#      same vector to every position;
#   3. "local" attention: rows delta = -1, -2, -3 of rel_embeddings get a common direction v and every
#      head's query bias gets the matching direction, so Q_i . posK[idx(i - j)] carries a constant
#      +LOCAL_LOGIT softmax logit for the three tokens after position i (c2p term, T:313-324) — peaked,
#      position-specific attention, which is also what makes a rel-pos indexing bug visible.
# Applied to the >= 6-layer architectures (small / base / large); tiny / mini keep the plain init their
# committed fixtures were generated with.
DEEP_INIT_MIN_LAYERS = 6
LOCAL_LOGIT = 8.0
LOCAL_OFFSETS = (-1, -2, -3)
BRANCH_GAIN = 0.5


def _trained_like_structure(w: dict, cfg: ArchConfig) -> None:
    g = torch.Generator().manual_seed(977)
    H, d, span = cfg.hidden_size, cfg.head_dim, cfg.position_buckets
    v = torch.randn(H, generator=g)
    v = v - v.mean()
    v = v / v.norm()
    rel = w[ENC + "encoder.rel_embeddings.weight"]
    rows = [o + span for o in LOCAL_OFFSETS]
    amp = rel[0].norm()
    for r in rows:
        rel[r] += amp * v
    gam, bet = w[ENC + "encoder.LayerNorm.weight"], w[ENC + "encoder.LayerNorm.bias"]
    mu = rel.mean(-1, keepdim=True)
    var = ((rel - mu) ** 2).mean(-1, keepdim=True)
    relln = (rel - mu) / torch.sqrt(var + cfg.layer_norm_eps) * gam + bet
    scale = math.sqrt(3 * d)
    for l in range(cfg.num_layers):
        p = f"{ENC}encoder.layer.{l}."
        w[p + "attention.output.dense.weight"] *= BRANCH_GAIN
        w2 = w[p + "output.dense.weight"]
        w2 -= w2.mean(dim=1, keepdim=True)
        w2 *= BRANCH_GAIN
        Wk, bk = w[p + "attention.self.key_proj.weight"], w[p + "attention.self.key_proj.bias"]
        bq = w[p + "attention.self.query_proj.bias"]
        posk = relln[: 2 * span] @ Wk.T + bk
        for hh in range(cfg.num_heads):
            sl = slice(hh * d, (hh + 1) * d)
            u = posk[rows][:, sl].mean(0) - posk[:, sl].mean(0)
            bq[sl] += u / (u.norm() ** 2) * LOCAL_LOGIT * scale


def _init_head(w: dict, cfg: ArchConfig, n) -> None:
    """the GLiClass head tensors (same draws as in init_weights, shared with the decoder backbones)"""
    H, Hh = cfg.hidden_size, cfg.head_hidden_size
    s2 = 1.2 / math.sqrt(Hh) / (Hh ** 0.25)
    for pj in ("text_projector", "classes_projector"):
        w[f"model.{pj}.linear_1.weight"] = n(Hh, H, std=1.4 / math.sqrt(H))
        w[f"model.{pj}.linear_1.bias"] = n(Hh, std=0.02)
        w[f"model.{pj}.linear_2.weight"] = n(Hh, Hh, std=s2)
        w[f"model.{pj}.linear_2.bias"] = n(Hh, std=0.02)
    if cfg.normalize_features:
        w["model.logit_scale"] = torch.tensor(2.6592) + n(1, std=0.05)[0]
    assert cfg.scorer_type == "simple", "decoder-backbone fixtures use the dot scorer"


QWEN_LOCAL_PAIRS = 16      # rotary pairs (the highest frequencies) that carry the local-attention bias
QWEN_LOCAL_LOGIT = 8.0


def init_weights_qwen2(cfg: ArchConfig, seed: int = 0) -> dict[str, torch.Tensor]:
    """Random-init Qwen2 backbone (HF Qwen2Model parameter names under model.encoder_model.) + GLiClass head.

    Pre-norm residual stream: the branch outputs (o_proj, down_proj) are scaled by 1/sqrt(2L) so the stream stays O(1)
    over the depth.  As for the DeBERTa fixtures, a random decoder averages over all keys and cannot tell two <<LABEL>>
    tokens apart; here the q / k biases of every head get a common component on the QWEN_LOCAL_PAIRS highest-frequency
    rotary pairs, so that (R_i b_q).(R_j b_k) = beta^2 sum_p cos((i - j) theta_p) peaks at |i - j| <= 3 with a softmax
    logit of QWEN_LOCAL_LOGIT: position-specific, peaked attention through the RoPE path itself."""
    g = torch.Generator().manual_seed(seed)
    H, I, L = cfg.hidden_size, cfg.intermediate_size, cfg.num_layers
    d, nh, nkv = cfg.head_dim, cfg.num_heads, cfg.num_kv_heads

    def n(*shape, std=0.05, mean=0.0):
        return (torch.randn(*shape, generator=g) * std + mean).float()

    w: dict[str, torch.Tensor] = {}
    w[ENC + "embed_tokens.weight"] = n(cfg.vocab_size, H, std=1.0)
    branch = 1.0 / math.sqrt(2.0 * L)
    beta = math.sqrt(QWEN_LOCAL_LOGIT * math.sqrt(d) / QWEN_LOCAL_PAIRS)
    for l in range(L):
        p = f"{ENC}layers.{l}."
        w[p + "input_layernorm.weight"] = n(H, std=0.1, mean=1.0)
        w[p + "post_attention_layernorm.weight"] = n(H, std=0.1, mean=1.0)
        w[p + "self_attn.q_proj.weight"] = n(nh * d, H, std=1.2 / math.sqrt(H))
        w[p + "self_attn.k_proj.weight"] = n(nkv * d, H, std=1.2 / math.sqrt(H))
        w[p + "self_attn.v_proj.weight"] = n(nkv * d, H, std=1.0 / math.sqrt(H))
        bq, bk = n(nh * d, std=0.02), n(nkv * d, std=0.02)
        bq.view(nh, d)[:, :QWEN_LOCAL_PAIRS] += beta
        bk.view(nkv, d)[:, :QWEN_LOCAL_PAIRS] += beta
        w[p + "self_attn.q_proj.bias"] = bq
        w[p + "self_attn.k_proj.bias"] = bk
        w[p + "self_attn.v_proj.bias"] = n(nkv * d, std=0.02)
        w[p + "self_attn.o_proj.weight"] = n(H, nh * d, std=branch * 2.0 / math.sqrt(nh * d))
        w[p + "mlp.gate_proj.weight"] = n(I, H, std=1.0 / math.sqrt(H))
        w[p + "mlp.up_proj.weight"] = n(I, H, std=1.0 / math.sqrt(H))
        w[p + "mlp.down_proj.weight"] = n(H, I, std=branch * 2.0 / math.sqrt(I))
    w[ENC + "norm.weight"] = n(H, std=0.1, mean=1.0)
    _init_head(w, cfg, n)
    return w


# --------------------------------------------------------------------------------------------
# synthetic inputs (layout of reference src/preprocessor.c:96-108 with
# prompt_first=false and src/tokenizer.c:44-84 pad-to-longest, pad id 0 / mask 0)
# --------------------------------------------------------------------------------------------


def synth_inputs(cfg: ArchConfig, B: int, S: int, n_labels, seed: int, ragged: bool = False,
                 min_frac: float = 0.25):
    """Returns (input_ids, attention_mask) int64 [B,S].

    n_labels: int or list of per-row label counts.  Row layout: [CLS]=1, text tokens uniform in
    [3, text_hi), then per label <<LABEL>> + 2 tokens, then <<SEP>>, then [SEP]=2; if ragged the row
    length is uniform in [S*min_frac, S] and the tail is padded with id 0 / mask 0.
    """
    g = torch.Generator().manual_seed(seed)
    text_hi = min(cfg.class_token_index, 128000)
    if isinstance(n_labels, int):
        n_labels = [n_labels] * B
    ids = torch.zeros(B, S, dtype=torch.long)
    mask = torch.zeros(B, S, dtype=torch.long)
    for b in range(B):
        nl = n_labels[b]
        tail = 3 * nl + 2
        L = S
        if ragged:
            lo = max(int(S * min_frac), tail + 2)
            L = int(torch.randint(lo, S + 1, (1,), generator=g).item())
        row = torch.randint(3, text_hi, (L,), generator=g)
        row[0] = 1
        p = L - tail
        for c in range(nl):
            row[p + 3 * c] = cfg.class_token_index
        row[L - 2] = cfg.sep_token_index
        row[L - 1] = 2
        ids[b, :L] = row
        mask[b, :L] = 1
    return ids, mask


# --------------------------------------------------------------------------------------------
# the traced module + export (reference ONNX_CONVERTING/convert_to_onnx.py:62-79)
# --------------------------------------------------------------------------------------------


def build_hf_module(cfg: ArchConfig, w: dict):
    """GLiClassModel-shaped nn.Module around transformers.DebertaV2Model, loaded with `w`.

    Attribute names mirror the gliclass package (model.encoder_model / text_projector /
    classes_projector) so exported node / initializer names carry the same scopes.
    """
    import torch.nn as nn

    if cfg.backbone == "qwen2":
        return _build_hf_module_qwen2(cfg, w)
    from transformers import DebertaV2Config, DebertaV2Model

    hf_cfg = DebertaV2Config(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_layers,
        num_attention_heads=cfg.num_heads, intermediate_size=cfg.intermediate_size, hidden_act="gelu",
        relative_attention=True, position_buckets=cfg.position_buckets, norm_rel_ebd="layer_norm",
        share_att_key=True, pos_att_type=["p2c", "c2p"], position_biased_input=False, type_vocab_size=0,
        max_relative_positions=-1, max_position_embeddings=cfg.max_relative_positions,
        layer_norm_eps=cfg.layer_norm_eps, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0,
        pad_token_id=0)

    class FeaturesProjector(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear_1 = nn.Linear(cfg.hidden_size, cfg.head_hidden_size)
            self.linear_2 = nn.Linear(cfg.head_hidden_size, cfg.head_hidden_size)

        def forward(self, t):
            act = {"gelu": F.gelu, "relu": F.relu, "tanh": torch.tanh}[cfg.projector_hidden_act]
            return self.linear_2(act(self.linear_1(t)))

    Hh = cfg.head_hidden_size

    class Pooler(nn.Module):   # gliclass poolings.py (`M:`): first / last token, masked mean, masked max
        def forward(self, hs, attention_mask):
            if cfg.pooling_strategy == "first":
                return hs[:, 0, :]
            if cfg.pooling_strategy == "last":
                return hs[:, -1, :]
            m = attention_mask.unsqueeze(-1).to(hs.dtype)
            if cfg.pooling_strategy == "avg":
                return (hs * m).sum(dim=1) / m.sum(dim=1)
            return hs.masked_fill(m == 0, torch.finfo(hs.dtype).min).max(dim=1)[0]

    class ScorerDot(nn.Module):
        def forward(self, t, l):
            return torch.einsum("BD,BCD->BC", t, l)

    class MLPScorer(nn.Module):   # cat[t, l] -> Linear-ReLU-Linear-ReLU-Linear(1)
        def __init__(self):
            super().__init__()
            m1, m2 = cfg.mlp_hidden_size, cfg.mlp_hidden_size // 2
            self.mlp = nn.Sequential(nn.Linear(2 * Hh, m1), nn.ReLU(), nn.Linear(m1, m2), nn.ReLU(), nn.Linear(m2, 1))

        def forward(self, t, l):
            te = t.unsqueeze(1).expand(-1, l.shape[1], -1)
            return self.mlp(torch.cat([te, l], dim=-1)).squeeze(-1)

    class ScorerWeightedDot(nn.Module):   # proj -> (.., Hh, 2) halves -> cat[t0, l0, t1*l1] -> Linear-ReLU-Linear(1)
        def __init__(self):
            super().__init__()
            self.proj_text = nn.Linear(Hh, 2 * Hh)
            self.proj_label = nn.Linear(Hh, 2 * Hh)
            self.out_mlp = nn.Sequential(nn.Linear(3 * Hh, 4 * Hh), nn.Dropout(0.0), nn.ReLU(), nn.Linear(4 * Hh, 1))

        def forward(self, t, l):
            B, C = l.shape[0], l.shape[1]
            lr = self.proj_label(l).view(B, C, -1, 2)
            tr = self.proj_text(t).view(B, 1, -1, 2).expand(-1, C, -1, -1)
            cat = torch.cat([tr[..., 0], lr[..., 0], tr[..., 1] * lr[..., 1]], dim=-1)
            return self.out_mlp(cat).squeeze(-1)

    class UniEncoder(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder_model = DebertaV2Model(hf_cfg)
            self.text_projector = FeaturesProjector()
            self.classes_projector = FeaturesProjector()
            self.pooler = Pooler()
            self.scorer = {"simple": ScorerDot, "mlp": MLPScorer, "weighted-dot": ScorerWeightedDot}[cfg.scorer_type]()
            if cfg.normalize_features:
                self.logit_scale = nn.Parameter(torch.tensor(2.6592))

        def forward(self, input_ids, attention_mask):
            hs = self.encoder_model(input_ids, attention_mask=attention_mask)[0]
            B, S, D = hs.shape
            class_token_mask = input_ids == cfg.class_token_index
            num_class_tokens = torch.sum(class_token_mask, dim=-1, keepdim=True)
            max_c = num_class_tokens.max()
            ar = torch.arange(max_c, dtype=attention_mask.dtype).unsqueeze(0).expand(B, -1)
            batch_idx, target_idx = torch.where(ar < num_class_tokens)
            bi_cls, pos_cls = torch.where(class_token_mask)
            if not cfg.embed_class_token:
                pos_cls = pos_cls + 1
            cls = torch.zeros(B, max_c, D, dtype=hs.dtype)
            cls[batch_idx, target_idx] = hs[bi_cls, pos_cls]
            pooled = self.text_projector(self.pooler(hs, attention_mask))
            cls = self.classes_projector(cls)
            if cfg.normalize_features:
                pooled = pooled / (pooled.norm(p=2, dim=-1, keepdim=True) + 1e-8)
                cls = cls / (cls.norm(p=2, dim=-1, keepdim=True) + 1e-8)
            logits = self.scorer(pooled, cls)
            if cfg.normalize_features:
                logits = logits * self.logit_scale
            return logits

    class GLiClassModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = UniEncoder()

        def forward(self, input_ids, attention_mask):
            return self.model(input_ids, attention_mask)

    m = GLiClassModel().eval()
    sd = m.state_dict()
    missing = [k for k in sd if k not in w and "position_ids" not in k]
    assert not missing, missing
    m.load_state_dict({k: v for k, v in w.items()}, strict=False)
    return m


def _build_hf_module_qwen2(cfg: ArchConfig, w: dict):
    """GLiClassModel-shaped module around transformers.Qwen2Model with BIDIRECTIONAL attention (the causal mask is replaced
    by a key-padding mask, as LLM2Vec-style encoders built on decoder checkpoints do; `M:` for the gliclass package)."""
    from transformers import Qwen2Config, Qwen2Model
    import torch.nn as nn

    hf_cfg = Qwen2Config(
        vocab_size=cfg.vocab_size, hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_layers,
        num_attention_heads=cfg.num_heads, num_key_value_heads=cfg.num_kv_heads, head_dim=cfg.head_dim,
        intermediate_size=cfg.intermediate_size, hidden_act="silu", rms_norm_eps=cfg.rms_norm_eps,
        rope_parameters={"rope_type": "default", "rope_theta": cfg.rope_theta}, max_position_embeddings=4096,
        attention_dropout=0.0, use_sliding_window=False, tie_word_embeddings=False, pad_token_id=0)
    hf_cfg._attn_implementation = "eager"
    Hh = cfg.head_hidden_size

    class FeaturesProjector(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear_1 = nn.Linear(cfg.hidden_size, Hh)
            self.linear_2 = nn.Linear(Hh, Hh)

        def forward(self, t):
            act = {"gelu": F.gelu, "relu": F.relu, "tanh": torch.tanh}[cfg.projector_hidden_act]
            return self.linear_2(act(self.linear_1(t)))

    class UniEncoder(nn.Module):
        def __init__(self):
            super().__init__()
            self.encoder_model = Qwen2Model(hf_cfg)
            self.text_projector = FeaturesProjector()
            self.classes_projector = FeaturesProjector()
            if cfg.normalize_features:
                self.logit_scale = nn.Parameter(torch.tensor(2.6592))

        def forward(self, input_ids, attention_mask):
            B, S = input_ids.shape
            # bidirectional: additive [B,1,S,S] mask, 0 where the KEY is valid, finfo.min where it is padding
            neg = torch.finfo(torch.float32).min
            add = (1.0 - attention_mask.to(torch.float32))[:, None, None, :] * neg
            add = add.expand(B, 1, S, S)
            hs = self.encoder_model(input_ids, attention_mask={"full_attention": add})[0]
            D = hs.shape[-1]
            class_token_mask = input_ids == cfg.class_token_index
            num_class_tokens = torch.sum(class_token_mask, dim=-1, keepdim=True)
            max_c = num_class_tokens.max()
            ar = torch.arange(max_c, dtype=attention_mask.dtype).unsqueeze(0).expand(B, -1)
            batch_idx, target_idx = torch.where(ar < num_class_tokens)
            bi_cls, pos_cls = torch.where(class_token_mask)
            if not cfg.embed_class_token:
                pos_cls = pos_cls + 1
            cls = torch.zeros(B, max_c, D, dtype=hs.dtype)
            cls[batch_idx, target_idx] = hs[bi_cls, pos_cls]
            if cfg.pooling_strategy == "first":
                pooled = hs[:, 0, :]
            elif cfg.pooling_strategy == "last":
                pooled = hs[:, -1, :]
            else:
                m = attention_mask.unsqueeze(-1).to(hs.dtype)
                pooled = (hs * m).sum(dim=1) / m.sum(dim=1) if cfg.pooling_strategy == "avg" else \
                    hs.masked_fill(m == 0, torch.finfo(hs.dtype).min).max(dim=1)[0]
            pooled = self.text_projector(pooled)
            cls = self.classes_projector(cls)
            if cfg.normalize_features:
                pooled = pooled / (pooled.norm(p=2, dim=-1, keepdim=True) + 1e-8)
                cls = cls / (cls.norm(p=2, dim=-1, keepdim=True) + 1e-8)
            logits = torch.einsum("BD,BCD->BC", pooled, cls)
            if cfg.normalize_features:
                logits = logits * self.logit_scale
            return logits

    class GLiClassModel(nn.Module):
        def __init__(self):
            super().__init__()
            self.model = UniEncoder()

        def forward(self, input_ids, attention_mask):
            return self.model(input_ids, attention_mask)

    m = GLiClassModel().eval()
    sd = m.state_dict()
    missing = [k for k in sd if k not in w and "inv_freq" not in k]
    assert not missing, missing
    m.load_state_dict({k: v for k, v in w.items()}, strict=False)
    return m


def export_onnx(module, cfg: ArchConfig, path: str, S: int = 24, n_labels: int = 3) -> None:
    """The reference's export call (convert_to_onnx.py:62-79), opset 14, legacy TorchScript path.

    The final onnxscript-function splice needs the `onnx` package (absent here); it is a no-op for
    this graph, so it is bypassed.
    """
    import warnings
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    ids, mask = synth_inputs(cfg, 2, S, [n_labels, max(1, n_labels - 1)], seed=7, ragged=True, min_frac=0.7)
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.onnx.export(
            module, (ids, mask), path,
            input_names=["input_ids", "attention_mask"], output_names=["logits"],
            dynamic_axes={"input_ids": {0: "batch_size", 1: "sequence_length"},
                          "attention_mask": {0: "batch_size", 1: "sequence_length"},
                          "logits": {0: "position", 1: "batch_size"}},
            opset_version=14, dynamo=False)


def make_model_file(arch: str, path: str, seed: int = 0, **over):
    """weights -> HF module -> model.onnx at `path`.  Returns (cfg, weights)."""
    cfg = make_config(arch, **over)
    w = init_weights(cfg, seed)
    if not os.path.exists(path):
        m = build_hf_module(cfg, w)
        tmp = path + ".tmp%d" % os.getpid()
        export_onnx(m, cfg, tmp)
        os.replace(tmp, path)
    return cfg, w


def flops_per_text(cfg: ArchConfig, S: int, C: int) -> float:
    """Algorithmic FLOPs per text: pos projections hoisted and excluded."""
    Hh = cfg.head_hidden_size
    if cfg.backbone == "qwen2":
        # L (2 S H (2 H + 2 H_kv) + 6 S H I + 4 S^2 H) with H = heads * head_dim for the attention terms
        L, H, I = cfg.num_layers, cfg.hidden_size, cfg.intermediate_size
        Hq, Hkv = cfg.num_heads * cfg.head_dim, cfg.num_kv_heads * cfg.head_dim
        return (L * (2.0 * S * H * (2 * Hq + 2 * Hkv) + 6.0 * S * H * I + 4.0 * S * S * Hq) +
                (1 + C) * (2.0 * H * Hh + 2.0 * Hh * Hh) + 2.0 * C * Hh)
    L, H, R = cfg.num_layers, cfg.hidden_size, 2 * cfg.position_buckets
    return L * (24.0 * S * H * H + 4.0 * S * S * H + 4.0 * S * R * H) + (1 + C) * (2.0 * H * Hh + 2.0 * Hh * Hh) + 2.0 * C * Hh


def config_dict(cfg: ArchConfig) -> dict:
    return asdict(cfg)
