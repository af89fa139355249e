"""CPU oracle for the GLiClass hot path (TEST INFRASTRUCTURE — never imported by the product path).

What this restates
------------------
The arithmetic that GLiClass.c hands to ONNX Runtime inside ``run_inference`` →
``g_ort->Run`` (reference ``src/model.c:173-182``): the DeBERTa‑v3 encoder forward traced by
``ONNX_CONVERTING/convert_to_onnx.py:71-79`` plus the GLiClass uni‑encoder head.  The arithmetic
itself lives in third‑party code that is *not* in ``/root/reference``:

* ``transformers`` DeBERTa‑v2 (installed here, 5.5.0; ``T:`` = models/deberta_v2/modeling_deberta_v2.py),
* the ``gliclass`` PyPI package (NOT installed; the head below is restated from its published
  structure — FeaturesProjector / first‑token pooler / dot scorer),
* ONNX Runtime 1.19.2 (NOT installed; it only executes the traced fp32 graph).

Parity pin status
-----------------
* encoder: pinned against ``transformers.DebertaV2Model`` (tests/test_oracle.py runs both on the
  same weights; fixtures in tests/golden/ were generated with ``oracle/make_golden.py``).
* head + end‑to‑end logits: **parity unpinned** — the reference's only golden vector
  (``original_logits`` in the HF‑hosted onnx/config.json, ``ONNX_CONVERTING/test_onnx.py:25-31``) is
  unreachable offline and neither ``gliclass`` nor ``onnxruntime`` can be imported here.

Everything is fp32 on CPU, written with plain tensor ops so that each line can be checked against
the cited source.  Only tests/, ``__graft_entry__.smoke()`` and bench.py's cpu_baseline / ``--impl
reference`` legs may import this module.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, asdict

import numpy as np
import torch
import torch.nn.functional as F

# architecture table, deterministic init, synthetic inputs and the ONNX export live in
# tools/synth_model.py (shared with bench.py, which must not import the oracle for its GPU arm)
import sys as _sys

_sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
from synth_model import (ARCHS, ENC, ArchConfig, build_hf_module, config_dict, export_onnx, flops_per_text,  # noqa: E402,F401
                         init_weights, make_config, make_model_file, synth_inputs)

# --------------------------------------------------------------------------------------------
# relative-position buckets  (T:57-69 make_log_bucket_position, T:72-101 build_relative_position)
# --------------------------------------------------------------------------------------------


def log_bucket(rel: np.ndarray, bucket_size: int, max_position: int) -> np.ndarray:
    """bucket(r) with the reference's fp32 arithmetic (torch.log/ceil on float32)."""
    rel_t = torch.as_tensor(rel, dtype=torch.long)
    sign = torch.sign(rel_t)
    mid = bucket_size // 2
    abs_pos = torch.where((rel_t < mid) & (rel_t > -mid), torch.tensor(mid - 1).type_as(rel_t), torch.abs(rel_t))
    log_pos = torch.ceil(torch.log(abs_pos / mid) / torch.log(torch.tensor((max_position - 1) / mid)) * (mid - 1)) + mid
    bucket_pos = torch.where(abs_pos <= mid, rel_t.type_as(log_pos), log_pos * sign)
    return bucket_pos.to(torch.long).numpy()


def rel_index_table(S: int, cfg: ArchConfig) -> np.ndarray:
    """idx[delta + S - 1] = clamp(bucket(delta) + span, 0, 2*span-1) for delta = i - j in (-S, S).

    c2p uses clamp(relpos + span) (T:318); p2c uses clamp(-relpos[j,i] + span) on the key axis and
    transposes (T:336-343); bucket() is odd so both reduce to this one table.
    """
    span = cfg.position_buckets
    delta = np.arange(-(S - 1), S)
    b = log_bucket(delta, cfg.position_buckets, cfg.max_relative_positions)
    return np.clip(b + span, 0, 2 * span - 1).astype(np.int64)


# --------------------------------------------------------------------------------------------
# the restated forward
# --------------------------------------------------------------------------------------------


def _ln(x, g, b, eps):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * g + b


def _gelu(x):
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


@torch.no_grad()
def forward_restated(w: dict, cfg: ArchConfig, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                     return_intermediates: bool = False):
    """fp32 forward: input_ids/attention_mask int64 [B,S] -> logits fp32 [B,C].

    Follows T:520-564 (embeddings), T:597-628 (rel-emb LN, mask, rel-pos), T:229-345 (attention),
    T:42-53 / T:384-446 (out/FFN/LN) and SURVEY.md App. B (head).
    """
    if cfg.backbone == "qwen2":
        return forward_restated_qwen2(w, cfg, input_ids, attention_mask, return_intermediates)
    B, S = input_ids.shape
    H, h, d = cfg.hidden_size, cfg.num_heads, cfg.head_dim
    eps = cfg.layer_norm_eps
    span = cfg.position_buckets
    inter = {}

    maskf = attention_mask.to(torch.float32)
    # 1. embeddings: LN(word_emb[ids]) * mask  (no position / token-type terms for v3)
    x = w[ENC + "embeddings.word_embeddings.weight"][input_ids]
    x = _ln(x, w[ENC + "embeddings.LayerNorm.weight"], w[ENC + "embeddings.LayerNorm.bias"], eps)
    x = x * maskf[..., None]
    inter["emb"] = x

    # 2. per-run constants
    rel = _ln(w[ENC + "encoder.rel_embeddings.weight"], w[ENC + "encoder.LayerNorm.weight"],
              w[ENC + "encoder.LayerNorm.bias"], eps)[: 2 * span]                      # [2span,H]
    tab = torch.from_numpy(rel_index_table(S, cfg))                                     # [2S-1]
    ii = torch.arange(S)
    idx = tab[(ii[:, None] - ii[None, :]) + (S - 1)]                                    # [S,S]
    # 3. attention mask: outer product (T:603-610)
    amask = (attention_mask[:, None, :, None] * attention_mask[:, None, None, :]).bool()  # [B,1,S,S]
    scale = math.sqrt(d * 3)
    fmin = torch.finfo(torch.float32).min

    def heads(t):  # [B,S,H] -> [B,h,S,d]
        return t.view(t.shape[0], t.shape[1], h, d).permute(0, 2, 1, 3)

    for l in range(cfg.num_layers):
        p = f"{ENC}encoder.layer.{l}."
        Wq, bq = w[p + "attention.self.query_proj.weight"], w[p + "attention.self.query_proj.bias"]
        Wk, bk = w[p + "attention.self.key_proj.weight"], w[p + "attention.self.key_proj.bias"]
        Wv, bv = w[p + "attention.self.value_proj.weight"], w[p + "attention.self.value_proj.bias"]
        q = heads(x @ Wq.T + bq)
        k = heads(x @ Wk.T + bk)
        v = heads(x @ Wv.T + bv)
        pos_q = heads((rel @ Wq.T + bq)[None])[0]                                        # [h,2span,d]
        pos_k = heads((rel @ Wk.T + bk)[None])[0]
        scores = q @ (k.transpose(-1, -2) / scale)                                       # [B,h,S,S]
        c2p_full = q @ pos_k.transpose(-1, -2)                                           # [B,h,S,2span]
        c2p = torch.gather(c2p_full, -1, idx[None, None].expand(B, h, S, S)) / scale
        p2c_full = k @ pos_q.transpose(-1, -2)                                           # [B,h,S(j),2span]
        # p2c[i,j] = p2c_full[j, idx(i-j)]
        p2c = torch.gather(p2c_full, -1, idx.T[None, None].expand(B, h, S, S)).transpose(-1, -2) / scale
        scores = scores + c2p + p2c
        scores = scores.masked_fill(~amask, fmin)
        probs = torch.softmax(scores, dim=-1)
        ctx = (probs @ v).permute(0, 2, 1, 3).reshape(B, S, H)
        if return_intermediates and l == 0:
            inter["qkv0"] = torch.cat([x @ Wq.T + bq, x @ Wk.T + bk, x @ Wv.T + bv], -1)
            inter["ctx0"] = ctx
        a = _ln(ctx @ w[p + "attention.output.dense.weight"].T + w[p + "attention.output.dense.bias"] + x,
                w[p + "attention.output.LayerNorm.weight"], w[p + "attention.output.LayerNorm.bias"], eps)
        f = _gelu(a @ w[p + "intermediate.dense.weight"].T + w[p + "intermediate.dense.bias"])
        pre2 = f @ w[p + "output.dense.weight"].T + w[p + "output.dense.bias"] + a
        if return_intermediates and l == 0:
            inter["preln2_0"] = pre2                  # the pre-LayerNorm sum of layer 0's FFN (fp16-range stress tests)
        x = _ln(pre2, w[p + "output.LayerNorm.weight"], w[p + "output.LayerNorm.bias"], eps)
        inter[f"h{l}"] = x

    logits = head_restated(w, cfg, x, input_ids, attention_mask)
    if return_intermediates:
        return logits, inter
    return logits


# --------------------------------------------------------------------------------------------
# decoder backbone: Qwen2 with bidirectional attention
# (Q: = transformers/models/qwen2/modeling_qwen2.py, 5.5.0: MLP Q:46-48, rotary Q:102-114, rotate_half /
#  apply_rotary_pos_emb Q:117-147, eager attention Q:160-185, attention Q:206-246, RMSNorm Q:258-264, decoder layer
#  Q:280-310, model Q:353-410.  The causal mask of Q:378-393 is replaced by a key-padding mask: GLiClass's decoder
#  backbones are used as bidirectional encoders, `M:` SURVEY.md §8 f4.)
# --------------------------------------------------------------------------------------------


def _rms(x, g, eps):
    return x * torch.rsqrt((x * x).mean(-1, keepdim=True) + eps) * g


@torch.no_grad()
def forward_restated_qwen2(w: dict, cfg: ArchConfig, input_ids: torch.Tensor, attention_mask: torch.Tensor,
                           return_intermediates: bool = False):
    B, S = input_ids.shape
    d, nh, nkv = cfg.head_dim, cfg.num_heads, cfg.num_kv_heads
    eps = cfg.rms_norm_eps
    inter = {}
    h = w[ENC + "embed_tokens.weight"][input_ids]                                     # [B,S,H]
    inter["emb"] = h
    inv_freq = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).float() / d))
    freqs = torch.arange(S).float()[:, None] * inv_freq[None, :]                      # [S,d/2]
    emb = torch.cat([freqs, freqs], -1)
    cos, sin = emb.cos()[None, None], emb.sin()[None, None]                           # [1,1,S,d]

    def rot(x):
        x1, x2 = x[..., : d // 2], x[..., d // 2:]
        return torch.cat([-x2, x1], -1)

    fmin = torch.finfo(torch.float32).min
    add = (1.0 - attention_mask.float())[:, None, None, :] * fmin                     # key padding only (bidirectional)
    for l in range(cfg.num_layers):
        p = f"{ENC}layers.{l}."
        x = _rms(h, w[p + "input_layernorm.weight"], eps)
        q = (x @ w[p + "self_attn.q_proj.weight"].T + w[p + "self_attn.q_proj.bias"]).view(B, S, nh, d).transpose(1, 2)
        k = (x @ w[p + "self_attn.k_proj.weight"].T + w[p + "self_attn.k_proj.bias"]).view(B, S, nkv, d).transpose(1, 2)
        v = (x @ w[p + "self_attn.v_proj.weight"].T + w[p + "self_attn.v_proj.bias"]).view(B, S, nkv, d).transpose(1, 2)
        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
        rep = nh // nkv
        kk = k[:, :, None].expand(B, nkv, rep, S, d).reshape(B, nh, S, d)
        vv = v[:, :, None].expand(B, nkv, rep, S, d).reshape(B, nh, S, d)
        a = torch.softmax(q @ kk.transpose(-1, -2) * (d ** -0.5) + add, dim=-1)
        ctx = (a @ vv).transpose(1, 2).reshape(B, S, nh * d)
        if return_intermediates and l == 0:
            inter["ctx0"] = ctx
        h = h + ctx @ w[p + "self_attn.o_proj.weight"].T
        x = _rms(h, w[p + "post_attention_layernorm.weight"], eps)
        g = x @ w[p + "mlp.gate_proj.weight"].T
        u = x @ w[p + "mlp.up_proj.weight"].T
        h = h + (F.silu(g) * u) @ w[p + "mlp.down_proj.weight"].T
        inter[f"h{l}"] = h
    hs = _rms(h, w[ENC + "norm.weight"], eps)
    inter["final"] = hs
    logits = head_restated(w, cfg, hs, input_ids, attention_mask)
    if return_intermediates:
        return logits, inter
    return logits


@torch.no_grad()
def head_restated(w: dict, cfg: ArchConfig, hseq: torch.Tensor, input_ids: torch.Tensor,
                  attention_mask: torch.Tensor | None = None) -> torch.Tensor:
    """GLiClass uni-encoder head (`gliclass` package, not installed -> unpinned).

    class rows = hidden state at each <<LABEL>> position (embed_class_token=True), zero rows for
    c >= count(b); text = pooled sequence (first / last token, masked mean, masked max); both through
    Linear-GELU-Linear; optional L2 normalisation (x / (|x| + 1e-8)) and logit_scale; scorer = dot
    (einsum BD,BCD->BC), MLP (cat[t,l] -> 256 -> 128 -> 1, ReLU) or weighted dot (proj to (Hh,2)
    halves, cat[t0, l0, t1*l1] -> 4Hh -> 1, ReLU).
    """
    B, S, H = hseq.shape
    m = input_ids == cfg.class_token_index
    n = m.sum(-1)
    C = int(n.max().item()) if B > 0 else 0
    cls = torch.zeros(B, C, H, dtype=hseq.dtype)
    for b in range(B):
        pos = torch.nonzero(m[b]).flatten()
        if not cfg.embed_class_token:        # gliclass: class_indices += 1 (the token after <<LABEL>>)
            pos = pos + 1
        cls[b, : len(pos)] = hseq[b, pos]
    if cfg.pooling_strategy == "first":
        pooled = hseq[:, 0, :]
    elif cfg.pooling_strategy == "last":
        pooled = hseq[:, S - 1, :]
    else:
        am = attention_mask.to(hseq.dtype)
        pooled = torch.zeros(B, H, dtype=hseq.dtype)
        for b in range(B):
            valid = hseq[b][am[b] != 0]
            pooled[b] = valid.sum(0) / am[b].sum() if cfg.pooling_strategy == "avg" else valid.max(0)[0]

    def lin(t, name):
        return t @ w[name + ".weight"].T + w[name + ".bias"]

    act = {"gelu": _gelu, "relu": torch.relu, "tanh": torch.tanh}[cfg.projector_hidden_act]

    def proj(t, name):
        return lin(act(lin(t, f"model.{name}.linear_1")), f"model.{name}.linear_2")

    t = proj(pooled, "text_projector")          # [B,Hh]
    kcls = proj(cls, "classes_projector")       # [B,C,Hh]   (zero rows still get the biases)
    if cfg.normalize_features:
        t = t / (torch.sqrt((t * t).sum(-1, keepdim=True)) + 1e-8)
        kcls = kcls / (torch.sqrt((kcls * kcls).sum(-1, keepdim=True)) + 1e-8)
    if cfg.scorer_type == "simple":
        logits = torch.einsum("bd,bcd->bc", t, kcls)
    elif cfg.scorer_type == "mlp":
        cat = torch.cat([t[:, None, :].expand(B, C, -1), kcls], -1)
        h1 = torch.relu(lin(cat, "model.scorer.mlp.0"))
        h2 = torch.relu(lin(h1, "model.scorer.mlp.2"))
        logits = lin(h2, "model.scorer.mlp.4")[..., 0]
    else:
        pt = lin(t, "model.scorer.proj_text")            # [B,2Hh], interleaved (d, half)
        pl = lin(kcls, "model.scorer.proj_label")        # [B,C,2Hh]
        t0, t1 = pt[:, None, 0::2].expand(B, C, -1), pt[:, None, 1::2].expand(B, C, -1)
        l0, l1 = pl[..., 0::2], pl[..., 1::2]
        cat = torch.cat([t0, l0, t1 * l1], -1)
        h1 = torch.relu(lin(cat, "model.scorer.out_mlp.0"))
        logits = lin(h1, "model.scorer.out_mlp.3")[..., 0]
    if cfg.normalize_features:
        logits = logits * w["model.logit_scale"]
    return logits


# --------------------------------------------------------------------------------------------
# decisions (reference src/postprocessor.c:14-16, 85-150)
# --------------------------------------------------------------------------------------------


def sigmoid32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-x, dtype=np.float32))).astype(np.float32)


def decisions_multilabel(logits: np.ndarray, threshold: float = 0.5) -> np.ndarray:
    """prob > threshold, strict (postprocessor.c:93-95)."""
    return sigmoid32(logits) > np.float32(threshold)


def decisions_singlelabel(logits: np.ndarray) -> np.ndarray:
    """argmax of sigmoid starting from max_prob = 0, max_idx = -1 (postprocessor.c:119-128)."""
    p = sigmoid32(logits)
    out = np.full(p.shape[0], -1, dtype=np.int64)
    for i in range(p.shape[0]):
        mp = np.float32(0.0)
        for j in range(p.shape[1]):
            if p[i, j] > mp:
                mp, out[i] = p[i, j], j
    return out


