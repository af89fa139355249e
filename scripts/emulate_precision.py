"""CPU emulation of the engine's storage precisions (which tensors are rounded to bf16/fp16 and
where), to predict logit error vs the fp32 oracle before spending GPU time.  Not part of the
product or the tests."""
import math, sys, os, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import gliclass_oracle as O

def bf(x): return x.to(torch.bfloat16).float()
def hf(x): return x.to(torch.float16).float()
def ident(x): return x
def mx8(x, blk=32):
    # MXFP8-style: e4m3 elements, one power-of-two scale per `blk` consecutive elements of the last (contraction) dim
    sh = x.shape; K = sh[-1]; pad = (-K) % blk
    y = torch.nn.functional.pad(x, (0, pad)).reshape(-1, blk)
    amax = y.abs().amax(-1, keepdim=True).clamp_min(1e-30)
    sc = torch.exp2(torch.ceil(torch.log2(amax / 448.0)))
    q = (y / sc).to(torch.float8_e4m3fn).float() * sc
    return q.reshape(*sh[:-1], K + pad)[..., :K]

@torch.no_grad()
def forward_emul(w, cfg, ids, mask, act=bf, wt=bf, tmp_round=bf, resid_round=bf, bias_stage=hf, p_round=bf, head_round=bf, gin=ident):
    B, S = ids.shape; H, h, d = cfg.hidden_size, cfg.num_heads, cfg.head_dim
    eps = cfg.layer_norm_eps; span = cfg.position_buckets; E = O.ENC
    W = {k: (wt(v) if v.dim() == 2 else v) for k, v in w.items()}
    maskf = mask.float()
    x = O._ln(W[E+"embeddings.word_embeddings.weight"][ids], w[E+"embeddings.LayerNorm.weight"], w[E+"embeddings.LayerNorm.bias"], eps) * maskf[..., None]
    xr = resid_round(x); x = act(x)
    rel = act(O._ln(w[E+"encoder.rel_embeddings.weight"], w[E+"encoder.LayerNorm.weight"], w[E+"encoder.LayerNorm.bias"], eps))
    tab = torch.from_numpy(O.rel_index_table(S, cfg)); ii = torch.arange(S)
    idx = tab[(ii[:, None] - ii[None, :]) + (S - 1)]
    scale = math.sqrt(3 * d)
    heads = lambda t: t.view(t.shape[0], t.shape[1], h, d).permute(0, 2, 1, 3)
    for l in range(cfg.num_layers):
        p = f"{E}encoder.layer.{l}."
        Wq, bq = W[p+"attention.self.query_proj.weight"], w[p+"attention.self.query_proj.bias"]
        Wk, bk = W[p+"attention.self.key_proj.weight"], w[p+"attention.self.key_proj.bias"]
        Wv, bv = W[p+"attention.self.value_proj.weight"], w[p+"attention.self.value_proj.bias"]
        xg = gin(x); q = heads(act(xg @ Wq.T + bq)); k = heads(act(xg @ Wk.T + bk)); v = heads(act(xg @ Wv.T + bv))
        pq = heads(act(rel @ Wq.T + bq)[None])[0]; pk = heads(act(rel @ Wk.T + bk)[None])[0]
        s = q @ k.transpose(-1, -2)
        c2p = torch.gather(bias_stage(q @ pk.transpose(-1, -2)), -1, idx[None, None].expand(B, h, S, S))
        p2c = torch.gather(bias_stage(k @ pq.transpose(-1, -2)), -1, idx.T[None, None].expand(B, h, S, S)).transpose(-1, -2)
        s = (s + c2p + p2c) / scale
        s = s.masked_fill(~mask[:, None, None, :].bool(), float("-inf"))
        m = s.max(-1, keepdim=True).values
        pe = torch.exp(s - m); l_ = pe.sum(-1, keepdim=True)
        ctx = act(((p_round(pe) @ v) / l_).permute(0, 2, 1, 3).reshape(B, S, H))
        t = tmp_round(gin(ctx) @ W[p+"attention.output.dense.weight"].T + w[p+"attention.output.dense.bias"])
        a_full = O._ln(t + xr, w[p+"attention.output.LayerNorm.weight"], w[p+"attention.output.LayerNorm.bias"], eps)
        ar = resid_round(a_full); a = act(a_full)
        f = act(O._gelu(gin(a) @ W[p+"intermediate.dense.weight"].T + w[p+"intermediate.dense.bias"]))
        t = tmp_round(gin(f) @ W[p+"output.dense.weight"].T + w[p+"output.dense.bias"])
        x_full = O._ln(t + ar, w[p+"output.LayerNorm.weight"], w[p+"output.LayerNorm.bias"], eps)
        xr = resid_round(x_full); x = act(x_full)
    # head
    mcl = ids == cfg.class_token_index; n = mcl.sum(-1); C = int(n.max())
    cls = torch.zeros(B, C, H)
    for b in range(B):
        pos = torch.nonzero(mcl[b]).flatten(); cls[b, :len(pos)] = x[b, pos]
    def proj(t, name):
        t = head_round(O._gelu(t @ W[f"model.{name}.linear_1.weight"].T + w[f"model.{name}.linear_1.bias"]))
        return t @ W[f"model.{name}.linear_2.weight"].T + w[f"model.{name}.linear_2.bias"]
    return torch.einsum("bd,bcd->bc", proj(x[:, 0], "text_projector"), proj(cls, "classes_projector"))

if __name__ == "__main__":
    arch = sys.argv[1] if len(sys.argv) > 1 else "tiny"
    B, S, C = (int(a) for a in (sys.argv[2:5] if len(sys.argv) > 4 else (6, 200, 4)))
    cfg = O.make_config(arch); w = O.init_weights(cfg, 0)
    errs = {}
    variants = {
        "all-bf16 (as built)": dict(),
        "tmp fp32": dict(tmp_round=ident),
        "tmp+resid fp32": dict(tmp_round=ident, resid_round=ident),
        "weights fp32 only": dict(wt=ident),
        "act fp16": dict(act=hf, tmp_round=hf, resid_round=hf, p_round=hf, head_round=hf),
        "bias stage fp32": dict(bias_stage=ident),
        "fp16 + W8 (mx e4m3 weights)": dict(act=hf, tmp_round=hf, resid_round=hf, p_round=hf, head_round=hf, wt=lambda x: mx8(hf(x))),
        "fp16 + W8A8 (mx e4m3 both)": dict(act=hf, tmp_round=hf, resid_round=hf, p_round=hf, head_round=hf, wt=lambda x: mx8(hf(x)), gin=mx8),
        "fp16 + W8 FFN only A8": dict(act=hf, tmp_round=hf, resid_round=hf, p_round=hf, head_round=hf, wt=lambda x: mx8(hf(x)), gin=ident),
        "all fp32 (sanity)": dict(act=ident, wt=ident, tmp_round=ident, resid_round=ident, bias_stage=ident, p_round=ident, head_round=ident),
    }
    for seed in (1, 2, 3):
        ids, mask = O.synth_inputs(cfg, B, S, C, seed=seed, ragged=True)
        ref = O.forward_restated(w, cfg, ids, mask)
        for name, kw in variants.items():
            out = forward_emul(w, cfg, ids, mask, **kw)
            errs.setdefault(name, []).append((out - ref).abs())
    for name, e in errs.items():
        e = torch.cat([x.flatten() for x in e])
        print(f"{arch:6s} {name:24s} max {e.max():.4f}  p99 {e.quantile(0.99):.4f}  mean {e.mean():.4f}  (n={e.numel()})")
