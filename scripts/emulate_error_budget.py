"""Error budget of fp16 storage on the discriminating fixtures: the fp32 oracle with the engine's rounding points switched on
one group at a time (scripts/emulate_precision.py), CPU only.  Usage: python scripts/emulate_error_budget.py base 6 384 10"""
import sys, os, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import emulate_precision as E
import gliclass_oracle as O
arch, B, S, C = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = O.make_config(arch); w = O.init_weights(cfg, 0)
ids, mask = O.synth_inputs(cfg, B, S, C, seed=1235, ragged=True, min_frac=0.6)
torch.set_num_threads(8)
ref = O.forward_restated(w, cfg, ids, mask)
hf, ident = E.hf, E.ident
base = dict(act=hf, wt=hf, tmp_round=hf, resid_round=hf, p_round=hf, head_round=hf, bias_stage=hf)
variants = {
  "engine-like fp16 everywhere": base,
  "residual path fp32": dict(base, resid_round=ident),
  "tmp + residual fp32": dict(base, resid_round=ident, tmp_round=ident),
  "weights fp32": dict(base, wt=ident),
  "only weights fp16": dict(act=ident, wt=hf, tmp_round=ident, resid_round=ident, p_round=ident, head_round=ident, bias_stage=ident),
  "only bias stage fp16": dict(act=ident, wt=ident, tmp_round=ident, resid_round=ident, p_round=ident, head_round=ident, bias_stage=hf),
  "bias stage fp32 (rest fp16)": dict(base, bias_stage=ident),
}
for name, kw in variants.items():
    t0 = time.time()
    out = E.forward_emul(w, cfg, ids, mask, **kw)
    e = (out - ref).abs()
    print(f"{arch} B{B} S{S}: {name:32s} max {e.max():.3e} mean {e.mean():.3e}  ({time.time()-t0:.0f}s)", flush=True)
