"""Engine-vs-oracle logit error on sampled rows of a BASELINE config, for A/B runs of engine switches (env vars are read
at session load, so run one process per setting).  Usage: python scripts/precision_probe.py [arch S labels B rows]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

arch, S, NL, B, nrows = (sys.argv[1:6] + ["base", "512", "10", "64", "16"][len(sys.argv) - 1:])[:5]
S, NL, B, nrows = int(S), int(NL), int(B), int(nrows)
pkg, orc = graft.load_package(), graft.load_oracle()
d = os.environ.get("GLC_MODEL_CACHE", "/tmp/glc_models")
os.makedirs(d, exist_ok=True)
path = os.path.join(d, f"{arch}.onnx")
cfg, w = orc.make_model_file(arch, path, seed=0)
ids, mask = orc.synth_inputs(cfg, B, S, NL, seed=1235, ragged=True, min_frac=0.6)
rows = list(range(0, B, max(1, B // nrows)))[:nrows]
ref_file = os.path.join(d, f"probe_ref_{arch}_{S}_{NL}_{B}_{nrows}.npy")
if os.path.exists(ref_file):
    ref = np.load(ref_file)
else:
    ref = orc.forward_restated(w, cfg, ids[rows], mask[rows]).numpy()
    np.save(ref_file, ref)
s = pkg.Session(path, preln_f32=os.environ.get("PROBE_PRELN_F32") == "1")
out = s.run_inference(ids.numpy(), mask.numpy())[rows]
s.close()
e = np.abs(out - ref)
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith(("GLC_ATTN", "PROBE_", "GLC_PRELN")))
print(f"probe {arch} S{S} L{NL} rows {len(rows)} [{tag}]: max|d| {e.max():.4e} mean|d| {e.mean():.4e} p99 {np.quantile(e, 0.99):.4e}", flush=True)
