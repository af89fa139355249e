"""Generates tests/golden/tiny/*: a random-init model.onnx exported with the reference's own
export call (ONNX_CONVERTING/convert_to_onnx.py:62-79), config.json in the schema of
convert_to_onnx.py:19-28 (incl. original_logits rounded to 5 dp), seeded inputs, fp32 oracle
logits from BOTH the traced HF module and the restated forward, and layer intermediates.

TEST INFRASTRUCTURE.  Run here (needs transformers); the outputs are committed.
    python oracle/make_golden.py
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gliclass_oracle as O  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = os.path.join(ROOT, "tests", "golden", "tiny")
    os.makedirs(out, exist_ok=True)
    cfg = O.make_config("tiny")
    w = O.init_weights(cfg, 0)
    m = O.build_hf_module(cfg, w)
    path = os.path.join(out, "model.onnx")
    if os.path.exists(path):
        os.remove(path)
    O.export_onnx(m, cfg, path)

    cases = {
        "full": O.synth_inputs(cfg, 4, 128, 4, seed=1235),                                  # all rows full length
        "ragged": O.synth_inputs(cfg, 5, 200, [4, 2, 3, 1, 4], seed=1236, ragged=True),      # pad-to-longest + mixed label counts
        "short": O.synth_inputs(cfg, 1, 37, 3, seed=1237),                                   # b < BATCH_SIZE, odd S
        "long": O.synth_inputs(cfg, 2, 700, 5, seed=1238, ragged=True, min_frac=0.8),        # S > 512: bucket clamp
    }
    blob = {}
    for name, (ids, mask) in cases.items():
        with torch.no_grad():
            hf = m(ids, mask).numpy()
        lg, inter = O.forward_restated(w, cfg, ids, mask, return_intermediates=True)
        lg = lg.numpy()
        assert np.abs(hf - lg).max() < 2e-5, (name, np.abs(hf - lg).max())
        blob[f"{name}.input_ids"] = ids.numpy()
        blob[f"{name}.attention_mask"] = mask.numpy()
        blob[f"{name}.logits_hf"] = hf
        blob[f"{name}.logits"] = lg
        if name == "ragged":
            blob["ragged.emb"] = inter["emb"].numpy().astype(np.float16)
            blob["ragged.qkv0"] = inter["qkv0"].numpy().astype(np.float16)
            blob["ragged.ctx0"] = inter["ctx0"].numpy().astype(np.float16)
            blob["ragged.h1"] = inter["h1"].numpy().astype(np.float16)
    np.savez_compressed(os.path.join(out, "cases.npz"), **blob)

    # the reference's config.json schema with its fixed 1 text x 4 labels style vector
    ids, mask = cases["short"]
    with torch.no_grad():
        ol = m(ids, mask).round(decimals=5).tolist()
    with open(os.path.join(out, "config.json"), "w") as f:
        json.dump({"original_model_name": "random-init/gliclass-tiny-arch", "architecture_type": "uni-encoder",
                   "prompt_first": False, "original_logits": ol, "oracle_config": O.config_dict(cfg)}, f, indent=4)
    # relative-position index tables straight from HF's build_relative_position
    from transformers.models.deberta_v2.modeling_deberta_v2 import build_relative_position
    tabs = {}
    for S in (37, 128, 512, 700, 1024, 2048):
        q = torch.zeros(1, S, 1)
        rp = build_relative_position(q, q, bucket_size=cfg.position_buckets, max_position=cfg.max_relative_positions)[0]
        idx = torch.clamp(rp + cfg.position_buckets, 0, 2 * cfg.position_buckets - 1)
        # idx[i, j] depends on i - j only: store first column reversed + first row
        col = idx[:, 0].numpy()            # delta = i      (0..S-1)
        row = idx[0, :].numpy()            # delta = -j     (0..-(S-1))
        assert all((idx.diagonal(d) == idx.diagonal(d)[0]).all() for d in range(-S + 1, S, max(1, S // 50)))
        tabs[f"S{S}"] = np.concatenate([row[::-1][:-1], col]).astype(np.int32)   # delta = -(S-1)..S-1
    np.savez_compressed(os.path.join(out, "rel_tables.npz"), **tabs)
    print("wrote", out, {k: os.path.getsize(os.path.join(out, k)) for k in os.listdir(out)})


if __name__ == "__main__":
    main()
