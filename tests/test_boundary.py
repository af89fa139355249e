"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol the headers
declare, the ONNX reader recovers exactly the oracle's weights and config from a file exported
with the reference's own export call, the rel-pos table / decision helpers agree with the
oracle, and the ORT shim behaves like the API the unchanged reference sources call."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _declared(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    return sorted(set(re.findall(r"\b(glc_[a-z0-9_]+|Ort[A-Za-z_]+)\s*\(", src)) - {"glc_opts", "glc_info"})


def test_library_exports_every_declared_symbol(pkg):
    L = ctypes.CDLL(pkg.LIB_PATH)
    names = [n for n in _declared("gliclass_b200.h") if n.startswith("glc_")]
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/gliclass_b200.h but not exported"
    for n in ("OrtGetApiBase", "OrtSessionOptionsAppendExecutionProvider_CUDA"):
        assert hasattr(L, n)
    # and the python binding covers them all
    assert set(names) <= set(pkg._SIGS)


def test_onnx_reader_recovers_oracle_weights(pkg, orc, golden_onnx):
    cfg = orc.make_config("tiny")
    w = orc.init_weights(cfg, 0)
    f = pkg.OnnxFile(golden_onnx)
    i = f.info
    assert (i["vocab"], i["hidden"], i["layers"], i["heads"], i["inter"]) == (1027, 128, 2, 2, 512)
    assert i["head_hidden"] == 128 and i["buckets"] == 256 and i["max_rel_pos"] == 512
    assert i["class_token"] == cfg.class_token_index
    assert abs(i["ln_eps"] - 1e-7) < 1e-12
    E = orc.ENC
    expect = {"emb.word": E + "embeddings.word_embeddings.weight", "emb.ln.g": E + "embeddings.LayerNorm.weight",
              "emb.ln.b": E + "embeddings.LayerNorm.bias", "rel.emb": E + "encoder.rel_embeddings.weight",
              "rel.ln.g": E + "encoder.LayerNorm.weight", "rel.ln.b": E + "encoder.LayerNorm.bias"}
    for l in range(cfg.num_layers):
        p = f"{E}encoder.layer.{l}."
        for r, n in (("q", "attention.self.query_proj"), ("k", "attention.self.key_proj"), ("v", "attention.self.value_proj"),
                     ("o", "attention.output.dense"), ("ffn1", "intermediate.dense"), ("ffn2", "output.dense")):
            expect[f"layer.{l}.{r}.w"] = p + n + ".weight"       # [out,in], as torch stores it
            expect[f"layer.{l}.{r}.b"] = p + n + ".bias"
        for r, n in (("ln1", "attention.output.LayerNorm"), ("ln2", "output.LayerNorm")):
            expect[f"layer.{l}.{r}.g"] = p + n + ".weight"
            expect[f"layer.{l}.{r}.b"] = p + n + ".bias"
    for r, n in (("text", "text_projector"), ("cls", "classes_projector")):
        for k in (1, 2):
            expect[f"{r}.{k}.w"] = f"model.{n}.linear_{k}.weight"
            expect[f"{r}.{k}.b"] = f"model.{n}.linear_{k}.bias"
    assert sorted(f.roles()) == sorted(expect)
    for role, name in expect.items():
        got = f.tensor(role)
        ref = w[name].numpy()
        assert got.shape == ref.shape, role
        assert np.array_equal(got, ref), role      # bit exact: fp32 in, fp32 out
    f.close()


def test_onnx_reader_resolves_dedup_aliases(pkg, orc, tmp_path):
    # HF-default-style degenerate init: identical biases / gammas collapse into Identity aliases
    # in torch's exporter; the reader must still resolve every role.
    import torch
    cfg = orc.make_config("tiny")
    w = orc.init_weights(cfg, 3)
    for k in w:
        if k.endswith(".bias"):
            w[k] = torch.zeros_like(w[k])
        elif "LayerNorm.weight" in k:
            w[k] = torch.ones_like(w[k])
    path = str(tmp_path / "dedup.onnx")
    orc.export_onnx(orc.build_hf_module(cfg, w), cfg, path)
    f = pkg.OnnxFile(path)
    assert np.array_equal(f.tensor("layer.1.ln2.g"), np.ones(128, np.float32))
    assert np.array_equal(f.tensor("layer.0.k.b"), np.zeros(128, np.float32))
    assert np.array_equal(f.tensor("layer.1.ffn2.w"), w[orc.ENC + "encoder.layer.1.output.dense.weight"].numpy())
    f.close()


def test_onnx_reader_errors_are_loud(pkg, tmp_path):
    with pytest.raises(pkg.GlcError, match="cannot open"):
        pkg.OnnxFile(str(tmp_path / "nope.onnx"))
    bad = tmp_path / "bad.onnx"
    bad.write_bytes(b"\x0a\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\x01garbage")
    with pytest.raises(pkg.GlcError):
        pkg.OnnxFile(str(bad))
    trunc = tmp_path / "trunc.onnx"
    trunc.write_bytes(open(os.path.join(GOLDEN, "model.onnx"), "rb").read()[:100000])
    with pytest.raises(pkg.GlcError):
        pkg.OnnxFile(str(trunc))


@pytest.mark.parametrize("S", [1, 37, 128, 512, 700, 1024, 2048])
def test_rel_index_table_bit_exact(pkg, orc, S):
    cfg = orc.make_config("tiny")
    assert np.array_equal(pkg.rel_index_table(S, 256, 512), orc.rel_index_table(S, cfg).astype(np.int32))


def test_decide_matches_oracle(pkg, orc):
    rng = np.random.default_rng(0)
    lg = (rng.standard_normal((64, 10)) * 3).astype(np.float32)
    lg[0, :3] = [0.0, np.float32(1e-8), -np.float32(1e-8)]
    lg[1] = -np.inf
    m, a, p = pkg.decide(lg, 0.5)
    assert np.array_equal(m, orc.decisions_multilabel(lg, 0.5))
    assert np.array_equal(a, orc.decisions_singlelabel(lg))
    assert a[1] == -1
    assert np.allclose(p, orc.sigmoid32(lg), atol=1e-7)


def test_no_cpu_fallback(pkg, golden_onnx):
    """On a box without a B200 the product path must fail loudly, never compute on the CPU."""
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.GlcError, match="no usable sm_100|no CPU fallback"):
        pkg.Session(golden_onnx)


def test_product_path_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gliclass.c_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cc", ".h", ".cuh")):
                src = open(os.path.join(dirpath, fn), errors="ignore").read()
                assert "oracle" not in src.lower() or fn == "Makefile", f"{fn} mentions the oracle"


def test_ort_shim_host_behaviour(pkg, golden_onnx, tmp_path):
    exe = str(tmp_path / "shim_host_test")
    subprocess.run(["/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc", "-O1", "-Wall", "-Werror",
                    "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "shim_host_test.c"),
                    "-o", exe, "-L" + os.path.dirname(pkg.LIB_PATH), "-lgliclass_b200",
                    "-Wl,-rpath," + os.path.dirname(pkg.LIB_PATH)], check=True)
    r = subprocess.run([exe, golden_onnx], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK" in r.stdout and "missing-file:" in r.stdout


@pytest.mark.skipif(not os.path.exists("/root/reference/main.c"), reason="reference sources not present")
def test_unchanged_reference_sources_link_against_shim(pkg):
    """main.c, model.c, postprocessor.c, parallel_processor.c compile and link
    UNCHANGED against include/onnxruntime_c_api.h + libgliclass_b200.so."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    exe = os.path.join(ROOT, "oracle", "_ref", "gliclass_ref_main")
    assert os.path.exists(exe)
    syms = subprocess.run(["nm", "-u", exe], capture_output=True, text=True).stdout
    assert "OrtGetApiBase" in syms
    r = subprocess.run([exe], capture_output=True, text=True)
    assert "Usage:" in r.stdout


HEAD_VARIANTS = {
    "mlp_max": dict(scorer_type="mlp", pooling_strategy="max"),
    "wdot_avg_norm": dict(scorer_type="weighted-dot", pooling_strategy="avg", normalize_features=True),
    "dot_last_norm": dict(pooling_strategy="last", normalize_features=True),
}


@pytest.mark.parametrize("variant", sorted(HEAD_VARIANTS))
def test_onnx_reader_detects_head_variant(pkg, orc, model_cache, variant):
    """pooling strategy / scorer type / feature normalisation are read off the graph structure (the
    reference's onnx/config.json, convert_to_onnx.py:19-28, does not carry them) and the scorer weights
    come back bit-exact under their roles."""
    kw = HEAD_VARIANTS[variant]
    cfg, w = orc.make_model_file("tiny", os.path.join(model_cache, f"tiny_{variant}.onnx"), **kw)
    f = pkg.OnnxFile(os.path.join(model_cache, f"tiny_{variant}.onnx"))
    i = f.info
    assert i["pooling"] == {"first": 0, "last": 1, "avg": 2, "max": 3}[cfg.pooling_strategy]
    assert i["scorer"] == {"simple": 0, "mlp": 1, "weighted-dot": 2}[cfg.scorer_type]
    assert i["normalize_features"] == int(cfg.normalize_features)
    if cfg.normalize_features:
        assert i["logit_scale"] == np.float32(w["model.logit_scale"].item())
    roles = {"mlp": {"scorer.mlp.0": "model.scorer.mlp.0", "scorer.mlp.2": "model.scorer.mlp.2", "scorer.mlp.4": "model.scorer.mlp.4"},
             "weighted-dot": {"scorer.pt": "model.scorer.proj_text", "scorer.pl": "model.scorer.proj_label",
                              "scorer.o1": "model.scorer.out_mlp.0", "scorer.o2": "model.scorer.out_mlp.3"},
             "simple": {}}[cfg.scorer_type]
    for role, name in roles.items():
        assert np.array_equal(f.tensor(role + ".w"), w[name + ".weight"].numpy()), role
        assert np.array_equal(f.tensor(role + ".b"), w[name + ".bias"].numpy()), role
    f.close()


def test_onnx_reader_detects_projector_act_and_class_offset(pkg, orc, model_cache):
    """projector_hidden_act and embed_class_token are config-dependent in the gliclass package; the loader
    reads them off the graph (Erf vs Relu under /text_projector, Add(+1) on the NonZero-derived class positions)."""
    f = pkg.OnnxFile(os.path.join(GOLDEN, "model.onnx"))
    assert f.info["projector_act"] == 1 and f.info["class_pos_offset"] == 0
    f.close()
    path = os.path.join(model_cache, "tiny_relu_noembed.onnx")
    orc.make_model_file("tiny", path, projector_hidden_act="relu", embed_class_token=False)
    f = pkg.OnnxFile(path)
    assert f.info["projector_act"] == 2 and f.info["class_pos_offset"] == 1
    f.close()


def test_onnx_reader_refuses_unsupported_variants(pkg, orc, model_cache, tmp_path):
    """variants the engine does not implement must fail at load, not return wrong logits"""
    path = os.path.join(model_cache, "tiny_tanh.onnx")
    orc.make_model_file("tiny", path, projector_hidden_act="tanh")
    with pytest.raises(pkg.GlcError, match="unsupported projector activation"):
        pkg.OnnxFile(path)
    # an LSTM anywhere in the graph (gliclass use_lstm): append a minimal NodeProto{op_type="LSTM"} to the GraphProto
    raw = open(os.path.join(GOLDEN, "model.onnx"), "rb").read()

    def varint(n):
        out = b""
        while True:
            b7 = n & 0x7f
            n >>= 7
            out += bytes([b7 | (0x80 if n else 0)])
            if not n:
                return out

    # ModelProto: find the graph field (tag 0x3a) — rebuild the file with one more node appended inside it
    pos, fields = 0, []
    while pos < len(raw):
        tag = raw[pos]
        assert tag & 0x80 == 0
        wt, p2 = tag & 7, pos + 1
        if wt == 0:
            while raw[p2] & 0x80:
                p2 += 1
            p2 += 1
            fields.append((tag, raw[pos:p2], None))
        elif wt == 2:
            n, sh = 0, 0
            while True:
                b7 = raw[p2]
                n |= (b7 & 0x7f) << sh
                sh += 7
                p2 += 1
                if not b7 & 0x80:
                    break
            fields.append((tag, None, raw[p2:p2 + n]))
            p2 += n
        else:
            raise AssertionError(wt)
        pos = p2
    node = b"\x1a\x05/lstm" + b"\x22\x04LSTM"          # NodeProto{name=3, op_type=4}
    out = b""
    for tag, whole, payload in fields:
        if whole is not None:
            out += whole
        else:
            if tag == 0x3a:
                payload = payload + b"\x0a" + varint(len(node)) + node
            out += bytes([tag]) + varint(len(payload)) + payload
    bad = tmp_path / "lstm.onnx"
    bad.write_bytes(out)
    with pytest.raises(pkg.GlcError, match="use_lstm"):
        pkg.OnnxFile(str(bad))


def test_glc_load_validates_options(pkg, golden_onnx, monkeypatch):
    """a wrong struct_size, a garbage GLC_DEVICES or an out-of-range ordinal fail loudly"""
    L = pkg.lib()
    o = pkg.glc_opts()
    o.struct_size = 12
    assert not L.glc_load(os.fsencode(golden_onnx), ctypes.byref(o))
    assert "struct_size" in pkg.last_error()
    monkeypatch.setenv("GLC_DEVICES", "zero")
    assert not L.glc_load(os.fsencode(golden_onnx), None)
    assert "GLC_DEVICES" in pkg.last_error()


def test_onnx_reader_qwen2_backbone(pkg, orc, model_cache):
    """decoder backbone: roles, grouped-query geometry and rotary frequencies read off a Qwen2Model trace, bit exact"""
    path = os.path.join(model_cache, "qwen-mini.onnx")
    cfg, w = orc.make_model_file("qwen-mini", path, seed=0)
    f = pkg.OnnxFile(path)
    i = f.info
    assert (i["backbone"], i["layers"], i["hidden"], i["heads"], i["kv_heads"], i["head_dim"], i["inter"]) == (1, 3, 512, 4, 2, 128, 1536)
    assert i["class_token"] == cfg.class_token_index
    E = orc.ENC
    for role, name in (("emb.word", "embed_tokens.weight"), ("layer.1.q.w", "layers.1.self_attn.q_proj.weight"),
                       ("layer.1.k.b", "layers.1.self_attn.k_proj.bias"), ("layer.2.o.w", "layers.2.self_attn.o_proj.weight"),
                       ("layer.0.gate.w", "layers.0.mlp.gate_proj.weight"), ("layer.0.up.w", "layers.0.mlp.up_proj.weight"),
                       ("layer.2.down.w", "layers.2.mlp.down_proj.weight"), ("layer.1.ln1.g", "layers.1.input_layernorm.weight"),
                       ("layer.1.ln2.g", "layers.1.post_attention_layernorm.weight"), ("norm.g", "norm.weight")):
        assert np.array_equal(f.tensor(role), w[E + name].numpy()), role
    inv = 1.0 / (cfg.rope_theta ** (np.arange(0, 128, 2, dtype=np.float32) / 128))
    assert np.allclose(f.tensor("rope.inv_freq"), inv, rtol=1e-6)
    f.close()


def test_pack_plan_host_logic(pkg):
    """Host side of the varlen packing (f2; replaces the pad-to-longest batches of reference src/tokenizer.c:44-54): per text
    the kept length, its 128-aligned row count, the split into device launches, and the cases that keep the padded layout."""
    import ctypes as C
    L = pkg.lib()
    rng = np.random.default_rng(7)
    B, S, CLS = 40, 512, 900
    lens = rng.integers(40, S + 1, size=B)
    lens[3], lens[4], lens[5] = 128, 129, S
    ids = np.zeros((B, S), dtype=np.int64)
    mask = np.zeros((B, S), dtype=np.int64)
    for b, n in enumerate(lens):
        ids[b, :n] = rng.integers(3, 800, size=n)
        ids[b, n - 8], ids[b, n - 5] = CLS, CLS
        mask[b, :n] = 1
    mask[7, lens[7] // 2] = 0                      # an interior hole does not shorten the text
    kv = np.zeros(B, dtype=np.int32); rows = np.zeros(B, dtype=np.int32); launch = np.full(B, -1, dtype=np.int32)
    def plan(i, m, max_rows=65536, off=0):
        return L.glc_pack_plan(i.ctypes.data, m.ctypes.data, B, S, CLS, off, max_rows, kv.ctypes.data, rows.ctypes.data, launch.ctypes.data)
    total = plan(ids, mask)
    want_rows = np.maximum(128, (lens + 127) // 128 * 128)
    assert total == int(want_rows.sum()) and np.array_equal(kv, lens) and np.array_equal(rows, want_rows)
    assert (launch == 0).all()
    # launches of at most 4096 packed rows: contiguous, in order, never above the cap (a text larger than the cap goes alone)
    assert plan(ids, mask, max_rows=4096) == total
    assert launch[0] == 0 and (np.diff(launch) >= 0).all() and (np.diff(launch) <= 1).all() and launch[-1] >= 2
    for k in range(launch[-1] + 1):
        assert rows[launch == k].sum() <= 4096
    # a <<LABEL>> id in the padded tail: the padded layout would count it as a class, so the request is not packed
    ids2 = ids.copy(); ids2[11, S - 1] = CLS
    assert plan(ids2, mask) == 0
    # embed_class_token=false reads the token AFTER the class token: a class token on the last kept position disqualifies
    ids3 = ids.copy(); ids3[12, lens[12] - 1] = CLS
    assert plan(ids3, mask, off=0) == total and plan(ids3, mask, off=1) == 0
    # full-length batches save nothing: padded layout
    full = np.ones((B, S), dtype=np.int64)
    assert plan(ids, full) == 0
    assert L.glc_pack_plan(None, None, B, S, CLS, 0, 65536, None, None, None) == -1
