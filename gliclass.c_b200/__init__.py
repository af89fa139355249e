"""Host-side binding of libgliclass_b200.so (ctypes over the C ABI in include/gliclass_b200.h).

The directory name carries the reference's name ("GLiClass.c" -> gliclass.c_b200), which is not an
importable identifier; ``__graft_entry__.load_package()`` registers it as ``gliclass_c_b200``.

This module mirrors the reference's model.h interface for the one hot path
(reference include/model.h:8-19, src/model.c):

    initialize_ort_api / initialize_ort_environment / create_ort_session  ->  Session(model_path)
    flatten_int_array + create_tensor + prepare_input_tensors             ->  prepare_input_tensors()
    run_inference                                                         ->  Session.run_inference()

There is no CPU fallback anywhere in here: if the shared library is missing, or no sm_100 device
is usable, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgliclass_b200.so")

GLC_OK, GLC_ERR, GLC_ERR_ARG, GLC_ERR_CUDA, GLC_ERR_CAPACITY = 0, -1, -2, -3, -4


class GlcError(RuntimeError):
    pass


class glc_opts(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("num_devices", C.c_int32), ("device_ids", C.c_int32 * 8),
                ("weight_dtype", C.c_int32), ("max_tokens", C.c_int32), ("num_heads", C.c_int32),
                ("preln_f32", C.c_int32), ("reserved", C.c_int32 * 7)]


class glc_info(C.Structure):
    _fields_ = [("vocab", C.c_int32), ("hidden", C.c_int32), ("layers", C.c_int32), ("heads", C.c_int32),
                ("inter", C.c_int32), ("head_hidden", C.c_int32), ("buckets", C.c_int32), ("max_rel_pos", C.c_int32),
                ("ln_eps", C.c_float), ("class_token", C.c_int64), ("num_devices", C.c_int32),
                ("weight_dtype", C.c_int32), ("pooling", C.c_int32), ("scorer", C.c_int32),
                ("normalize_features", C.c_int32), ("logit_scale", C.c_float), ("projector_act", C.c_int32),
                ("class_pos_offset", C.c_int32), ("backbone", C.c_int32), ("kv_heads", C.c_int32), ("head_dim", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

_SIGS = {
    "glc_last_error": (C.c_char_p, []),
    "glc_device_count": (_i, []),
    "glc_load": (_vp, [C.c_char_p, C.POINTER(glc_opts)]),
    "glc_free": (None, [_vp]),
    "glc_model_info": (_i, [_vp, C.POINTER(glc_info)]),
    "glc_num_classes": (_i, [_vp, _vp, _i, _i]),
    "glc_run": (_i, [_vp, _vp, _vp, _i, _i, _vp, C.c_size_t, C.POINTER(_i)]),
    "glc_run_decisions": (_i, [_vp, _vp, _vp, _i, _i, _f, _vp, _vp, _vp, C.c_size_t, C.POINTER(_i)]),
    "glc_run_device": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _vp, _i]),
    "glc_sync": (_i, [_vp, _i]),
    "glc_stream": (_vp, [_vp, _i]),
    "glc_launch_count": (C.c_uint64, [_vp]),
    "glc_submit": (_vp, [_vp, _vp, _vp, _i, _i, _vp, C.c_size_t, C.POINTER(_i)]),
    "glc_poll": (_i, [_vp]),
    "glc_collect": (_i, [_vp]),
    "glc_coalesce_stats": (_i, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "glc_packed_stats": (_i, [_vp, _vp, _vp, _vp]),
    "glc_pack_plan": (_i64, [_vp, _vp, _i, _i, _i64, _i, _i, _vp, _vp, _vp]),
    "glc_profile_enable": (_i, [_vp, _i, _i]),
    "glc_profile_collect": (_i, [_vp, _i, _vp, _vp, _i]),
    "glc_debug_fetch": (_i64, [_vp, _i, C.c_char_p, _vp, C.c_size_t]),
    "glc_decide": (_i, [_vp, _i, _i, _f, _vp, _vp, _vp]),
    "glc_onnx_open": (_vp, [C.c_char_p]),
    "glc_onnx_close": (None, [_vp]),
    "glc_onnx_info": (_i, [_vp, C.POINTER(glc_info)]),
    "glc_onnx_tensor": (_i, [_vp, C.c_char_p, C.POINTER(C.POINTER(C.c_float)), C.POINTER(_i64)]),
    "glc_onnx_num_roles": (_i, [_vp]),
    "glc_onnx_role_name": (C.c_char_p, [_vp, _i]),
    "glc_rel_index_table": (_i, [_i, _i, _i, _vp]),
    "glc_op_gemm": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "glc_op_gemm_resid": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _vp, _i64, _i, _i, _i, _i, _i, _vp]),
    "glc_op_gemm_rope": (_i, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _vp, _i, _i, _vp]),
    "glc_op_gemm_e4m3": (_i, [_vp, _i64, _vp, _i64, _vp, _f, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _f, _vp]),
    "glc_op_quantize_rows_e4m3": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _vp]),
    "glc_op_residual_ln_e4m3": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _vp, _vp, _i, _i, _vp]),
    "glc_op_embed_ln": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _vp, _i, _i, _i, _vp]),
    "glc_op_residual_ln": (_i, [_vp, _vp, _vp, _vp, _f, _vp, _i, _i, _vp]),
    "glc_op_mask_prep": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "glc_op_attention_naive": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "glc_op_attention_persist": (_i, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "glc_expanded_pos_rows": (_i, []),
    "glc_op_expand_pos": (_i, [_vp, _i64, _i, _i, _vp, _i64, _i, _vp]),
    "glc_op_expand_pos_rev": (_i, [_vp, _i64, _i, _i, _vp, _i64, _i, _vp]),
    "glc_op_add_rmsnorm": (_i, [_vp, _vp, _vp, _f, _vp, _i, _i, _vp]),
    "glc_op_rope": (_i, [_vp, _i64, _vp, _i, _i, _i, _i, _vp]),
    "glc_op_attention_flash128": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "glc_op_head_gather": (_i, [_vp, _vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp]),
    "glc_op_head_score": (_i, [_vp, _vp, _vp, _vp, _vp, _f, _i, _i, _i, _vp]),
}


def lib() -> C.CDLL:
    """The loaded shared library.  Raises (never falls back) if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GlcError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no Python/CPU fallback for the hot path)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().glc_last_error() or b"").decode("utf-8", "replace")


def _check(rc: int, what: str) -> None:
    if rc != GLC_OK:
        raise GlcError(f"{what} failed ({rc}): {last_error()}")


def device_count() -> int:
    return int(lib().glc_device_count())


# ---------------------------------------------------------------------------------------------
# reference model.h mirror
# ---------------------------------------------------------------------------------------------


def flatten_int_array(rows: Sequence[Sequence[int]]) -> np.ndarray:
    """reference src/model.c:17-29: int[rows][cols] -> contiguous int64[rows*cols]."""
    a = np.ascontiguousarray(np.asarray(rows, dtype=np.int64))
    if a.ndim != 2:
        raise ValueError("expected a rectangular 2-D array of token ids")
    return a


def prepare_input_tensors(input_ids, attention_mask):
    """reference src/model.c:81-108: (TokenizedInputs) -> (input_ids, attention_mask) int64 [B,S]."""
    ids = flatten_int_array(input_ids)
    mask = flatten_int_array(attention_mask)
    if ids.shape != mask.shape:
        raise ValueError("input_ids and attention_mask must have the same [batch, seq] shape")
    return ids, mask


class Session:
    """reference create_ort_session (src/model.c:217-281) + run_inference (src/model.c:122-207)."""

    def __init__(self, model_path: str, devices: Optional[Sequence[int]] = None, max_tokens: int = 0,
                 weight_dtype: str = "default", preln_f32: bool = False, num_heads: int = 0):
        L = lib()
        o = glc_opts()
        o.struct_size = C.sizeof(glc_opts)
        if devices:
            o.num_devices = len(devices)
            for k, d in enumerate(devices):
                o.device_ids[k] = int(d)
        o.max_tokens = int(max_tokens)
        o.preln_f32 = 1 if preln_f32 else 0
        o.num_heads = int(num_heads)
        o.weight_dtype = {"default": 0, "fp16": 1, "bf16": 2, "fp8": 3}[weight_dtype]
        self._h = L.glc_load(os.fsencode(model_path), C.byref(o))
        if not self._h:
            raise GlcError(f"create session failed: {last_error()}")
        info = glc_info()
        _check(L.glc_model_info(self._h, C.byref(info)), "glc_model_info")
        self.info = info.as_dict()

    def close(self):
        if getattr(self, "_h", None):
            lib().glc_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- hot path, host buffers (what the reference's Run does: H2D, forward, D2H inside the call)
    def run_inference(self, input_ids: np.ndarray, attention_mask: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int64)
        if ids.ndim != 2 or ids.shape != mask.shape:
            raise ValueError("inputs must both be int64 [batch_size, sequence_length]")
        B, S = ids.shape
        L = lib()
        ncls = L.glc_num_classes(self._h, ids.ctypes.data, B, S)
        if ncls < 0:
            raise GlcError(last_error())
        if out is None:
            out = np.empty((B, ncls), dtype=np.float32)
        cc = C.c_int(0)
        _check(L.glc_run(self._h, ids.ctypes.data, mask.ctypes.data, B, S, out.ctypes.data, out.size, C.byref(cc)), "glc_run")
        return out.reshape(B, cc.value) if out.size == B * cc.value else out.ravel()[: B * cc.value].reshape(B, cc.value)

    def run_decisions(self, input_ids: np.ndarray, attention_mask: np.ndarray, threshold: float = 0.5):
        """run_inference + the reference's post-processing arithmetic (src/postprocessor.c:14-16,93-95)
        fused into the scorer kernel: returns (logits, probs, decisions[bool]) each [B, C]."""
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int64)
        B, S = ids.shape
        L = lib()
        ncls = max(0, L.glc_num_classes(self._h, ids.ctypes.data, B, S))
        lg = np.empty((B, ncls), dtype=np.float32)
        pr = np.empty((B, ncls), dtype=np.float32)
        de = np.zeros((B, ncls), dtype=np.uint8)
        cc = C.c_int(0)
        _check(L.glc_run_decisions(self._h, ids.ctypes.data, mask.ctypes.data, B, S, threshold, lg.ctypes.data, pr.ctypes.data,
                                   de.ctypes.data, lg.size, C.byref(cc)), "glc_run_decisions")
        return lg, pr, de.astype(bool)

    # -- asynchronous submit / collect: tokenise the next batch while this one runs
    def submit(self, input_ids: np.ndarray, attention_mask: np.ndarray):
        """Returns a ticket; collect(ticket) waits and returns logits [B, C]."""
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int64)
        B, S = ids.shape
        L = lib()
        ncls = max(0, L.glc_num_classes(self._h, ids.ctypes.data, B, S))
        out = np.empty((B, ncls), dtype=np.float32)
        cc = C.c_int(0)
        t = L.glc_submit(self._h, ids.ctypes.data, mask.ctypes.data, B, S, out.ctypes.data, out.size, C.byref(cc))
        if not t:
            raise GlcError(f"glc_submit failed: {last_error()}")
        return (t, ids, mask, out)   # keeps the buffers alive until collect

    def poll(self, ticket) -> bool:
        return lib().glc_poll(ticket[0]) == 1

    def collect(self, ticket) -> np.ndarray:
        _check(lib().glc_collect(ticket[0]), "glc_collect")
        return ticket[3]

    def submit_pinned(self, ids_ptr: int, mask_ptr: int, B: int, S: int, out_ptr: int, out_capacity: int):
        """glc_submit on raw host pointers (bench.py's pipelined e2e leg); the buffers must stay valid until collect_raw."""
        cc = C.c_int(0)
        t = lib().glc_submit(self._h, ids_ptr, mask_ptr, B, S, out_ptr, out_capacity, C.byref(cc))
        if not t:
            raise GlcError(f"glc_submit failed: {last_error()}")
        return t

    def collect_raw(self, ticket) -> None:
        _check(lib().glc_collect(ticket), "glc_collect")

    def coalesce_stats(self):
        """(merged launches, requests served by them) since load"""
        g, r = C.c_uint64(0), C.c_uint64(0)
        _check(lib().glc_coalesce_stats(self._h, C.byref(g), C.byref(r)), "glc_coalesce_stats")
        return int(g.value), int(r.value)

    def packed_stats(self):
        """(packed launches, rows computed, rows the padded [B,S] layout would have computed) since load"""
        a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        _check(lib().glc_packed_stats(self._h, C.byref(a), C.byref(b), C.byref(c)), "glc_packed_stats")
        return int(a.value), int(b.value), int(c.value)

    def run_pinned(self, ids_ptr: int, mask_ptr: int, B: int, S: int, out_ptr: int, out_capacity: int) -> int:
        """Same call on raw host pointers (pinned buffers in bench.py's e2e leg). Returns C."""
        cc = C.c_int(0)
        _check(lib().glc_run(self._h, ids_ptr, mask_ptr, B, S, out_ptr, out_capacity, C.byref(cc)), "glc_run")
        return cc.value

    # -- same forward on device-resident tensors (kernel-only timing, parity)
    def run_device(self, ids_ptr: int, mask_ptr: int, B: int, S: int, ncls: int, logits_ptr: int, slot: int = 0,
                   sync: bool = True) -> None:
        _check(lib().glc_run_device(self._h, slot, ids_ptr, mask_ptr, B, S, ncls, logits_ptr, 0 if sync else 1),
               "glc_run_device")

    def num_classes(self, input_ids: np.ndarray) -> int:
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        return int(lib().glc_num_classes(self._h, ids.ctypes.data, ids.shape[0], ids.shape[1]))

    def sync(self, slot: int = 0) -> None:
        _check(lib().glc_sync(self._h, slot), "glc_sync")

    def stream(self, slot: int = 0) -> int:
        return int(lib().glc_stream(self._h, slot) or 0)

    def launch_count(self) -> int:
        return int(lib().glc_launch_count(self._h))

    PROFILE_CATEGORIES = ("embed", "gemm_qkv", "attention", "gemm_out", "residual_ln", "gemm_ffn1", "gemm_ffn2",
                          "head_gemm", "head_misc")

    def profile_enable(self, on: bool, slot: int = 0) -> None:
        _check(lib().glc_profile_enable(self._h, slot, 1 if on else 0), "glc_profile_enable")

    def profile_collect(self, slot: int = 0) -> dict:
        """{category: (total_ms, launches)} accumulated since the last collect."""
        ms = np.zeros(16, dtype=np.float64)
        n = np.zeros(16, dtype=np.uint64)
        k = lib().glc_profile_collect(self._h, slot, ms.ctypes.data, n.ctypes.data, 16)
        if k < 0:
            raise GlcError(last_error())
        return {c: (float(ms[i]), int(n[i])) for i, c in enumerate(self.PROFILE_CATEGORIES[:k])}

    def debug_fetch(self, name: str, count: int, slot: int = 0) -> np.ndarray:
        out = np.empty(count, dtype=np.float32)
        n = lib().glc_debug_fetch(self._h, slot, name.encode(), out.ctypes.data, out.size)
        if n < 0:
            raise GlcError(f"debug_fetch({name}) -> {n}: {last_error()}")
        return out[:n]


SHIM_E2E_PATH = os.path.join(_HERE, "lib", "libglc_shim_e2e.so")


class ShimSession:
    """The reference's own calling sequence (src/model.c: flatten_int_array -> create_tensor -> run_inference, outputs
    read as src/postprocessor.c does) over the ORT-named entry points of the library, through tools/shim_e2e.c.  Every
    call mallocs pageable int64 copies of the inputs exactly like flatten_int_array (model.c:17-29)."""

    def __init__(self, model_path: str, num_threads: int = 8):
        lib()   # the engine library must be loaded first (the helper links against it)
        if not os.path.exists(SHIM_E2E_PATH):
            raise GlcError(f"{SHIM_E2E_PATH} not found: run the build")
        L = C.CDLL(SHIM_E2E_PATH)
        L.shim_e2e_open.restype = _vp
        L.shim_e2e_open.argtypes = [C.c_char_p, _i]
        L.shim_e2e_ok.restype = _i
        L.shim_e2e_ok.argtypes = [_vp]
        L.shim_e2e_error.restype = C.c_char_p
        L.shim_e2e_error.argtypes = [_vp]
        L.shim_e2e_close.restype = None
        L.shim_e2e_close.argtypes = [_vp]
        L.shim_e2e_run.restype = _i
        L.shim_e2e_run.argtypes = [_vp, _vp, _vp, _i, _i, _vp, C.c_size_t, C.POINTER(_i)]
        L.shim_e2e_time.restype = _i
        L.shim_e2e_time.argtypes = [_vp, _vp, _vp, _i, _i, _vp, C.c_size_t, _i, C.POINTER(C.c_double)]
        self._L = L
        self._h = L.shim_e2e_open(os.fsencode(model_path), int(num_threads))
        if not self._h or not L.shim_e2e_ok(self._h):
            msg = (L.shim_e2e_error(self._h) or b"").decode("utf-8", "replace")
            if self._h:
                L.shim_e2e_close(self._h)
                self._h = None
            raise GlcError(f"ORT shim CreateSession failed: {msg}")

    def run(self, input_ids: np.ndarray, attention_mask: np.ndarray, max_classes: int = 512) -> np.ndarray:
        ids = np.ascontiguousarray(input_ids, dtype=np.int64)
        mask = np.ascontiguousarray(attention_mask, dtype=np.int64)
        B, S = ids.shape
        out = np.empty(B * max_classes, dtype=np.float32)
        cc = C.c_int(0)
        if self._L.shim_e2e_run(self._h, ids.ctypes.data, mask.ctypes.data, B, S, out.ctypes.data, out.size, C.byref(cc)) != 0:
            raise GlcError("shim Run failed: " + (self._L.shim_e2e_error(self._h) or b"").decode("utf-8", "replace"))
        return out[: B * cc.value].reshape(B, cc.value).copy()

    def time_runs(self, ids: np.ndarray, mask: np.ndarray, out: np.ndarray, steps: int) -> float:
        """wall seconds of `steps` back-to-back reference-style run_inference calls on pageable buffers"""
        B, S = ids.shape
        sec = C.c_double(0.0)
        if self._L.shim_e2e_time(self._h, ids.ctypes.data, mask.ctypes.data, B, S, out.ctypes.data, out.size, int(steps),
                                 C.byref(sec)) != 0:
            raise GlcError("shim Run failed: " + (self._L.shim_e2e_error(self._h) or b"").decode("utf-8", "replace"))
        return float(sec.value)

    def close(self):
        if getattr(self, "_h", None):
            self._L.shim_e2e_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def decide(logits: np.ndarray, threshold: float = 0.5):
    """reference src/postprocessor.c:85-150 → (multi-label mask [B,C], single-label argmax [B], probs [B,C])."""
    lg = np.ascontiguousarray(logits, dtype=np.float32)
    B, ncls = lg.shape
    m = np.zeros((B, ncls), dtype=np.uint8)
    a = np.zeros((B,), dtype=np.int32)
    p = np.zeros((B, ncls), dtype=np.float32)
    _check(lib().glc_decide(lg.ctypes.data, B, ncls, threshold, m.ctypes.data, a.ctypes.data, p.ctypes.data), "glc_decide")
    return m.astype(bool), a, p


# ---------------------------------------------------------------------------------------------
# host-only model inspection
# ---------------------------------------------------------------------------------------------


class OnnxFile:
    def __init__(self, path: str):
        self._h = lib().glc_onnx_open(os.fsencode(path))
        if not self._h:
            raise GlcError(f"glc_onnx_open failed: {last_error()}")
        info = glc_info()
        _check(lib().glc_onnx_info(self._h, C.byref(info)), "glc_onnx_info")
        self.info = info.as_dict()

    def roles(self):
        L = lib()
        return [L.glc_onnx_role_name(self._h, i).decode() for i in range(L.glc_onnx_num_roles(self._h))]

    def tensor(self, role: str) -> np.ndarray:
        data = C.POINTER(C.c_float)()
        dims = (C.c_int64 * 4)()
        nd = lib().glc_onnx_tensor(self._h, role.encode(), C.byref(data), dims)
        if nd < 0:
            raise GlcError(last_error())
        shape = tuple(dims[i] for i in range(nd))
        n = int(np.prod(shape)) if shape else 1
        return np.ctypeslib.as_array(data, shape=(n,)).reshape(shape).copy()

    def close(self):
        if self._h:
            lib().glc_onnx_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rel_index_table(S: int, buckets: int = 256, max_pos: int = 512) -> np.ndarray:
    out = np.zeros(2 * S - 1, dtype=np.int32)
    _check(lib().glc_rel_index_table(S, buckets, max_pos, out.ctypes.data), "glc_rel_index_table")
    return out
