"""Attention-kernel micro benchmark on the bench shape (base arch: B=64, S=512, 12 heads) through
the C ABI (glc_op_attention), plus a parity check of the fused kernel against the slow CUDA-core
restatement on the same inputs.  Usage: python scripts/bench_attn.py [B S heads iters]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
L = pkg.lib()
B, S, heads, iters = (list(map(int, sys.argv[1:5])) + [64, 512, 12, 20][len(sys.argv) - 1:])[:4]
H = heads * 64
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(5)
qkv = torch.randn(B, S, 3 * H, generator=g)
qkv[..., :2 * H] *= 1.8
qkv = qkv.to(torch.float16).to(dev)
pos = (torch.randn(512, 2 * H, generator=g) * 1.8).to(torch.float16).to(dev)
mask = torch.ones(B, S, dtype=torch.long, device=dev)
Spad = (S + 127) // 128 * 128
rel = torch.from_numpy(pkg.rel_index_table(Spad, 256, 512)).to(dev)
bits = torch.zeros(B, (S + 31) // 32, dtype=torch.int32, device=dev)
kv = torch.zeros(B, dtype=torch.int32, device=dev)
assert L.glc_op_mask_prep(mask.data_ptr(), bits.data_ptr(), kv.data_ptr(), B, S, None) == 0
ctx = torch.zeros(B, S, H, dtype=torch.float16, device=dev)
pos_q, pos_k = pos[:, :H], pos[:, H:]


ER = L.glc_expanded_pos_rows()
exp = torch.zeros(ER, 2 * H, dtype=torch.float16, device=dev)
assert L.glc_op_expand_pos(pos.data_ptr(), 2 * H, 256, 512, exp.data_ptr(), 2 * H, 2 * H, None) == 0, pkg.last_error()
MODE = os.environ.get("GLC_ATTN", "toeplitz" if os.environ.get("GLC_ATTN_TOEPLITZ") == "1" else "gather")
LEGACY = MODE == "gather"
exps = torch.zeros(ER, 2 * H, dtype=torch.float16, device=dev)
assert L.glc_op_expand_pos_rev(pos.data_ptr(), 2 * H, 256, 512, exps.data_ptr(), 2 * H, H, None) == 0, pkg.last_error()
assert L.glc_op_expand_pos(pos[:, H:].data_ptr(), 2 * H, 256, 512, exps[:, H:].data_ptr(), 2 * H, H, None) == 0, pkg.last_error()
print("attention mode:", MODE)


def run(naive, out, nb=B):
    if not naive and MODE in ("shift", "stream"):
        op = L.glc_op_attention_shift if MODE == "shift" else L.glc_op_attention_stream
        rc = op(qkv.data_ptr(), exps[:, H:].data_ptr(), exps.data_ptr(), 2 * H, bits.data_ptr(),
                                      kv.data_ptr(), out.data_ptr(), nb, S, heads, None)
        assert rc == 0, pkg.last_error()
        return
    if not naive and not LEGACY:
        rc = L.glc_op_attention_toeplitz(qkv.data_ptr(), exp[:, H:].data_ptr(), exp.data_ptr(), 2 * H, bits.data_ptr(),
                                         kv.data_ptr(), out.data_ptr(), nb, S, heads, None)
        assert rc == 0, pkg.last_error()
        return
    rc = L.glc_op_attention(qkv.data_ptr(), pos_k.data_ptr(), pos_q.data_ptr(), 2 * H, rel.data_ptr(), bits.data_ptr(),
                            kv.data_ptr(), out.data_ptr(), nb, S, heads, 256, int(naive), None)
    assert rc == 0, pkg.last_error()


run(False, ctx)
torch.cuda.synchronize()
nb = min(B, 4)
ref = torch.zeros(nb, S, H, dtype=torch.float16, device=dev)
run(True, ref, nb)
torch.cuda.synchronize()
err = (ctx[:nb].float() - ref.float()).abs().max().item()
print(f"parity fused vs naive (first {nb} rows): max abs err {err:.3e}")
for _ in range(3):
    run(False, ctx)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(iters):
    run(False, ctx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 4.0 * B * S * S * H + 4.0 * B * S * 512 * H
print(f"attention B{B} S{S} h{heads}: {ms*1e3:.1f} us/launch, {fl/ms/1e9:.1f} algorithmic TFLOP/s")
assert err < 2e-2 or os.environ.get('GLC_ATTN_FLAGS')
