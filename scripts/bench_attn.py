"""Attention-kernel micro benchmark on the bench shape (base arch: B=64, S=512, 12 heads) through the C ABI
(glc_op_attention_persist), plus a parity check against the slow CUDA-core restatement on the
same inputs.  Usage: python scripts/bench_attn.py [B S heads iters]"""
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
pos_q, pos_k = pos[:, :H], pos[:, H:]
ER = L.glc_expanded_pos_rows()
exps = torch.zeros(ER, 2 * H, dtype=torch.float16, device=dev)
assert L.glc_op_expand_pos_rev(pos.data_ptr(), 2 * H, 256, 512, exps.data_ptr(), 2 * H, H, None) == 0, pkg.last_error()
assert L.glc_op_expand_pos(pos[:, H:].data_ptr(), 2 * H, 256, 512, exps[:, H:].data_ptr(), 2 * H, H, None) == 0, pkg.last_error()
nb = min(B, 4)
ref = torch.zeros(nb, S, H, dtype=torch.float16, device=dev)
rc = L.glc_op_attention_naive(qkv.data_ptr(), pos_k.data_ptr(), pos_q.data_ptr(), 2 * H, rel.data_ptr(), bits.data_ptr(),
                              ref.data_ptr(), nb, S, heads, 256, None)
assert rc == 0, pkg.last_error()
torch.cuda.synchronize()
fl = 4.0 * B * S * S * H + 4.0 * B * S * 512 * H
for name in ("persist",):
    op = L.glc_op_attention_persist
    ctx = torch.zeros(B, S, H, dtype=torch.float16, device=dev)

    def run():
        rc = op(qkv.data_ptr(), exps[:, H:].data_ptr(), exps.data_ptr(), 2 * H, bits.data_ptr(), kv.data_ptr(),
                ctx.data_ptr(), B, S, heads, None)
        assert rc == 0, pkg.last_error()

    run()
    torch.cuda.synchronize()
    err = (ctx[:nb].float() - ref.float()).abs().max().item()
    for _ in range(3):
        run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print(f"attention_{name} B{B} S{S} h{heads}: {ms*1e3:.1f} us/launch, {fl/ms/1e9:.1f} algorithmic TFLOP/s, "
          f"max abs err vs naive (first {nb} rows) {err:.3e}", flush=True)
    assert err < 2e-2, name
