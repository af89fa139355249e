"""GEMM-kernel micro benchmark on the bench shapes (base arch, M = 64*512 tokens) through the C ABI
(glc_op_gemm), each checked against torch.matmul on the same fp16 operands."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

pkg = graft.load_package()
L = pkg.lib()
dev = torch.device("cuda", 0)
M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
iters = 20
shapes = [("qkv", 2304, 768, 0), ("out", 768, 768, 0), ("ffn1", 3072, 768, 1), ("ffn2", 768, 3072, 0)]
g = torch.Generator().manual_seed(1)
tot_ms, tot_fl = 0.0, 0.0
for name, N, K, act in shapes:
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) * 0.05).to(torch.float16).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    C = torch.empty(M, N, dtype=torch.float16, device=dev)

    def run():
        rc = L.glc_op_gemm(A.data_ptr(), K, W.data_ptr(), K, bias.data_ptr(), C.data_ptr(), N, M, N, K, act, 0, None)
        assert rc == 0, pkg.last_error()

    run()
    torch.cuda.synchronize()
    ref = A[:4096].float() @ W.float().t() + bias
    if act:
        ref = torch.nn.functional.gelu(ref)
    err = (C[:4096].float() - ref).abs().max().item()
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
    fl = 2.0 * M * N * K
    tot_ms += ms
    tot_fl += fl
    print(f"{name:5s} M{M} N{N} K{K} act{act}: {ms*1e3:7.1f} us  {fl/ms/1e9:7.1f} TFLOP/s  max|err| {err:.3e}")
    assert err < 3e-2
print(f"layer total {tot_ms*1e3:.1f} us, {tot_fl/tot_ms/1e9:.1f} TFLOP/s ; x12 layers = {tot_ms*12:.3f} ms")
