"""Per-kernel parity on a B200, through the C ABI (glc_op_*), against plain fp32 torch restatements
of the same op on the same fp16-rounded inputs.  Tolerances are fp16 output rounding (2^-11
relative) plus fp32-accumulation-order noise; a layout / descriptor bug shows up as O(1) error."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(pkg):
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    assert pkg.device_count() >= 1, "no usable sm_100 device"
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device("cuda:0")


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _sync_check(pkg, rc, what):
    assert rc == 0, f"{what}: rc={rc} {pkg.last_error()}"
    torch.cuda.synchronize()


def _report(name, got, ref, atol, rtol):
    d = (got.float() - ref.float()).abs()
    tol = atol + rtol * ref.float().abs()
    bad = d > tol
    msg = (f"{name}: max|d|={d.max().item():.4e} mean|d|={d.mean().item():.4e} ref_absmax={ref.abs().max().item():.3f} "
           f"bad={int(bad.sum())}/{bad.numel()}")
    print(msg)
    if bad.any():
        idx = torch.nonzero(bad)[:8].tolist()
        for ix in idx:
            print("   at", ix, "got", got[tuple(ix)].item(), "ref", ref[tuple(ix)].item())
        if got.dim() == 2:
            rows = torch.nonzero(bad.any(1)).flatten()
            cols = torch.nonzero(bad.any(0)).flatten()
            print(f"   bad rows: {rows.numel()} (first {rows[:10].tolist()})  bad cols: {cols.numel()} (first {cols[:10].tolist()})")
    assert not bad.any(), msg


# ---------------------------------------------------------------------------------------------
# K2 GEMM
# ---------------------------------------------------------------------------------------------

GEMM_CASES = [
    # M, N, K, act, out_f32
    (128, 128, 64, 0, False),      # one tile, one k-block
    (128, 256, 128, 0, False),
    (300, 384, 128, 0, False),     # ragged M, tiny arch QKV
    (77, 128, 512, 0, False),      # M < 128, tiny arch FFN2
    (1000, 512, 128, 1, False),    # GELU epilogue, tiny arch FFN1
    (512, 1536, 768, 0, False),    # pos projection shape (R x 2H)
    (4096, 2304, 768, 0, False),   # QKV, persistent loop with several tiles per CTA
    (4096, 768, 768, 0, False),
    (8192, 3072, 768, 1, False),   # FFN1 + GELU, 128x256 tiles
    (4096, 768, 3072, 0, False),   # FFN2, K = 3072
    (64, 768, 768, 0, True),       # head projector, fp32 out
    (640, 768, 768, 1, False),     # head projector 1 (GELU)
    (40000, 768, 768, 0, False),   # > 2 waves of 128x256 tiles -> wide path for N % 256 == 0
    (32768, 768, 768, 0, False),   # out-proj at the bench shape: 256x192 cluster tiles (7 rounds instead of 6 x 256)
    (32768, 768, 3072, 0, False),  # FFN2 at the bench shape, 256x192 cluster tiles
    (1000, 256, 256, 2, False),    # ReLU epilogue (MLP scorer layer 1)
    (1000, 128, 256, 2, True),     # ReLU epilogue, fp32 out (MLP scorer layer 2)
    (20000, 512, 384, 2, False),   # ReLU epilogue on the CTA-pair path (weighted-dot scorer at reranker scale)
]


@pytest.mark.parametrize("M,N,K,act,out_f32", GEMM_CASES)
def test_gemm(pkg, dev, M, N, K, act, out_f32):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    A = (torch.randn(M, K, generator=g) * 1.0).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    bias = (torch.randn(N, generator=g) * 0.5).float().to(dev)
    C = torch.full((M, N), float("nan"), dtype=torch.float32 if out_f32 else torch.float16, device=dev)
    rc = pkg.lib().glc_op_gemm(_ptr(A), K, _ptr(W), K, _ptr(bias), _ptr(C), N, M, N, K, act, int(out_f32), None)
    _sync_check(pkg, rc, "glc_op_gemm")
    ref = A.float() @ W.float().t() + bias
    if act == 1:
        ref = torch.nn.functional.gelu(ref)   # erf form
    if act == 2:
        ref = torch.relu(ref)
    if out_f32:
        _report(f"gemm {M}x{N}x{K} f32", C, ref, 2e-4, 2e-4)
    else:
        _report(f"gemm {M}x{N}x{K} act{act}", C, ref, 2e-3, 2e-3)


@pytest.mark.parametrize("M,N,K,act,out_f32", [
    (32768, 768, 768, 0, False),    # out-proj + x on the CTA-pair path (256x192 tiles, TMA-store epilogue)
    (4096, 768, 3072, 0, False),    # FFN2 + a at the batch-8 latency shape
    (300, 128, 128, 1, False),      # single-CTA kernel, ragged M, activation before the add
    (1000, 256, 128, 0, True),      # fp32 output
    (20000, 1024, 1024, 0, False),  # large arch
])
def test_gemm_fused_residual(pkg, dev, M, N, K, act, out_f32):
    """C = act(A W^T + b) + r: the residual add of the layer rides in the GEMM epilogue (T:49-53, T:408-412)"""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    bias = (torch.randn(N, generator=g) * 0.5).float().to(dev)
    Rbig = torch.randn(M, N + 8, generator=g).to(torch.float16).to(dev)   # strided residual view
    R = Rbig[:, :N]
    C = torch.full((M, N), float("nan"), dtype=torch.float32 if out_f32 else torch.float16, device=dev)
    rc = pkg.lib().glc_op_gemm_resid(_ptr(A), K, _ptr(W), K, _ptr(bias), R.data_ptr(), N + 8, _ptr(C), N, M, N, K, act,
                                     int(out_f32), None)
    _sync_check(pkg, rc, "glc_op_gemm_resid")
    ref = A.float() @ W.float().t() + bias
    if act == 1:
        ref = torch.nn.functional.gelu(ref)
    ref = ref + R.float()
    _report(f"gemm+resid {M}x{N}x{K}", C, ref, 3e-3, 3e-3)


# ---------------------------------------------------------------------------------------------
# K2 on e4m3 operands (opt-in FP8 FFN path): the kernel must reproduce, to fp32-accumulation accuracy, the product of the
# DEQUANTISED operands — quantisation error itself is measured end to end (test_gpu_e2e.py::test_fp8_ffn_mode)
# ---------------------------------------------------------------------------------------------


def _quantize_rows(pkg, dev, x16):
    M, K = x16.shape
    q = torch.empty(M, K, dtype=torch.uint8, device=dev)
    sc = torch.empty(M, dtype=torch.float32, device=dev)
    rc = pkg.lib().glc_op_quantize_rows_e4m3(_ptr(x16), K, _ptr(q), K, _ptr(sc), M, K, None)
    _sync_check(pkg, rc, "glc_op_quantize_rows_e4m3")
    return q, sc


def test_quantize_rows_e4m3(pkg, dev):
    g = torch.Generator().manual_seed(11)
    x = (torch.randn(300, 784, generator=g) * torch.rand(300, 1, generator=g) * 3).to(torch.float16).to(dev)
    x[7] = 0
    q, sc = _quantize_rows(pkg, dev, x)
    amax = x.float().abs().amax(1)
    want_sc = torch.where(amax > 0, amax / 448.0, torch.ones_like(amax))
    assert torch.allclose(sc, want_sc, rtol=1e-6, atol=0)
    deq = q.view(torch.float8_e4m3fn).float() * sc[:, None]
    # round-to-nearest e4m3: relative error <= 2^-4 of the value, or half a subnormal step of the row scale
    err = (deq - x.float()).abs()
    bound = torch.maximum(x.float().abs() * 2.0 ** -4, sc[:, None] * 2.0 ** -10) * 1.001
    assert (err <= bound).all(), float((err - bound).max())
    want_q = (x.float() / sc[:, None]).to(torch.float8_e4m3fn).view(torch.uint8)
    assert (q == want_q).float().mean() > 0.999   # ties / fp32 division vs multiplication by the reciprocal


@pytest.mark.parametrize("M,N,K", [(32768, 768, 3072), (4096, 768, 3072), (300, 192, 128), (1000, 256, 784), (77, 768, 3072)])
def test_gemm_e4m3_fp16_out(pkg, dev, M, N, K):
    """FFN2 shape: static activation scale, per-channel weight scale, fp16 output"""
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    bias = (torch.randn(N, generator=g) * 0.5).float().to(dev)
    mult = 16.0
    a8 = (A.float() * mult).to(torch.float8_e4m3fn)
    w8, ws = _quantize_rows(pkg, dev, W)
    C = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_gemm_e4m3(_ptr(a8), K, _ptr(w8), K, None, 1.0 / mult, _ptr(ws), _ptr(bias), _ptr(C), N, M, N, K, 0, 0,
                                    1.0, None)
    _sync_check(pkg, rc, "glc_op_gemm_e4m3")
    ref = (a8.float() @ w8.view(torch.float8_e4m3fn).float().t()) * (ws[None, :] / mult) + bias
    _report(f"gemm e4m3 {M}x{N}x{K}", C, ref, 2e-3, 2e-3)


@pytest.mark.parametrize("M,N,K", [(32768, 3072, 768), (4096, 3072, 768), (300, 256, 128), (999, 1024, 400)])
def test_gemm_e4m3_gelu_e4m3_out(pkg, dev, M, N, K):
    """FFN1 shape: per-row activation scales, per-channel weight scales, erf-GELU, e4m3 output under a static multiplier"""
    g = torch.Generator().manual_seed(M + N + K + 1)
    A = (torch.randn(M, K, generator=g) * (0.5 + torch.rand(M, 1, generator=g))).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    bias = (torch.randn(N, generator=g) * 0.5).float().to(dev)
    a8, a_s = _quantize_rows(pkg, dev, A)
    w8, ws = _quantize_rows(pkg, dev, W)
    mult = 4.0
    C = torch.full((M, N), 0x7f, dtype=torch.uint8, device=dev)
    rc = pkg.lib().glc_op_gemm_e4m3(_ptr(a8), K, _ptr(w8), K, _ptr(a_s), 1.0, _ptr(ws), _ptr(bias), _ptr(C), N, M, N, K, 1, 1,
                                    mult, None)
    _sync_check(pkg, rc, "glc_op_gemm_e4m3 (e4m3 out)")
    pre = (a8.view(torch.float8_e4m3fn).float() @ w8.view(torch.float8_e4m3fn).float().t()) * a_s[:, None] * ws[None, :] + bias
    ref = torch.nn.functional.gelu(pre) * mult
    got = C.view(torch.float8_e4m3fn).float()
    assert torch.isfinite(got).all()
    # one e4m3 step of slack: the result sits within 2^-3 relative (or one subnormal step 2^-9) of the fp32 reference.
    # The absolute term covers the packed-fp16 GELU of this epilogue: in the negative tail hx * (1 + tanh) cancels against
    # fp16's 2^-11, an absolute error of up to ~1.5e-3 (x mult) on outputs that are themselves ~1e-3
    err = (got - ref.clamp(-448, 448)).abs()
    bound = torch.maximum(ref.abs() * 2.0 ** -3, torch.full_like(ref, 2.0 ** -9)) + 1.5e-3 * mult
    assert (err <= bound).all(), f"max excess {(err - bound).max().item():.3e}"
    exact = (C == ref.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)).float().mean().item()
    print(f"gemm e4m3->e4m3 {M}x{N}x{K}: {100 * exact:.2f}% of output bytes identical to torch's rounding of the fp32 reference")
    assert exact > 0.95


@pytest.mark.parametrize("M,I,K", [(4096, 8960, 1536), (300, 128, 128), (1000, 1024, 512)])
def test_gemm_e4m3_swiglu_e4m3_out(pkg, dev, M, I, K):
    """decoder MLP of the FP8 tier: gate / up rows interleaved in blocks of 32, e4m3 operands, silu(gate) * up as e4m3"""
    g = torch.Generator().manual_seed(M + I + K)
    A = (torch.randn(M, K, generator=g) * (0.5 + torch.rand(M, 1, generator=g))).to(torch.float16).to(dev)
    Wg = (torch.randn(I, K, generator=g) / math.sqrt(K)).to(torch.float16)
    Wu = (torch.randn(I, K, generator=g) / math.sqrt(K)).to(torch.float16)
    W = torch.stack([Wg.view(I // 32, 32, K), Wu.view(I // 32, 32, K)], 1).reshape(2 * I, K).contiguous().to(dev)
    a8, a_s = _quantize_rows(pkg, dev, A)
    w8, ws = _quantize_rows(pkg, dev, W)
    mult = 4.0
    C = torch.full((M, I), 0x7f, dtype=torch.uint8, device=dev)
    rc = pkg.lib().glc_op_gemm_e4m3(_ptr(a8), K, _ptr(w8), K, _ptr(a_s), 1.0, _ptr(ws), None, _ptr(C), I, M, 2 * I, K, 3, 1, mult, None)
    _sync_check(pkg, rc, "glc_op_gemm_e4m3 (swiglu)")
    pre = (a8.view(torch.float8_e4m3fn).float() @ w8.view(torch.float8_e4m3fn).float().t()) * a_s[:, None] * ws[None, :]
    pre = pre.view(M, I // 32, 2, 32)
    ref = (torch.nn.functional.silu(pre[:, :, 0]) * pre[:, :, 1]).reshape(M, I) * mult
    got = C.view(torch.float8_e4m3fn).float()
    assert torch.isfinite(got).all()
    err = (got - ref.clamp(-448, 448)).abs()
    bound = torch.maximum(ref.abs() * 2.0 ** -3, torch.full_like(ref, 2.0 ** -9)) + 1.5e-3 * mult
    assert (err <= bound).all(), f"max excess {(err - bound).max().item():.3e}"
    exact = (C == ref.clamp(-448, 448).to(torch.float8_e4m3fn).view(torch.uint8)).float().mean().item()
    print(f"gemm e4m3 swiglu {M}x{2 * I}x{K}: {100 * exact:.2f}% of output bytes identical to torch's rounding of the fp32 reference")
    assert exact > 0.93


@pytest.mark.parametrize("H,M", [(768, 5001), (1024, 700), (128, 300)])
def test_residual_ln_e4m3_output(pkg, dev, H, M):
    g = torch.Generator().manual_seed(H + M + 5)
    x = (torch.randn(M, H, generator=g) * 1.7).to(torch.float16).to(dev)
    r = torch.randn(M, H, generator=g).to(torch.float16).to(dev)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).float().to(dev)
    beta = (0.02 * torch.randn(H, generator=g)).float().to(dev)
    y = torch.empty(M, H, dtype=torch.float16, device=dev)
    y8 = torch.empty(M, H, dtype=torch.uint8, device=dev)
    ys = torch.empty(M, dtype=torch.float32, device=dev)
    rc = pkg.lib().glc_op_residual_ln_e4m3(_ptr(x), _ptr(r), _ptr(gamma), _ptr(beta), 1e-7, _ptr(y), _ptr(y8), _ptr(ys), M, H, None)
    _sync_check(pkg, rc, "glc_op_residual_ln_e4m3")
    ref = torch.nn.functional.layer_norm(x.float() + r.float(), (H,), gamma, beta, 1e-7)
    _report(f"ln+e4m3 H={H} (fp16 output)", y, ref, 2e-3, 2e-3)
    assert torch.allclose(ys, ref.abs().amax(1) / 448.0, rtol=2e-3)
    deq = y8.view(torch.float8_e4m3fn).float() * ys[:, None]
    err = (deq - ref).abs()
    bound = torch.maximum(ref.abs() * 2.0 ** -4, ys[:, None] * 2.0 ** -10) * 1.01 + 2e-3
    assert (err <= bound).all(), float((err - bound).max())


@pytest.mark.parametrize("H,M", [(768, 70001), (1024, 40003), (128, 300)])
def test_ln_without_residual_operand(pkg, dev, H, M):
    g = torch.Generator().manual_seed(H + M)
    x = (torch.randn(M, H, generator=g) * 1.7).to(torch.float16).to(dev)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).float().to(dev)
    beta = (0.02 * torch.randn(H, generator=g)).float().to(dev)
    y = torch.empty(M, H, dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_residual_ln(_ptr(x), None, _ptr(gamma), _ptr(beta), 1e-7, _ptr(y), M, H, None)
    _sync_check(pkg, rc, "glc_op_residual_ln(no residual)")
    ref = torch.nn.functional.layer_norm(x.float(), (H,), gamma, beta, 1e-7)
    _report(f"ln H={H}", y, ref, 2e-3, 2e-3)


def test_gemm_strided_views(pkg, dev):
    # A and C as column slices of wider matrices (how qkv thirds / pos tables are addressed)
    M, N, K = 512, 256, 128
    g = torch.Generator().manual_seed(5)
    Abig = torch.randn(M, 3 * K, generator=g).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    Cbig = torch.zeros(M, 2 * N, dtype=torch.float16, device=dev)
    A = Abig[:, K:2 * K]
    Cv = Cbig[:, N:]
    rc = pkg.lib().glc_op_gemm(A.data_ptr(), 3 * K, _ptr(W), K, None, Cv.data_ptr(), 2 * N, M, N, K, 0, 0, None)
    _sync_check(pkg, rc, "glc_op_gemm strided")
    _report("gemm strided", Cbig[:, N:], A.float() @ W.float().t(), 2e-3, 2e-3)
    assert (Cbig[:, :N] == 0).all()


# ---------------------------------------------------------------------------------------------
# K1 / K4
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("H", [128, 256, 768, 1024])
def test_embed_ln(pkg, dev, H):
    V, M = 5003, 1000
    g = torch.Generator().manual_seed(H)
    emb = (torch.randn(V, H, generator=g) * 0.5).to(torch.float16).to(dev)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).float().to(dev)
    beta = (0.02 * torch.randn(H, generator=g)).float().to(dev)
    ids = torch.randint(0, V, (M,), generator=g).to(dev)
    mask = (torch.rand(M, generator=g) > 0.2).long().to(dev)
    y = torch.empty(M, H, dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_embed_ln(_ptr(ids), _ptr(mask), _ptr(emb), _ptr(gamma), _ptr(beta), 1e-7, _ptr(y), M, H, V, None)
    _sync_check(pkg, rc, "glc_op_embed_ln")
    ref = torch.nn.functional.layer_norm(emb[ids].float(), (H,), gamma, beta, 1e-7) * mask[:, None].float()
    _report(f"embed_ln H={H}", y, ref, 2e-3, 2e-3)


@pytest.mark.parametrize("H,M", [(128, 300), (768, 4096), (1024, 1000), (768, 70001), (1024, 40003), (1536, 20000), (2048, 3000)])
def test_residual_ln(pkg, dev, H, M):
    g = torch.Generator().manual_seed(H + M)
    x = torch.randn(M, H, generator=g).to(torch.float16).to(dev)
    r = torch.randn(M, H, generator=g).to(torch.float16).to(dev)
    gamma = (1 + 0.1 * torch.randn(H, generator=g)).float().to(dev)
    beta = (0.02 * torch.randn(H, generator=g)).float().to(dev)
    y = torch.empty(M, H, dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_residual_ln(_ptr(x), _ptr(r), _ptr(gamma), _ptr(beta), 1e-7, _ptr(y), M, H, None)
    _sync_check(pkg, rc, "glc_op_residual_ln")
    ref = torch.nn.functional.layer_norm(x.float() + r.float(), (H,), gamma, beta, 1e-7)
    _report(f"residual_ln H={H}", y, ref, 2e-3, 2e-3)


def test_mask_prep(pkg, dev):
    B, S = 7, 333
    g = torch.Generator().manual_seed(1)
    mask = torch.zeros(B, S, dtype=torch.long)
    lens = [333, 1, 0, 64, 65, 200, 32]
    for b, L in enumerate(lens):
        mask[b, :L] = 1
    mask[5, 17] = 0   # a hole
    words = (S + 31) // 32
    bits = torch.zeros(B, words, dtype=torch.int32, device=dev)
    kv = torch.zeros(B, dtype=torch.int32, device=dev)
    md = mask.to(dev)
    rc = pkg.lib().glc_op_mask_prep(_ptr(md), _ptr(bits), _ptr(kv), B, S, None)
    _sync_check(pkg, rc, "glc_op_mask_prep")
    assert kv.cpu().tolist() == lens
    got = bits.cpu().numpy().view(np.uint32)
    for b in range(B):
        for j in range(S):
            assert ((got[b, j // 32] >> (j % 32)) & 1) == mask[b, j].item()


# ---------------------------------------------------------------------------------------------
# K3 attention
# ---------------------------------------------------------------------------------------------


def _attention_ref(qkv, pos_k, pos_q, idx, mask, heads):
    """fp32 restatement on the bf16-rounded inputs.  qkv [B,S,3H]; pos_* [R,H]; idx [S,S] long."""
    B, S, H3 = qkv.shape
    H = H3 // 3
    d = H // heads
    q, k, v = [t.view(B, S, heads, d).permute(0, 2, 1, 3).float() for t in qkv.split(H, dim=-1)]
    pk = pos_k.view(-1, heads, d).permute(1, 0, 2).float()    # [h,R,d]
    pq = pos_q.view(-1, heads, d).permute(1, 0, 2).float()
    scale = math.sqrt(3 * d)
    s = q @ k.transpose(-1, -2)
    c2p = torch.gather(q @ pk.transpose(-1, -2), -1, idx[None, None].expand(B, heads, S, S))
    p2c = torch.gather(k @ pq.transpose(-1, -2), -1, idx.t()[None, None].expand(B, heads, S, S)).transpose(-1, -2)
    s = (s + c2p + p2c) / scale
    s = s.masked_fill(~mask[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, dim=-1)
    return (p @ v).permute(0, 2, 1, 3).reshape(B, S, H)


def _run_attention(pkg, dev, B, S, heads, lens, seed, kernel, qk_std=1.8):
    """kernel: 'naive' (CUDA-core restatement on the unexpanded tables), 'rows' (production) or 'shift'"""
    H = heads * 64
    R = 512
    g = torch.Generator().manual_seed(seed)
    qkv = torch.randn(B, S, 3 * H, generator=g)
    qkv[..., :2 * H] *= qk_std
    qkv = qkv.to(torch.float16).to(dev)
    pos = (torch.randn(R, 2 * H, generator=g) * qk_std).to(torch.float16).to(dev)   # [:, :H] = posQ, [:, H:] = posK
    mask = torch.zeros(B, S, dtype=torch.long)
    for b, L in enumerate(lens):
        mask[b, :L] = 1
    mask = mask.to(dev)
    Spad = (S + 127) // 128 * 128
    rel = torch.from_numpy(pkg.rel_index_table(Spad, 256, 512)).to(dev)
    words = (S + 31) // 32
    bits = torch.zeros(B, words, dtype=torch.int32, device=dev)
    kv = torch.zeros(B, dtype=torch.int32, device=dev)
    L = pkg.lib()
    _sync_check(pkg, L.glc_op_mask_prep(_ptr(mask), _ptr(bits), _ptr(kv), B, S, None), "mask_prep")
    ctx = torch.full((B, S, H), float("nan"), dtype=torch.float16, device=dev)
    pos_q, pos_k = pos[:, :H], pos[:, H:]
    if kernel == "persist":
        # posK half expanded in rho order, posQ half in the opposite (sigma) order, one row per relative distance
        ER = L.glc_expanded_pos_rows()
        exp = torch.full((ER, 2 * H), float("nan"), dtype=torch.float16, device=dev)
        _sync_check(pkg, L.glc_op_expand_pos_rev(_ptr(pos), 2 * H, 256, 512, _ptr(exp), 2 * H, H, None), "expand_pos_rev")
        _sync_check(pkg, L.glc_op_expand_pos(pos[:, H:].data_ptr(), 2 * H, 256, 512, exp[:, H:].data_ptr(), 2 * H, H, None),
                    "expand_pos")
        full = torch.from_numpy(pkg.rel_index_table(2048, 256, 512)).long().to(dev)    # idx[delta + 2047]
        assert torch.equal(exp[:ER - 1, H:], pos[full.flip(0)][:, H:]) and (exp[ER - 1] == 0).all()   # row rho = posK[idx(2047 - rho)]
        assert torch.equal(exp[:ER - 1, :H], pos[full][:, :H])                                        # row sigma = posQ[idx(sigma - 2047)]
        rc = L.glc_op_attention_persist(_ptr(qkv), exp[:, H:].data_ptr(), exp.data_ptr(), 2 * H, _ptr(bits), _ptr(kv), _ptr(ctx), B, S, heads, None)
        _sync_check(pkg, rc, "glc_op_attention_" + kernel)
    else:
        rc = L.glc_op_attention_naive(_ptr(qkv), pos_k.data_ptr(), pos_q.data_ptr(), 2 * H, _ptr(rel), _ptr(bits),
                                      _ptr(ctx), B, S, heads, 256, None)
        _sync_check(pkg, rc, "glc_op_attention_naive")
    ii = torch.arange(S, device=dev)
    idx = rel.long()[(ii[:, None] - ii[None, :]) + (Spad - 1)]
    ref = _attention_ref(qkv, pos_k.contiguous(), pos_q.contiguous(), idx, mask, heads)
    return ctx, ref, mask


ATT_CASES = [
    # B, S, heads, lens
    (1, 64, 1, [64]),              # one key tile, one (partial) query tile
    (1, 128, 2, [128]),            # two key tiles
    (1, 192, 1, [192]),            # three key tiles: every softmax group gets exactly one
    (2, 256, 2, [256, 256]),       # 2 q tiles x 4 k tiles: off-diagonal slices
    (2, 512, 2, [512, 300]),       # log-bucket region + ragged
    (3, 200, 2, [200, 37, 129]),   # S not a multiple of 64/128
    (1, 700, 1, [700]),            # beyond 512: clamp
    (2, 1024, 1, [1024, 555]),
    (4, 512, 12, [512, 512, 100, 1]),   # base-arch head count
]


@pytest.mark.parametrize("B,S,heads,lens", ATT_CASES)
def test_attention_naive_kernel(pkg, dev, B, S, heads, lens):
    """the slow CUDA-core restatement must agree with torch: it is the on-GPU debugging oracle"""
    if B * S * heads > 2 * 1024 * 4:
        pytest.skip("naive kernel only checked on small cases")
    ctx, ref, mask = _run_attention(pkg, dev, B, S, heads, lens, seed=S + B, kernel="naive")
    v = mask.bool()
    _report(f"attn-naive B{B} S{S} h{heads}", ctx[v], ref[v], 3e-3, 3e-3)


@pytest.mark.parametrize("kernel", ["persist"])
@pytest.mark.parametrize("B,S,heads,lens", ATT_CASES + [(1, 2048, 1, [2048]), (2, 1500, 2, [1500, 1]), (70, 384, 3, [384] * 35 + list(range(1, 36)))])
def test_attention(pkg, dev, B, S, heads, lens, kernel):
    """the attention kernel (csrc/attention_persist.cu) against the fp32 restatement"""
    ctx, ref, mask = _run_attention(pkg, dev, B, S, heads, lens, seed=S + B, kernel=kernel)
    v = mask.bool()
    got, want = ctx[v], ref[v]
    d = (got.float() - want.float()).abs()
    if d.max().item() > 1e-2 or torch.isnan(got.float()).any():
        full = (ctx.float() - ref.float()).abs().nan_to_num(99.0) * mask[..., None].float()
        for b in range(B):
            for h in range(heads):
                row = [f"{full[b, q0:q0 + 128, h * 64:(h + 1) * 64].max().item():.3f}" for q0 in range(0, S, 128)]
                print(f"   b{b} h{h} per-q-tile max err: {row}")
        bad = torch.nonzero(full.max(-1).values > 1e-2)
        print("   first bad (b,row):", bad[:10].tolist())
    _report(f"attn-{kernel} B{B} S{S} h{heads}", got, want, 1e-2, 1e-2)


@pytest.mark.parametrize("kernel", ["persist"])
def test_attention_softmax_peaked(pkg, dev, kernel):
    # large score magnitudes: the row maximum keeps growing across key tiles (exercises the sticky-maximum chain and the
    # rescale of the TMEM-resident output accumulator)
    ctx, ref, mask = _run_attention(pkg, dev, 1, 512, 2, [512], seed=99, kernel=kernel, qk_std=3.0)
    _report(f"attn-{kernel} peaked", ctx[mask.bool()], ref[mask.bool()], 2e-2, 2e-2)


def test_attention_growing_maximum(pkg, dev):
    """keys sorted so that every tile raises the row maximum by far more than 2^8: every softmax group rescales
    O in turn and the partial row sums of the other groups must follow the chained maximum"""
    B, S, heads = 1, 512, 1
    H = 64
    L = pkg.lib()
    g = torch.Generator().manual_seed(4)
    q = torch.randn(S, H, generator=g)
    kdir = torch.randn(H, generator=g)
    kdir = kdir / kdir.norm()
    # K_j = a_j * dir with a_j growing along the sequence; Q rows have a positive component along dir
    q = q * 0.3 + 2.0 * kdir
    k = torch.linspace(0.0, 60.0, S)[:, None] * kdir[None, :] + 0.1 * torch.randn(S, H, generator=g)
    v = torch.randn(S, H, generator=g)
    qkv = torch.cat([q, k, v], -1)[None].to(torch.float16).to(dev)
    pos = (torch.randn(512, 2 * H, generator=g) * 0.5).to(torch.float16).to(dev)
    mask = torch.ones(B, S, dtype=torch.long, device=dev)
    bits = torch.zeros(B, S // 32, dtype=torch.int32, device=dev)
    kv = torch.zeros(B, dtype=torch.int32, device=dev)
    _sync_check(pkg, L.glc_op_mask_prep(_ptr(mask), _ptr(bits), _ptr(kv), B, S, None), "mask_prep")
    ER = L.glc_expanded_pos_rows()
    exp = torch.zeros(ER, 2 * H, dtype=torch.float16, device=dev)
    _sync_check(pkg, L.glc_op_expand_pos_rev(_ptr(pos), 2 * H, 256, 512, _ptr(exp), 2 * H, H, None), "expand_pos_rev")
    _sync_check(pkg, L.glc_op_expand_pos(pos[:, H:].data_ptr(), 2 * H, 256, 512, exp[:, H:].data_ptr(), 2 * H, H, None), "expand_pos")
    ctx = torch.full((B, S, H), float("nan"), dtype=torch.float16, device=dev)
    rc = L.glc_op_attention_persist(_ptr(qkv), exp[:, H:].data_ptr(), exp.data_ptr(), 2 * H, _ptr(bits), _ptr(kv), _ptr(ctx), B, S, heads, None)
    _sync_check(pkg, rc, "glc_op_attention_persist")
    rel = torch.from_numpy(pkg.rel_index_table(S, 256, 512)).to(dev)
    ii = torch.arange(S, device=dev)
    idx = rel.long()[(ii[:, None] - ii[None, :]) + (S - 1)]
    ref = _attention_ref(qkv, pos[:, H:].contiguous(), pos[:, :H].contiguous(), idx, mask, heads)
    _report("attn-persist growing maximum", ctx[0], ref[0], 2e-2, 2e-2)


# ---------------------------------------------------------------------------------------------
# K5
# ---------------------------------------------------------------------------------------------


def test_head_gather_and_score(pkg, dev):
    B, S, H, C, tok = 5, 100, 128, 4, 1025
    g = torch.Generator().manual_seed(3)
    h = torch.randn(B, S, H, generator=g).to(torch.float16).to(dev)
    ids = torch.randint(3, 1000, (B, S), generator=g)
    counts = [4, 0, 2, 1, 3]
    for b, n in enumerate(counts):
        pos = torch.randperm(S - 1, generator=g)[:n].sort().values + 1
        ids[b, pos] = tok
    idsd = ids.to(dev)
    pooled = torch.empty(B, H, dtype=torch.float16, device=dev)
    cls = torch.full((B, C, H), float("nan"), dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_head_gather(_ptr(h), _ptr(idsd), tok, _ptr(pooled), _ptr(cls), B, S, H, C, None)
    _sync_check(pkg, rc, "glc_op_head_gather")
    assert torch.equal(pooled, h[:, 0])
    for b, n in enumerate(counts):
        pos = torch.nonzero(ids[b] == tok).flatten()
        for c in range(C):
            if c < n:
                assert torch.equal(cls[b, c], h[b, pos[c]])
            else:
                assert (cls[b, c] == 0).all()
    t = torch.randn(B, H, generator=g).to(dev)
    k = torch.randn(B, C, H, generator=g).to(dev)
    logits = torch.empty(B, C, device=dev)
    probs = torch.empty(B, C, device=dev)
    dec = torch.empty(B, C, dtype=torch.uint8, device=dev)
    rc = pkg.lib().glc_op_head_score(_ptr(t), _ptr(k), _ptr(logits), _ptr(probs), _ptr(dec), 0.5, B, C, H, None)
    _sync_check(pkg, rc, "glc_op_head_score")
    ref = torch.einsum("bd,bcd->bc", t, k)
    _report("head_score", logits, ref, 1e-4, 1e-5)
    assert torch.equal(dec.bool(), torch.sigmoid(logits) > 0.5)


# ---------------------------------------------------------------------------------------------
# decoder-backbone kernels (Qwen2-style stack)
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("H,M", [(512, 300), (1536, 1000), (2048, 77)])
def test_add_rmsnorm(pkg, dev, H, M):
    g = torch.Generator().manual_seed(H + M)
    h = (torch.randn(M, H, generator=g) * 3).to(dev)
    delta = torch.randn(M, H, generator=g).to(torch.float16).to(dev)
    w = (torch.randn(H, generator=g) * 0.1 + 1).to(dev)
    for d in (delta, None):
        hh = h.clone()
        y = torch.empty(M, H, dtype=torch.float16, device=dev)
        _sync_check(pkg, pkg.lib().glc_op_add_rmsnorm(_ptr(hh), _ptr(d), _ptr(w), 1e-6, _ptr(y), M, H, None), "add_rmsnorm")
        ref_h = h + (d.float() if d is not None else 0)
        ref_y = ref_h * torch.rsqrt((ref_h * ref_h).mean(-1, keepdim=True) + 1e-6) * w
        assert torch.equal(hh, ref_h) or (hh - ref_h).abs().max().item() < 1e-6
        _report(f"add_rmsnorm H{H} M{M} delta={d is not None}", y, ref_y, 2e-3, 2e-3)


def test_rope(pkg, dev):
    B, S, nh, nkv, d = 2, 333, 4, 2, 128
    W = (nh + 2 * nkv) * d
    g = torch.Generator().manual_seed(1)
    qkv = torch.randn(B * S, W, generator=g).to(torch.float16).to(dev)
    inv = (1.0 / (1.0e6 ** (torch.arange(0, d, 2).float() / d))).to(dev)
    out = qkv.clone()
    _sync_check(pkg, pkg.lib().glc_op_rope(_ptr(out), W, _ptr(inv), B * S, S, nh + nkv, d, None), "rope")
    pos = torch.arange(S, device=dev).float().repeat(B)
    ang = pos[:, None] * inv[None, :]
    cos, sin = torch.cat([ang.cos(), ang.cos()], -1), torch.cat([ang.sin(), ang.sin()], -1)
    x = qkv[:, : (nh + nkv) * d].float().view(B * S, nh + nkv, d)
    rot = torch.cat([-x[..., d // 2:], x[..., : d // 2]], -1)
    ref = (x * cos[:, None, :] + rot * sin[:, None, :]).reshape(B * S, -1)
    _report("rope q|k", out[:, : (nh + nkv) * d], ref, 2e-3, 2e-3)
    assert torch.equal(out[:, (nh + nkv) * d:], qkv[:, (nh + nkv) * d:])     # V untouched


@pytest.mark.parametrize("B,S,nh,nkv,K", [(32, 1024, 12, 2, 1536), (3, 333, 4, 2, 512), (1, 100, 3, 1, 256)])
def test_gemm_rope(pkg, dev, B, S, nh, nkv, K):
    """QKV projection with the rotary embedding fused into the epilogue (what the decoder stack runs): q and k heads rotated
    by the row position, v heads untouched, one rounding of the fp32 result"""
    d = 128
    N, M = (nh + 2 * nkv) * d, B * S
    g = torch.Generator().manual_seed(B + S + nh)
    A = torch.randn(M, K, generator=g).to(torch.float16).to(dev)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    bias = (torch.randn(N, generator=g) * 0.5).float().to(dev)
    inv = (1.0 / (1.0e6 ** (torch.arange(0, d, 2).float() / d))).to(dev)
    C = torch.full((M, N), float("nan"), dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_gemm_rope(_ptr(A), K, _ptr(W), K, _ptr(bias), _ptr(C), N, M, N, K, _ptr(inv), S, (nh + nkv) * d, None)
    _sync_check(pkg, rc, "glc_op_gemm_rope")
    y = A.float() @ W.float().t() + bias
    pos = torch.arange(S, device=dev).float().repeat(B)
    ang = pos[:, None] * inv[None, :]
    cos, sin = torch.cat([ang.cos(), ang.cos()], -1), torch.cat([ang.sin(), ang.sin()], -1)
    x = y[:, : (nh + nkv) * d].view(M, nh + nkv, d)
    rot = torch.cat([-x[..., d // 2:], x[..., : d // 2]], -1)
    ref = torch.cat([(x * cos[:, None, :] + rot * sin[:, None, :]).reshape(M, -1), y[:, (nh + nkv) * d:]], -1)
    _report(f"gemm+rope M{M} N{N} K{K}", C, ref, 3e-3, 3e-3)


@pytest.mark.parametrize("M,I,K", [(300, 1536, 512), (4096, 8960, 1536), (100, 64, 128)])
def test_gemm_swiglu(pkg, dev, M, I, K):
    """SwiGLU GEMM epilogue (act = 3): gate / up rows interleaved in blocks of 32, output [M, I] = silu(gate) * up"""
    g = torch.Generator().manual_seed(M + I)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.float16).to(dev)
    Wg = (torch.randn(I, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    Wu = (torch.randn(I, K, generator=g) / math.sqrt(K)).to(torch.float16).to(dev)
    Wi = torch.stack([Wg.view(I // 32, 32, K), Wu.view(I // 32, 32, K)], 1).reshape(2 * I, K).contiguous()
    C = torch.full((M, I), float("nan"), dtype=torch.float16, device=dev)
    rc = pkg.lib().glc_op_gemm(_ptr(A), K, _ptr(Wi), K, None, _ptr(C), I, M, 2 * I, K, 3, 0, None)
    _sync_check(pkg, rc, "glc_op_gemm(swiglu)")
    gate, up = A.float() @ Wg.float().t(), A.float() @ Wu.float().t()
    _report(f"gemm swiglu M{M} I{I} K{K}", C, torch.nn.functional.silu(gate) * up, 3e-3, 3e-3)


FLASH_CASES = [
    # B, S, heads, kv_heads, lens
    (1, 64, 2, 1, [64]),
    (1, 192, 2, 2, [192]),
    (2, 320, 4, 2, [320, 111]),
    (1, 1024, 6, 1, [1000]),
    (3, 700, 12, 2, [700, 1, 450]),
]


@pytest.mark.parametrize("B,S,heads,kvh,lens", FLASH_CASES)
def test_attention_flash128(pkg, dev, B, S, heads, kvh, lens):
    d = 128
    W = (heads + 2 * kvh) * d
    g = torch.Generator().manual_seed(S + heads)
    qkv = torch.randn(B, S, W, generator=g)
    qkv[..., : (heads + kvh) * d] *= 1.5
    qkv = qkv.to(torch.float16).to(dev)
    mask = torch.zeros(B, S, dtype=torch.long)
    for b, L in enumerate(lens):
        mask[b, :L] = 1
    mask = mask.to(dev)
    bits = torch.zeros(B, (S + 31) // 32, dtype=torch.int32, device=dev)
    kv = torch.zeros(B, dtype=torch.int32, device=dev)
    L_ = pkg.lib()
    _sync_check(pkg, L_.glc_op_mask_prep(_ptr(mask), _ptr(bits), _ptr(kv), B, S, None), "mask_prep")
    ctx = torch.full((B, S, heads * d), float("nan"), dtype=torch.float16, device=dev)
    _sync_check(pkg, L_.glc_op_attention_flash128(_ptr(qkv), _ptr(bits), _ptr(kv), _ptr(ctx), B, S, heads, kvh, None), "flash128")
    q = qkv[..., : heads * d].float().view(B, S, heads, d).permute(0, 2, 1, 3)
    k = qkv[..., heads * d: (heads + kvh) * d].float().view(B, S, kvh, d).permute(0, 2, 1, 3).repeat_interleave(heads // kvh, 1)
    v = qkv[..., (heads + kvh) * d:].float().view(B, S, kvh, d).permute(0, 2, 1, 3).repeat_interleave(heads // kvh, 1)
    s = q @ k.transpose(-1, -2) / math.sqrt(d)
    s = s.masked_fill(~mask[:, None, None, :].bool(), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B, S, heads * d)
    vm = mask.bool()
    _report(f"flash128 B{B} S{S} h{heads}/{kvh}", ctx[vm], ref[vm], 5e-3, 5e-3)
