"""tcgen05 GEMM / implicit-GEMM conv kernel vs a plain fp32 torch reference of the same op
(floating-point kernel: bf16 operands, fp32 accumulation).  Tolerance: the inputs are
bf16-exact, so the only error sources are accumulation order and the bf16 rounding of the
output: |err| <= 2^-8 * |ref| + 1e-3 * sqrt(K)-scaled slack."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _lin(native_lib, M, K, N, act=0, bias=True, res=None, out_f32=False, BN=0, seed=0, resident=0):
    from tuatara_b200._native import check

    g = torch.Generator(device="cpu").manual_seed(seed)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    b = (torch.randn(N, generator=g)).float().cuda() if bias else None
    R = None
    if res == "bf16":
        R = torch.randn(M, N, generator=g).to(torch.bfloat16).cuda()
    elif res == "f32":
        R = torch.randn(M, N, generator=g).float().cuda()
    out = torch.full((M, N), float("nan"), dtype=torch.float32 if out_f32 else torch.bfloat16, device="cuda")
    check(native_lib.tt_linear_dev(A.data_ptr(), K, M, K, W.data_ptr(), N, b.data_ptr() if bias else None, act,
                                   R.data_ptr() if R is not None else None, int(res == "f32"), N,
                                   out.data_ptr(), int(out_f32), N, BN, resident, None), "tt_linear_dev")
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t()
    if bias:
        ref = ref + b
    if R is not None:
        ref = ref + R.float()
    if act == 1:
        ref = torch.relu(ref)
    elif act == 2:
        ref = torch.nn.functional.gelu(ref)
    return out.float().cpu(), ref.cpu()


def _check(out, ref, K, what):
    assert torch.isfinite(out).all(), f"{what}: non-finite output (unwritten tiles?)"
    err = (out - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 2e-3 * np.sqrt(K)
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()
        rows = idx[:, 0].unique()[:8].tolist()
        cols = idx[:, 1].unique()[:16].tolist()
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements off, max err {float(err.max()):.4g}; "
                             f"first bad rows {rows} cols {cols}; out[0,:8]={out[0,:8].tolist()} ref[0,:8]={ref[0,:8].tolist()}")


@pytest.mark.parametrize("M,K,N,BN", [
    (128, 64, 64, 64), (128, 128, 128, 128), (256, 64, 256, 256), (1000, 384, 1152, 0), (4096, 1536, 384, 0),
    (333, 96, 384, 0), (77, 384, 96, 96), (512, 384, 1536, 256), (128, 64, 16, 16), (300, 32, 32, 32),
    (20000, 384, 384, 192),
])
def test_linear_shapes(native_lib, M, K, N, BN):
    out, ref = _lin(native_lib, M, K, N, BN=BN)
    _check(out, ref, K, f"linear M{M} K{K} N{N} BN{BN}")


@pytest.mark.parametrize("M,K,N,BN", [(5000, 384, 1152, 192), (131072, 384, 384, 192), (3000, 96, 384, 192),
                                      (70000, 384, 1536, 192), (2000, 64, 64, 64), (1200, 384, 768, 128)])
def test_linear_weight_resident(native_lib, M, K, N, BN):
    """The weight-resident schedule (B slice loaded once per CTA, only A streams)."""
    out, ref = _lin(native_lib, M, K, N, BN=BN, act=2 if N == 1536 else 0, res="f32" if N == 384 else None,
                    out_f32=(N == 384), resident=1)
    _check(out, ref, K, f"resident linear M{M} K{K} N{N} BN{BN}")


@pytest.mark.parametrize("M,K,N,BN,sched", [
    (256, 64, 256, 256, 2), (128, 64, 64, 64, 2), (100, 384, 1152, 192, 2), (1000, 384, 1152, 192, 2), (4096, 1536, 384, 192, 2),
    (333, 96, 384, 128, 2), (20000, 384, 1536, 256, 2), (640, 384, 96, 96, 2), (300, 32, 32, 32, 2), (385, 64, 16, 16, 2),
    (20000, 384, 1536, 256, 3), (131072 + 128, 384, 1152, 192, 3), (70000, 384, 384, 128, 3), (3000, 96, 384, 192, 3),
    (2400, 384, 768, 256, 3),
])
def test_linear_pair(native_lib, M, K, N, BN, sched):
    """CTA-pair schedule (cta_group::2): 256-row tiles, odd tile counts, streaming (2) and weight-resident (3)."""
    out, ref = _lin(native_lib, M, K, N, BN=BN, act=2 if N == 1536 else 0, res="f32" if N == 384 else None,
                    out_f32=(N == 384), resident=sched)
    _check(out, ref, K, f"pair linear M{M} K{K} N{N} BN{BN} sched{sched}")


@pytest.mark.parametrize("act,res,out_f32", [(1, None, False), (2, None, False), (0, "f32", True), (0, None, True)])
def test_linear_pair_epilogues_auto_bn(native_lib, act, res, out_f32):
    out, ref = _lin(native_lib, 1700, 384, 384, act=act, res=res, out_f32=out_f32, resident=4)
    _check(out, ref, 384, f"pair linear act{act} res{res} f32{out_f32}")


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,taps,dil,BN,sched", [
    (1, 32, 48, 64, 0, 64, 9, 1, 64, 2), (2, 24, 38, 128, 0, 256, 9, 1, 256, 2), (1, 16, 16, 512, 0, 1024, 9, 6, 256, 2),
    (1, 48, 38, 1024, 512, 512, 1, 1, 256, 2), (1, 64, 64, 32, 0, 32, 9, 1, 32, 3), (3, 8, 8, 256, 128, 128, 1, 1, 128, 2),
    (3, 40, 24, 64, 0, 64, 9, 1, 64, 3), (1, 8, 16, 64, 0, 128, 9, 1, 128, 2), (5, 24, 24, 128, 0, 128, 9, 1, 0, 4),
])
def test_conv_pair(native_lib, B, H, W, C0, C1, Cout, taps, dil, BN, sched):
    out, ref = _conv(native_lib, B, H, W, C0, C1, Cout, taps, dil, BN=BN, resident=sched)
    _check(out.reshape(-1, Cout), ref.reshape(-1, Cout), taps * (C0 + C1), f"pair conv {B}x{H}x{W} C{C0}+{C1}->{Cout} t{taps} d{dil} BN{BN} s{sched}")


def test_linear_auto_plan_large(native_lib):
    for (M, K, N) in [(131072, 384, 1152), (131072, 1536, 384), (2400, 384, 384), (26 * 300, 384, 96)]:
        out, ref = _lin(native_lib, M, K, N)
        _check(out, ref, K, f"auto linear M{M} K{K} N{N}")


def test_conv_weight_resident(native_lib):
    for (B, H, W, C0, C1, Cout, taps, BN) in [(2, 64, 64, 64, 0, 64, 9, 64), (1, 128, 96, 128, 256, 128, 1, 128),
                                              (1, 96, 128, 32, 0, 32, 9, 32)]:
        out, ref = _conv(native_lib, B, H, W, C0, C1, Cout, taps, 1, BN=BN, resident=1)
        _check(out.reshape(-1, Cout), ref.reshape(-1, Cout), taps * (C0 + C1), f"resident conv {H}x{W} C{C0}+{C1}->{Cout}")


@pytest.mark.parametrize("act,res,out_f32", [(1, None, False), (2, None, False), (0, "f32", True), (0, None, True)])
def test_linear_epilogues(native_lib, act, res, out_f32):
    out, ref = _lin(native_lib, 700, 384, 384, act=act, res=res, out_f32=out_f32)
    _check(out, ref, 384, f"linear act{act} res{res} f32{out_f32}")


def _conv(native_lib, B, H, W, C0, C1, Cout, taps, dil, relu=1, BN=0, seed=0, resident=0):
    from tuatara_b200._native import check

    g = torch.Generator(device="cpu").manual_seed(seed)
    x0 = (torch.randn(B, H, W, C0, generator=g)).to(torch.bfloat16).cuda()
    x1 = (torch.randn(B, H, W, C1, generator=g)).to(torch.bfloat16).cuda() if C1 else None
    ct = C0 + C1
    k = 3 if taps == 9 else 1
    w = (torch.randn(Cout, k, k, ct, generator=g) * (1.0 / np.sqrt(taps * ct))).to(torch.bfloat16).cuda()
    b = torch.randn(Cout, generator=g).float().cuda()
    out = torch.full((B, H, W, Cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    check(native_lib.tt_conv_dev(x0.data_ptr(), C0, x1.data_ptr() if C1 else None, C1, B, H, W, taps, dil,
                                 w.data_ptr(), b.data_ptr(), Cout, relu, out.data_ptr(), BN, resident, None), "tt_conv_dev")
    torch.cuda.synchronize()
    xin = x0.float() if not C1 else torch.cat([x0.float(), x1.float()], -1)
    ref = torch.nn.functional.conv2d(xin.permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b,
                                     padding=(dil if taps == 9 else 0), dilation=dil)
    if relu:
        ref = torch.relu(ref)
    return out.float().cpu(), ref.permute(0, 2, 3, 1).contiguous().cpu()


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,taps,dil", [
    (1, 32, 48, 64, 0, 64, 9, 1), (2, 24, 38, 128, 0, 256, 9, 1), (1, 16, 16, 512, 0, 1024, 9, 6),
    (1, 48, 38, 1024, 512, 512, 1, 1), (1, 64, 64, 32, 0, 32, 9, 1), (1, 40, 24, 32, 0, 16, 9, 1),
    (1, 64, 80, 32, 0, 64, 1, 1), (1, 14, 18, 64, 0, 128, 9, 1), (3, 8, 8, 256, 128, 128, 1, 1),
])
def test_conv_shapes(native_lib, B, H, W, C0, C1, Cout, taps, dil):
    out, ref = _conv(native_lib, B, H, W, C0, C1, Cout, taps, dil)
    _check(out.reshape(-1, Cout), ref.reshape(-1, Cout), taps * (C0 + C1), f"conv {B}x{H}x{W} C{C0}+{C1}->{Cout} t{taps} d{dil}")


@pytest.mark.parametrize("M,K,N", [(2400, 384, 384), (307200 // 8, 1536, 384), (1000, 96, 384), (129, 384, 128)])
def test_linear_residual_in_place(native_lib, M, K, N):
    """x += A W^T + b with out == residual (the PARSeq residual stream; TMA epilogue: residual chunks are loaded,
    updated in place in smem and stored back to the same addresses)."""
    from tuatara_b200._native import check

    g = torch.Generator(device="cpu").manual_seed(3)
    A = (torch.randn(M, K, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16).cuda()
    b = torch.randn(N, generator=g).float().cuda()
    x = torch.randn(M, N, generator=g).float().cuda()
    ref = (A.float() @ W.float().t() + b + x).cpu()
    check(native_lib.tt_linear_dev(A.data_ptr(), K, M, K, W.data_ptr(), N, b.data_ptr(), 0, x.data_ptr(), 1, N,
                                   x.data_ptr(), 1, N, 0, 0, None), "tt_linear_dev")
    torch.cuda.synchronize()
    _check(x.cpu(), ref, K, f"in-place residual linear M{M} K{K} N{N}")


@pytest.mark.parametrize("B,H,W,C0,Cout", [
    (1, 32, 48, 64, 64), (2, 40, 24, 64, 32), (1, 64, 64, 128, 128), (3, 17, 9, 128, 64), (1, 128, 136, 64, 64),
    (1, 16, 8, 64, 16), (2, 50, 70, 128, 64), (8, 64, 64, 64, 128), (1, 64, 64, 32, 32), (2, 33, 41, 32, 16), (1, 96, 128, 32, 64),
])
def test_conv_halo(native_lib, B, H, W, C0, Cout):
    """3x3 convs with resident weights run in halo mode: one staged (16+2) x 16-pixel tile per 64 channels feeds all
    nine taps (the tap's A operand is that tile read at a pixel offset); borders, odd sizes, 1 and 2 channel blocks."""
    out, ref = _conv(native_lib, B, H, W, C0, 0, Cout, 9, 1)
    _check(out.reshape(-1, Cout), ref.reshape(-1, Cout), 9 * C0, f"halo conv {B}x{H}x{W} C{C0}->{Cout}")


@pytest.mark.parametrize("M,K1,N2,act", [(307200 // 8, 384, 1152, 0), (40000, 1536, 1536, 2), (1000, 96, 768, 0), (129, 384, 1536, 2),
                                          (128 * 300, 384, 768, 0)])
def test_linear_layernorm_fused_pair(native_lib, M, K1, N2, act):
    """LayerNorm fused away (gemm_tc.cuh Epilogue::ln_*): the residual GEMM emits bf16(x) and the rows' (sum, sum of
    squares) and keeps the stream as a (hi, lo) bf16 pair, the next GEMM reads hi and applies (mean, rstd) to its
    accumulators.  Reference: fp32 LayerNorm + Linear in torch.
    Sizes cover the GPU-filling and the small-launch regime, a ragged last tile and the K=96 patch-embedding shape."""
    from tuatara_b200._native import check

    D = 384
    g = torch.Generator(device="cpu").manual_seed(M + N2)
    A = (torch.randn(M, K1, generator=g) * 0.5).to(torch.bfloat16).cuda()
    W1 = (torch.randn(D, K1, generator=g) * 0.05).to(torch.bfloat16).cuda()
    b1 = torch.randn(D, generator=g).float().cuda() * 0.1
    X0 = (torch.randn(M, D, generator=g) * 1.5 + 0.3).float().cuda()  # residual stream with a non-zero mean
    gamma = (1.0 + 0.2 * (torch.rand(D, generator=g) - 0.5)).cuda()
    beta = (0.05 * torch.randn(D, generator=g)).cuda()
    W2 = (torch.randn(N2, D, generator=g) * 0.05).cuda()
    b2 = (torch.randn(N2, generator=g) * 0.02).cuda()
    W2f = (W2 * gamma[None, :]).to(torch.bfloat16)
    c1 = W2f.double().sum(1).float()
    c0 = (b2.double() + W2.double() @ beta.double()).float()
    XH = X0.to(torch.bfloat16)
    XL = (X0 - XH.float()).to(torch.bfloat16)   # the split residual stream: x = hi + lo
    stats = torch.zeros(M, 8, device="cuda")
    out = torch.full((M, N2), float("nan"), dtype=torch.bfloat16, device="cuda")
    x_in = XH.float() + XL.float()
    check(native_lib.tt_linear_ln_pair_dev(A.data_ptr(), M, K1, W1.data_ptr(), b1.data_ptr(), D, XH.data_ptr(), XL.data_ptr(),
                                           stats.data_ptr(), W2f.data_ptr(), c0.data_ptr(), c1.data_ptr(), N2, act, 1e-6,
                                           out.data_ptr(), None), "tt_linear_ln_pair_dev")
    torch.cuda.synchronize()
    X = XH.float() + XL.float()
    x_ref = x_in + A.float() @ W1.float().t() + b1
    _check(X.cpu(), x_ref.cpu(), K1, "residual stream")
    # the pair carries x to ~2^-16 relative: far inside the GEMM's own accumulation-order noise
    assert float((X - x_ref).abs().max() / x_ref.abs().max()) < 2e-3
    # hi is a nearest bf16 of x (what the consumer GEMM reads), lo the bf16-rounded remainder
    assert bool(((XH.float() - X).abs() <= X.abs() * 2.0 ** -8 + 1e-30).all()), "hi = bf16(x)"
    y = torch.nn.functional.layer_norm(X, (D,), gamma, beta, 1e-6) @ W2.t() + b2   # from the kernel's own x: isolates the LN + GEMM
    if act == 2:
        y = torch.nn.functional.gelu(y)
    err = (out.float() - y).norm() / y.norm()
    print("LN-fused pair rel-L2", float(err))
    assert torch.isfinite(out.float()).all()
    assert float(err) <= 6e-3, float(err)


@pytest.mark.parametrize("proj", [False, True])
@pytest.mark.parametrize("M", [128, 384, 2432, 38400])
def test_fused_encoder_mlp_kernel(native_lib, M, proj):
    """k_enc_mlp (enc_mlp.cu): the second half of a PARSeq-base encoder block in one kernel -- optionally the attention
    output projection (x += att Wp^T + bp, the residual rows added into the TMEM accumulator through an identity
    operand), then x += fc2(GELU(fc1(LN(x)))): CTA pairs, the hidden activations go from registers to the smem tile fc2
    reads, the residual stream stays the split (hi, lo) bf16 pair, the rows' LayerNorm sums are emitted for the next layer.
    Reference: fp32 torch.  Sizes: one half-filled pair tile, an odd number of 128-row tiles (the pair's second CTA past
    the end), more tiles than CTA pairs (several tiles per pair, ragged), and a batch of 300 crops."""
    from tuatara_b200._native import check

    D, H = 384, 1536
    g = torch.Generator(device="cpu").manual_seed(M + int(proj))
    X0 = (torch.randn(M, D, generator=g) * 1.5 + 0.3).float().cuda()
    gamma = (1.0 + 0.2 * (torch.rand(D, generator=g) - 0.5)).cuda()
    beta = (0.05 * torch.randn(D, generator=g)).cuda()
    W1 = (torch.randn(H, D, generator=g) * 0.05).cuda()
    b1 = (torch.randn(H, generator=g) * 0.02).cuda()
    W2 = (torch.randn(D, H, generator=g) * 0.03).to(torch.bfloat16).cuda()
    b2 = (torch.randn(D, generator=g) * 0.1).float().cuda()
    att = (torch.randn(M, D, generator=g) * 0.7).to(torch.bfloat16).cuda()
    Wp = (torch.randn(D, D, generator=g) * 0.05).to(torch.bfloat16).cuda()
    bp = (torch.randn(D, generator=g) * 0.1).float().cuda()
    W1f = (W1 * gamma[None, :]).to(torch.bfloat16)
    c1 = W1f.double().sum(1).float()
    c0 = (b1.double() + W1.double() @ beta.double()).float()
    XH = X0.to(torch.bfloat16)
    XL = (X0 - XH.float()).to(torch.bfloat16)
    x_in = XH.float() + XL.float()
    stats = torch.zeros(M, 4, device="cuda")
    if not proj:   # with the projection the kernel forms the LayerNorm sums of x1 itself
        stats[:, 0] = x_in[:, :192].sum(1); stats[:, 1] = (x_in[:, :192] ** 2).sum(1)
        stats[:, 2] = x_in[:, 192:].sum(1); stats[:, 3] = (x_in[:, 192:] ** 2).sum(1)
    check(native_lib.tt_enc_mlp_dev(XH.data_ptr(), XL.data_ptr(), stats.data_ptr(), M, att.data_ptr() if proj else None,
                                    Wp.data_ptr() if proj else None, bp.data_ptr() if proj else None, W1f.data_ptr(), c0.data_ptr(),
                                    c1.data_ptr(), W2.data_ptr(), b2.data_ptr(), 1e-6, None), "tt_enc_mlp_dev")
    torch.cuda.synchronize()
    X = XH.float() + XL.float()
    x1 = x_in + (att.float() @ Wp.float().t() + bp if proj else 0.0)
    h = torch.nn.functional.gelu(torch.nn.functional.layer_norm(x1, (D,), gamma, beta, 1e-6) @ W1.t() + b1)
    x_ref = x1 + h @ W2.float().t() + b2
    assert torch.isfinite(X).all()
    err = float((X - x_ref).norm() / x_ref.norm())
    upd = float(((X - x_in) - (x_ref - x_in)).norm() / (x_ref - x_in).norm())
    print(f"M={M} proj={proj}: fused block x' rel-L2 {err:.2e}, update rel-L2 {upd:.2e}")
    assert err <= 2e-3 and upd <= 8e-3, (err, upd)
    assert bool(((XH.float() - X).abs() <= X.abs() * 2.0 ** -8 + 1e-30).all()), "hi = bf16(x)"
    # the emitted partial sums are those of the new rows
    s1, s2 = stats[:, 0] + stats[:, 2], stats[:, 1] + stats[:, 3]
    assert torch.allclose(s1, X.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(s2, (X ** 2).sum(1), rtol=1e-4, atol=1e-2)
