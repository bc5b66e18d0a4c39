"""GPU (B200): the tcgen05 / TMA GEMM (acx_gemm_bf16) against fp64 matmul of the same bf16 operands
and against the SIMT fp32 GEMM of the same library."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from audioset_convnext_inf_b200 import _native as N          # noqa: E402

DEV = "cuda:0"

SHAPES = [
    # (M, N, K)  -- every (BN, K-tail, M-tail) combination the trunk uses
    (128, 96, 64), (128, 128, 64), (256, 256, 128), (300, 192, 96),
    (14112, 384, 96), (14112, 96, 384), (3528, 768, 192), (3528, 192, 768),
    (882, 1536, 384), (882, 384, 1536), (434, 3072, 768), (434, 768, 3072),
    (7056, 192, 384), (1764, 384, 768), (434, 768, 1536),
]


def _run(M, Nn, K, epi, seed=0):
    g = torch.Generator().manual_seed(seed + M + 3 * Nn + 7 * K)
    A = (torch.randn(M, K, generator=g)).to(torch.bfloat16)
    W = (torch.randn(Nn, K, generator=g) * (1.0 / K ** 0.5)).to(torch.bfloat16)
    bias = torch.randn(Nn, generator=g) * 0.1
    gamma = torch.rand(Nn, generator=g)
    resid = torch.randn(M, Nn, generator=g).to(torch.bfloat16)
    Ad, Wd, bd, gd, rd = (t.to(DEV) for t in (A, W, bias, gamma, resid))
    ref = (Ad.double() @ Wd.double().t()) + bd.double()
    if epi == N.EPI_BIAS_GELU:
        ref = F.gelu(ref)
    if epi == N.EPI_BIAS_SCALE_RESID:
        ref = rd.double() + gd.double() * ref
    out = torch.full((M, Nn), float("nan"), device=DEV, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    N.call("acx_gemm_bf16", Ad.data_ptr(), Wd.data_ptr(), out.data_ptr(), M, Nn, K, epi, bd.data_ptr(), gd.data_ptr(),
           rd.data_ptr(), st)
    torch.cuda.synchronize()
    return out.double(), ref


@pytest.mark.parametrize("M,Nn,K", SHAPES)
def test_umma_gemm_bias(M, Nn, K):
    out, ref = _run(M, Nn, K, N.EPI_BIAS)
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item()
    assert err < 0.02 + 0.01 * ref.abs().max().item(), err   # bf16 output rounding


@pytest.mark.parametrize("M,Nn,K", [(14112, 384, 96), (882, 1536, 384), (434, 3072, 768), (300, 192, 96)])
def test_umma_gemm_gelu(M, Nn, K):
    out, ref = _run(M, Nn, K, N.EPI_BIAS_GELU)
    err = (out - ref).abs().max().item()
    assert err < 0.02 + 0.01 * ref.abs().max().item(), err


@pytest.mark.parametrize("M,Nn,K", [(14112, 96, 384), (3528, 192, 768), (882, 384, 1536), (434, 768, 3072)])
def test_umma_gemm_scale_residual_in_place(M, Nn, K):
    out, ref = _run(M, Nn, K, N.EPI_BIAS_SCALE_RESID)
    err = (out - ref).abs().max().item()
    assert err < 0.03 + 0.01 * ref.abs().max().item(), err
    # in-place form used by the engine: out aliases resid
    g = torch.Generator().manual_seed(5)
    A = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
    W = (torch.randn(Nn, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    bias = torch.zeros(Nn, device=DEV)
    gamma = torch.ones(Nn, device=DEV)
    x = torch.randn(M, Nn, generator=g).to(torch.bfloat16).to(DEV)
    ref = x.double() + A.double() @ W.double().t()
    N.call("acx_gemm_bf16", A.data_ptr(), W.data_ptr(), x.data_ptr(), M, Nn, K, N.EPI_BIAS_SCALE_RESID,
           bias.data_ptr(), gamma.data_ptr(), x.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert (x.double() - ref).abs().max().item() < 0.03 + 0.01 * ref.abs().max().item()


def test_umma_gemm_many_tiles_persistent_loop():
    """More tiles than SMs x 2 accumulators: exercises the phase bookkeeping of both pipelines."""
    out, ref = _run(128 * 700, 192, 192, N.EPI_BIAS)
    assert (out - ref).abs().max().item() < 0.02 + 0.01 * ref.abs().max().item()


def test_umma_rejects_bad_arguments():
    lib = N.load()
    assert lib.acx_gemm_bf16(0, 0, 0, 16, 96, 64, 0, 0, 0, 0, 0) != 0
    assert "null" in N.last_error()
    t = torch.zeros(128, 100, device=DEV, dtype=torch.bfloat16)
    b = torch.zeros(100, device=DEV)
    rc = lib.acx_gemm_bf16(t.data_ptr(), t.data_ptr(), t.data_ptr(), 128, 100, 64, 0, b.data_ptr(), 0, 0, 0)
    assert rc != 0 and "multiple" in N.last_error()


# the last three cases give every persistent CTA 6-12 row tiles (ragged tail included), so the D1 ring (4 buffers), the
# per-group hidden buffers and the double-buffered D2 of mlp_fused96 wrap their mbarrier phases several times
@pytest.mark.parametrize("C,M", [(96, 128), (96, 1000), (96, 14112 * 2), (192, 300), (192, 3528 * 2), (96, 128 * 400),
                                 (96, 128 * 148 * 6 + 77), (96, 14112 * 16), (192, 3528 * 16)])
def test_mlp_fused_matches_reference_block_mlp(C, M):
    """acx_mlp_fused (hidden tile kept in TMEM/SMEM) vs fp64 evaluation of CX:79-86 on the same bf16 operands,
    and vs the two-kernel tcgen05 path (pw1+GELU, pw2+gamma+residual)."""
    g = torch.Generator().manual_seed(C + M)
    y = torch.randn(M, C, generator=g).to(torch.bfloat16)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(4 * C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) * 0.5 + 0.1
    yd, xd, w1d, w2d, b1d, b2d, gd = (t.to(DEV) for t in (y, x, w1, w2, b1, b2, gamma))
    hid = F.gelu(yd.double() @ w1d.double().t() + b1d.double())
    ref = xd.double() + gd.double() * (hid @ w2d.double().t() + b2d.double())
    st = torch.cuda.current_stream().cuda_stream
    # two-kernel path (hidden rounded to bf16 in HBM)
    hbuf = torch.empty(M, 4 * C, device=DEV, dtype=torch.bfloat16)
    x2 = xd.clone()
    N.call("acx_gemm_bf16", yd.data_ptr(), w1d.data_ptr(), hbuf.data_ptr(), M, 4 * C, C, N.EPI_BIAS_GELU, b1d.data_ptr(), 0, 0, st)
    N.call("acx_gemm_bf16", hbuf.data_ptr(), w2d.data_ptr(), x2.data_ptr(), M, C, 4 * C, N.EPI_BIAS_SCALE_RESID,
           b2d.data_ptr(), gd.data_ptr(), x2.data_ptr(), st)
    x1 = xd.clone()
    N.call("acx_mlp_fused", yd.data_ptr(), x1.data_ptr(), w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(),
           gd.data_ptr(), M, C, st)
    torch.cuda.synchronize()
    err = (x1.double() - ref).abs().max().item()
    err2 = (x2.double() - ref).abs().max().item()
    assert torch.isfinite(x1.float()).all()
    assert err < 0.03 + 0.01 * ref.abs().max().item(), (err, err2)
    assert (x1.double() - x2.double()).abs().max().item() < 0.05      # same arithmetic up to bf16 rounding of outputs


@pytest.mark.parametrize("C,M", [(96, 128 * 150 + 77), (96, 100), (192, 128 * 149 + 5), (192, 4000)])
def test_mlp_fused_ln_equals_layernorm_pass_then_mlp(C, M):
    """acx_mlp_fused_ln (LayerNorm applied to the operand tile in shared memory, CX:78-86 in one kernel) against
    acx_layernorm_rows + acx_mlp_fused on the same raw conv output -- same arithmetic up to the summation order of the
    LayerNorm statistics -- and against an fp64 evaluation.  Rows carry a large common offset (the shifted single-pass
    variance must not cancel) and the tile count exceeds the 148 persistent CTAs (several tiles per CTA)."""
    g = torch.Generator().manual_seed(7 * C + M)
    v = (torch.randn(M, C, generator=g) * torch.rand(M, 1, generator=g) * 3 + torch.randn(M, 1, generator=g) * 20).to(torch.bfloat16)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(4 * C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) * 0.5 + 0.1
    lw = torch.rand(C, generator=g) * 0.4 + 0.8
    lb = torch.randn(C, generator=g) * 0.05
    vd, xd, w1d, w2d, b1d, b2d, gd, lwd, lbd = (t.to(DEV) for t in (v, x, w1, w2, b1, b2, gamma, lw, lb))
    st = torch.cuda.current_stream().cuda_stream
    y = torch.empty_like(vd)
    N.call("acx_layernorm_rows", vd.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), y.data_ptr(), M, C, st)
    x_pass = xd.clone()
    N.call("acx_mlp_fused", y.data_ptr(), x_pass.data_ptr(), w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(),
           gd.data_ptr(), M, C, st)
    x_fused = xd.clone()
    v_before = vd.clone()
    N.call("acx_mlp_fused_ln", vd.data_ptr(), x_fused.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), w1d.data_ptr(), b1d.data_ptr(),
           w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    torch.cuda.synchronize()
    assert torch.equal(vd, v_before)                                   # the HBM copy of v is never written
    yn = F.layer_norm(vd.double(), (C,), lwd.double(), lbd.double(), 1e-6)
    ref = xd.double() + gd.double() * (F.gelu(yn @ w1d.double().t() + b1d.double()) @ w2d.double().t() + b2d.double())
    assert torch.isfinite(x_fused.float()).all()
    d = (x_fused.double() - x_pass.double()).abs()
    assert d.max().item() < 0.05 and (d > 0).double().mean().item() < 0.05, (d.max().item(), (d > 0).double().mean().item())
    assert (x_fused.double() - ref).abs().max().item() < 0.05 + 0.01 * ref.abs().max().item()


def _to_gp(t):
    """(M, C) -> group-planar [C/8][M][8]"""
    M, C = t.shape
    Mp = (M + 127) // 128 * 128                      # plane stride: rows rounded up to 128
    out = torch.zeros(C // 8, Mp, 8, device=t.device, dtype=t.dtype)
    out[:, :M] = t.view(M, C // 8, 8).permute(1, 0, 2)
    return out


def _from_gp(t, M, C):
    return t.view(C // 8, -1, 8)[:, :M].permute(1, 0, 2).reshape(M, C).contiguous()


@pytest.mark.parametrize("ln", [False, True])
@pytest.mark.parametrize("C,M", [(96, 128 * 150 + 77), (96, 100), (192, 128 * 149 + 5), (192, 300)])
def test_mlp_fused_group_planar(C, M, ln):
    """acx_mlp_fused_gp: the same fused MLP on group-planar activations ([C/8][M][8]; the operand tile arrives as one 3-D
    TMA box in the un-swizzled canonical K-major layout, residual / output move as 16-byte group pieces) must reproduce
    the row-major kernel bit for bit -- it is the same arithmetic in the same order; this also pins the LBO / SBO field
    assignment of the un-swizzled smem descriptor on hardware."""
    g = torch.Generator().manual_seed(11 * C + M)
    v = (torch.randn(M, C, generator=g) * 2 + torch.randn(M, 1, generator=g) * 5).to(torch.bfloat16)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w1 = (torch.randn(4 * C, C, generator=g) / C ** 0.5).to(torch.bfloat16)
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(4 * C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) * 0.5 + 0.1
    lw = torch.rand(C, generator=g) * 0.4 + 0.8
    lb = torch.randn(C, generator=g) * 0.05
    vd, xd, w1d, w2d, b1d, b2d, gd, lwd, lbd = (t.to(DEV) for t in (v, x, w1, w2, b1, b2, gamma, lw, lb))
    st = torch.cuda.current_stream().cuda_stream
    x_rm = xd.clone()
    if ln:
        N.call("acx_mlp_fused_ln", vd.data_ptr(), x_rm.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), w1d.data_ptr(), b1d.data_ptr(),
               w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    else:
        N.call("acx_mlp_fused", vd.data_ptr(), x_rm.data_ptr(), w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(),
               gd.data_ptr(), M, C, st)
    v_gp, x_gp = _to_gp(vd), _to_gp(xd)
    N.call("acx_mlp_fused_gp", v_gp.data_ptr(), x_gp.data_ptr(), lwd.data_ptr() if ln else 0, lbd.data_ptr() if ln else 0, 0,
           w1d.data_ptr(), b1d.data_ptr(), w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    torch.cuda.synchronize()
    got = _from_gp(x_gp, M, C)
    assert torch.isfinite(got.float()).all()
    d = (got.float() - x_rm.float()).abs().max().item()
    assert torch.equal(got, x_rm), d
    assert torch.equal(v_gp[:, :M], _to_gp(vd)[:, :M])


@pytest.mark.parametrize("C,M", [(96, 128 * 150 + 77), (96, 100), (192, 128 * 149 + 5), (192, 300)])
def test_mlp_fused_group_planar_folded_layernorm(C, M):
    """LayerNorm folded into pwconv1 (engine.fold_layernorm_into_pwconv1): the kernel multiplies the UN-normalised conv
    output by W1' = bf16(W1 ln_w) and applies  rstd (G - mean s) + b1'  in the GELU epilogue.  Checked against an fp64
    evaluation of CX:78-86 on the same bf16 inputs and against the in-shared-memory LayerNorm variant, on rows with a
    common offset several times their spread (where the fold's cancellation is at its worst)."""
    from audioset_convnext_inf_b200.engine import fold_layernorm_into_pwconv1
    g = torch.Generator().manual_seed(13 * C + M)
    v = (torch.randn(M, C, generator=g) * (torch.rand(M, 1, generator=g) * 2 + 0.2) + torch.randn(M, 1, generator=g) * 4).to(torch.bfloat16)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w1 = torch.randn(4 * C, C, generator=g) / C ** 0.5
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(4 * C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) * 0.5 + 0.1
    lw = torch.rand(C, generator=g) * 0.4 + 0.8
    lb = torch.randn(C, generator=g) * 0.05
    f = fold_layernorm_into_pwconv1(w1, b1, lw, lb)
    vd, xd, w2d, b2d, gd, lwd, lbd, b1d = (t.to(DEV) for t in (v, x, w2, b2, gamma, lw, lb, b1))
    w1d = w1.to(torch.bfloat16).to(DEV)
    w1f, s1, b1f = (f[k].to(DEV) for k in ("w1f", "s1", "b1f"))
    st = torch.cuda.current_stream().cuda_stream
    v_gp, x_fold, x_smem = _to_gp(vd), _to_gp(xd), _to_gp(xd)
    N.call("acx_mlp_fused_gp", v_gp.data_ptr(), x_fold.data_ptr(), 0, 0, s1.data_ptr(), w1f.data_ptr(), b1f.data_ptr(),
           w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    N.call("acx_mlp_fused_gp", v_gp.data_ptr(), x_smem.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), 0, w1d.data_ptr(), b1d.data_ptr(),
           w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    torch.cuda.synchronize()
    yn = F.layer_norm(vd.double(), (C,), lwd.double(), lbd.double(), 1e-6)
    ref = xd.double() + gd.double() * (F.gelu(yn @ w1.double().to(DEV).t() + b1d.double()) @ w2d.double().t() + b2d.double())
    got, alt = _from_gp(x_fold, M, C), _from_gp(x_smem, M, C)
    assert torch.isfinite(got.float()).all()
    e_fold = (got.double() - ref).abs().max().item()
    e_smem = (alt.double() - ref).abs().max().item()
    print(f"  folded LN max err {e_fold:.3e}, in-smem LN max err {e_smem:.3e} (ref max {ref.abs().max().item():.2f})")
    assert e_fold < 0.05 + 0.01 * ref.abs().max().item()
    assert e_fold < 2.5 * e_smem + 0.02
    rc = N.load().acx_mlp_fused_gp(v_gp.data_ptr(), x_fold.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), s1.data_ptr(), w1f.data_ptr(),
                                   b1f.data_ptr(), w2d.data_ptr(), b2d.data_ptr(), gd.data_ptr(), M, C, st)
    assert rc != 0 and "excludes" in N.last_error()


@pytest.mark.parametrize("M,Nn,K", [(64 * 126 * 28, 192, 384), (300, 192, 384), (128 * 5 + 17, 96, 64)])
def test_gemm_planar_output_equals_row_major(M, Nn, K):
    """acx_gemm_bf16_gp_out (the downsample GEMM handing its result over group-planar) == acx_gemm_bf16 + layout change."""
    g = torch.Generator().manual_seed(M + Nn)
    a = torch.randn(M, K, generator=g).to(torch.bfloat16).to(DEV)
    w = (torch.randn(Nn, K, generator=g) / K ** 0.5).to(torch.bfloat16).to(DEV)
    b = (torch.randn(Nn, generator=g) * 0.1).to(DEV)
    st = torch.cuda.current_stream().cuda_stream
    out = torch.empty(M, Nn, device=DEV, dtype=torch.bfloat16)
    N.call("acx_gemm_bf16", a.data_ptr(), w.data_ptr(), out.data_ptr(), M, Nn, K, N.EPI_BIAS, b.data_ptr(), 0, 0, st)
    Mp = (M + 127) // 128 * 128
    og = torch.zeros(Nn // 8, Mp, 8, device=DEV, dtype=torch.bfloat16)
    N.call("acx_gemm_bf16_gp_out", a.data_ptr(), w.data_ptr(), og.data_ptr(), M, Nn, K, b.data_ptr(), st)
    torch.cuda.synchronize()
    assert torch.equal(_from_gp(og, M, Nn), out)


@pytest.mark.parametrize("M", [64 * 63 * 14, 3 * 63 * 14, 200])
def test_stage2_planar_two_gemm_mlp(M):
    """Stage 2 on group-planar activations: acx_gp_row_stats + acx_gemm_bf16_pw1_gp (planar A operand, LayerNorm folded
    into the GELU epilogue) + acx_gemm_bf16_pw2_gp (planar residual / output, in place) against an fp64 evaluation of
    CX:78-86 and against the row-major kernels fed with an explicitly normalised operand."""
    from audioset_convnext_inf_b200.engine import fold_layernorm_into_pwconv1
    C = 384
    g = torch.Generator().manual_seed(M)
    v = (torch.randn(M, C, generator=g) * (torch.rand(M, 1, generator=g) * 2 + 0.2) + torch.randn(M, 1, generator=g) * 3).to(torch.bfloat16)
    x = torch.randn(M, C, generator=g).to(torch.bfloat16)
    w1 = torch.randn(4 * C, C, generator=g) / C ** 0.5
    w2 = (torch.randn(C, 4 * C, generator=g) / (4 * C) ** 0.5).to(torch.bfloat16)
    b1 = torch.randn(4 * C, generator=g) * 0.1
    b2 = torch.randn(C, generator=g) * 0.1
    gamma = torch.rand(C, generator=g) * 0.5 + 0.1
    lw = torch.rand(C, generator=g) * 0.4 + 0.8
    lb = torch.randn(C, generator=g) * 0.05
    f = fold_layernorm_into_pwconv1(w1, b1, lw, lb)
    vd, xd, w2d, b2d, gd, lwd, lbd, b1d = (t.to(DEV) for t in (v, x, w2, b2, gamma, lw, lb, b1))
    w1f, s1, b1f = (f[k].to(DEV) for k in ("w1f", "s1", "b1f"))
    st = torch.cuda.current_stream().cuda_stream
    v_gp, x_gp = _to_gp(vd), _to_gp(xd)
    stats = torch.empty(M, 2, device=DEV)
    N.call("acx_gp_row_stats", v_gp.data_ptr(), stats.data_ptr(), M, C, st)
    mean = vd.double().mean(1)
    rstd = 1.0 / torch.sqrt(vd.double().var(1, unbiased=False) + 1e-6)
    assert (stats[:, 0].double() - rstd).abs().max().item() < 1e-4 * rstd.max().item()
    assert (stats[:, 1].double() + mean * rstd).abs().max().item() < 1e-3
    hid = torch.empty(M, 4 * C, device=DEV, dtype=torch.bfloat16)
    N.call("acx_gemm_bf16_pw1_gp", v_gp.data_ptr(), w1f.data_ptr(), hid.data_ptr(), M, 4 * C, C, b1f.data_ptr(), stats.data_ptr(),
           s1.data_ptr(), st)
    N.call("acx_gemm_bf16_pw2_gp", hid.data_ptr(), w2d.data_ptr(), x_gp.data_ptr(), M, C, 4 * C, b2d.data_ptr(), gd.data_ptr(), st)
    torch.cuda.synchronize()
    yn = F.layer_norm(vd.double(), (C,), lwd.double(), lbd.double(), 1e-6)
    href = F.gelu(yn @ w1.double().to(DEV).t() + b1d.double())
    assert (hid.double() - href).abs().max().item() < 0.03 + 0.01 * href.abs().max().item()
    ref = xd.double() + gd.double() * (href @ w2d.double().t() + b2d.double())
    got = _from_gp(x_gp, M, C)
    assert torch.isfinite(got.float()).all()
    assert (got.double() - ref).abs().max().item() < 0.05 + 0.01 * ref.abs().max().item()
    # row-major path on an explicitly normalised operand (what stage 2 ran before): same result up to bf16 roundings
    y = torch.empty_like(vd)
    N.call("acx_layernorm_rows", vd.data_ptr(), lwd.data_ptr(), lbd.data_ptr(), y.data_ptr(), M, C, st)
    hid2 = torch.empty_like(hid)
    x2 = xd.clone()
    w1d = w1.to(torch.bfloat16).to(DEV)
    N.call("acx_gemm_bf16", y.data_ptr(), w1d.data_ptr(), hid2.data_ptr(), M, 4 * C, C, N.EPI_BIAS_GELU, b1d.data_ptr(), 0, 0, st)
    N.call("acx_gemm_bf16", hid2.data_ptr(), w2d.data_ptr(), x2.data_ptr(), M, C, 4 * C, N.EPI_BIAS_SCALE_RESID, b2d.data_ptr(),
           gd.data_ptr(), x2.data_ptr(), st)
    torch.cuda.synchronize()
    assert (got.double() - x2.double()).abs().max().item() < 0.06


def test_mlp_fused_rejects_wide_stages():
    t = torch.zeros(128, 384, device=DEV, dtype=torch.bfloat16)
    f = torch.zeros(1536, device=DEV)
    rc = N.load().acx_mlp_fused(t.data_ptr(), t.data_ptr(), t.data_ptr(), f.data_ptr(), t.data_ptr(), f.data_ptr(),
                                f.data_ptr(), 128, 384, 0)
    assert rc != 0 and "not supported" in N.last_error()
