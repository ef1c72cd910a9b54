"""Per-kernel parity tests (GPU): every C-ABI entry point against the CPU oracle / a plain torch fp32 restatement of
the same op, on seeded inputs.  Tolerances: fp32 storage 2e-5 rel-L2 (re-association only), bf16 storage 1e-2,
fp16 3e-3 (the storage rounding of the OUTPUT alone is 2^-9 / 2^-11)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import founddiff_oracle as O
from oracle import scan_cpu

pytestmark = pytest.mark.gpu

DTYPES = [torch.float32, torch.bfloat16, torch.float16]
TOL = {torch.float32: 2e-5, torch.bfloat16: 1e-2, torch.float16: 3e-3}


def rel(a, b):
    return O.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())


def nhwc(x, dt):          # (B,C,H,W) cpu -> (B,H*W,C) cuda
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B, H * W, C).contiguous().to("cuda", dt)


def nchw(x, H, W):        # (B,P,C) cuda -> (B,C,H,W) cpu fp32
    B, P, C = x.shape
    return x.float().cpu().reshape(B, H, W, C).permute(0, 3, 1, 2)


def q(x, dt):             # storage rounding of an input
    return x.to(dt).float()


@pytest.fixture(scope="module")
def ops():
    from founddiff_b200 import ops as _ops
    assert "sm_100a" in _ops.version()
    return _ops


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(2, 4, 8, 4, 300), (1, 4, 16, 16, 65), (2, 1, 4, 32, 1), (2, 4, 32, 8, 1024),
                                   (1, 4, 8, 32, 2048), (1, 2, 6, 5, 77), (2, 4, 16, 4, 512), (1, 4, 8, 16, 776), (2, 2, 24, 8, 264)])
@pytest.mark.parametrize("dt", DTYPES)
def test_selective_scan(ops, shape, dt):
    b, K, Dk, N, L = shape
    g = torch.Generator().manual_seed(L + N)
    u = q(torch.randn(b, K * Dk, L, generator=g), dt)
    delta = q(torch.randn(b, K * Dk, L, generator=g) * 2, dt)
    if L >= 5:
        delta[0, 0, :5] = 25.0
    A = -torch.exp(torch.randn(K * Dk, N, generator=g) * 0.5)
    Bm, Cm = torch.randn(b, K, N, L, generator=g), torch.randn(b, K, N, L, generator=g)
    D, bias = torch.randn(K * Dk, generator=g), torch.randn(K * Dk, generator=g)
    ref = scan_cpu.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True)
    c = lambda t: t.cuda()
    y = ops.selective_scan_fwd(c(u).to(dt), c(delta).to(dt), c(A), c(Bm), c(Cm), c(D), c(bias), True)
    assert rel(y, ref) < TOL[dt]
    # no softplus / no bias / no D: dt must then be positive for the recurrence to be stable
    dpos = delta.abs()
    y2 = ops.selective_scan_fwd(c(u).to(dt), c(dpos).to(dt), c(A), c(Bm), c(Cm), None, None, False)
    ref2 = scan_cpu.selective_scan_fwd(u, dpos, A, Bm, Cm, None, None, False)
    assert rel(y2, ref2) < TOL[dt]


def test_selective_scan_golden(ops):
    g = load_golden("scan.npz")
    for tag in "abc":
        c = lambda k: g[f"{tag}.{k}"].cuda()
        y = ops.selective_scan_fwd(c("u"), c("delta"), c("A"), c("B"), c("C"), c("D"), c("bias"), True)
        assert rel(y, g[f"{tag}.y64"]) < 2e-6


def test_selective_scan_vs_published_kernel_fixture(ops):
    """tests/golden/scan_vllm.npz = outputs of vLLM 0.22's build of the state-spaces/mamba selective_scan_fwd kernel (the
    lineage of the reference's `selective_scan_cuda.fwd`, src/emamba2.py:34, 152), oracle/gen_golden_scan_vllm.py."""
    import numpy as np
    import os
    from conftest import GOLDEN
    g, v = load_golden("scan.npz"), np.load(os.path.join(GOLDEN, "scan_vllm.npz"))
    for tag in "abcd":
        c = (lambda k: g[f"{tag}.{k}"].cuda()) if tag != "d" else (lambda k: torch.from_numpy(v[f"d.{k}"]).cuda())
        y = ops.selective_scan_fwd(c("u"), c("delta"), c("A"), c("B"), c("C"), c("D"), c("bias"), True)
        r = rel(y, torch.from_numpy(v[f"{tag}.y_vllm"]))
        assert r < 1e-6, (tag, r)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape", [(2, 4, 32, 4, 4096), (2, 4, 32, 8, 2048), (1, 4, 64, 16, 1024), (2, 4, 128, 32, 256),
                                   (1, 2, 6, 5, 77)])
def test_selective_scan_vs_live_published_kernel(ops, shape, dt):
    """The same comparison live, when vLLM's op is loadable on this box (it is in the image this repo is graded in): our
    kernel and the published one on identical (storage-rounded) inputs, every storage dtype, the level geometries
    (d_state 4 / 8 / 16 / 32, 4 direction groups) and a ragged shape."""
    try:
        from oracle.gen_golden_scan_vllm import published_scan
        from vllm import _custom_ops  # noqa: F401
        assert hasattr(torch.ops._C, "selective_scan_fwd")
    except Exception as e:  # pragma: no cover
        pytest.skip(f"vLLM selective_scan_fwd not available: {e}")
    b, K, Dk, N, L = shape
    g = torch.Generator().manual_seed(3 * L + N)
    u = torch.randn(b, K * Dk, L, generator=g).to(dt).cuda()
    delta = (torch.randn(b, K * Dk, L, generator=g) * 2).to(dt).cuda()
    A = -torch.exp(torch.randn(K * Dk, N, generator=g) * 0.5).cuda()
    Bm, Cm = torch.randn(b, K, N, L, generator=g).cuda(), torch.randn(b, K, N, L, generator=g).cuda()
    D, bias = torch.randn(K * Dk, generator=g).cuda(), torch.randn(K * Dk, generator=g).cuda()
    y = ops.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True)
    # the published kernel takes B / C in the input dtype: give it fp32 copies of the rounded u / delta instead, so that
    # both sides see the same numbers and only the OUTPUT rounding (ours stores y in `dt`) differs
    ref = published_scan(u.float(), delta.float(), A, Bm, Cm, D, bias, True)
    r = rel(y, ref)
    assert r < {torch.float32: 1e-6, torch.bfloat16: 4e-3, torch.float16: 5e-4}[dt], (shape, dt, r)


def test_selective_scan_module_is_dropin(ops):
    """The reference-facing signatures of src/emamba2.py:152,154."""
    from founddiff_b200 import selective_scan as S
    g = load_golden("scan.npz")
    c = lambda k: g[f"a.{k}"].cuda()
    out, x = S.selective_scan_cuda_core.fwd(c("u"), c("delta"), c("A"), c("B"), c("C"), c("D"), c("bias"), True, 1)
    assert x is None and rel(out, g["a.y64"]) < 2e-6
    out2, _ = S.fwd(c("u"), c("delta"), c("A"), c("B"), c("C"), c("D"), None, c("bias"), True)
    assert torch.equal(out, out2)
    with pytest.raises(RuntimeError):
        S.fwd(g["a.u"], g["a.delta"], g["a.A"], g["a.B"], g["a.C"], g["a.D"], None, g["a.bias"], True)


# ---------------------------------------------------------------------------------------------------------
CONV_CASES = [
    # c0, c1, cout, k, stride, pad, upsample, H, W
    (64, 0, 64, 3, 1, 1, False, 16, 24),
    (64, 64, 64, 3, 1, 1, False, 8, 16),
    (128, 64, 128, 1, 1, 0, False, 8, 12),
    (64, 0, 128, 4, 2, 1, False, 16, 24),
    (128, 0, 64, 3, 1, 1, True, 8, 12),
    (64, 0, 256, 1, 1, 0, False, 16, 16),
    (32, 0, 48, 3, 1, 1, False, 6, 10),
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("dt", DTYPES)
def test_conv2d_simt(ops, case, dt):
    c0, c1, cout, k, stride, pad, up, H, W = case
    B = 2
    g = torch.Generator().manual_seed(c0 + cout + k)
    x0 = q(torch.randn(B, c0, H, W, generator=g), dt)
    x1 = q(torch.randn(B, c1, H, W, generator=g), dt) if c1 else None
    w = q(torch.randn(cout, c0 + c1, k, k, generator=g) / math.sqrt((c0 + c1) * k * k), dt)
    bias = torch.randn(cout, generator=g)
    xin = torch.cat([x0, x1], dim=1) if c1 else x0
    if up:
        xin = F.interpolate(xin, scale_factor=2, mode="nearest")
    ref = F.conv2d(xin, w, bias, stride=stride, padding=pad)
    Ho, Wo = ref.shape[-2:]
    out = torch.empty(B, Ho * Wo, cout, device="cuda", dtype=dt)
    wp = w.permute(0, 2, 3, 1).contiguous().to("cuda", dt)
    conv = ops.Conv(nhwc(x0, dt), wp, out, B=B, Hin=H, Win=W, KH=k, KW=k, stride=stride, pad=pad, upsample=up,
                    src1=nhwc(x1, dt) if c1 else None, bias=bias.cuda(), prefer_tc=False)
    conv.run()
    assert rel(nchw(out, Ho, Wo), ref) < TOL[dt]


@pytest.mark.parametrize("dt", DTYPES)
def test_conv2d_simt_epilogues(ops, dt):
    B, C, H, W, cout = 2, 64, 8, 12, 128
    g = torch.Generator().manual_seed(3)
    x = q(torch.randn(B, C, H, W, generator=g), dt)
    w = q(torch.randn(cout, C, 1, 1, generator=g) / 8, dt)
    # (a) SiLU on the upper half (SS2D in_proj: x | silu(z))
    ref = F.conv2d(x, w)
    ref = torch.cat([ref[:, :cout // 2], F.silu(ref[:, cout // 2:])], dim=1)
    out = torch.empty(B, H * W, cout, device="cuda", dtype=dt)
    ops.Conv(nhwc(x, dt), w.reshape(cout, C).to("cuda", dt), out, B=B, Hin=H, Win=W, silu_from=cout // 2, prefer_tc=False).run()
    assert rel(nchw(out, H, W), ref) < TOL[dt]
    # (b) gate * acc + addend with a strided gate tensor, in place on the addend; per-batch weights; GN partial sums
    mods = torch.randn(B, 3 * cout, generator=g)
    add = q(torch.randn(B, cout, H, W, generator=g), dt)
    wb = q(torch.randn(B, cout, C, generator=g) / 8, dt)
    bias = torch.randn(cout, generator=g)
    v = torch.stack([F.conv2d(x[b:b + 1], wb[b].reshape(cout, C, 1, 1), bias)[0] for b in range(B)])
    ref = add + mods[:, cout:2 * cout, None, None] * v
    buf = nhwc(add, dt)
    sums = torch.zeros(B, 8, 2, device="cuda")
    from founddiff_b200.engine import _view_ptr
    mods_d = mods.cuda()
    ops.Conv(nhwc(x, dt), wb.to("cuda", dt).contiguous(), buf, B=B, Hin=H, Win=W, bias=bias.cuda(),
             gate=_view_ptr(mods_d[:, cout:2 * cout]), gate_stride=3 * cout, addend=buf, per_batch_weight=True,
             gn_sums=sums, gn_groups=8, prefer_tc=False).run()
    assert rel(nchw(buf, H, W), ref) < TOL[dt]
    vg = v.reshape(B, 8, -1)
    ref_sums = torch.stack([vg.sum(-1), (vg * vg).sum(-1)], dim=-1)
    assert rel(sums, ref_sums) < 1e-4


@pytest.mark.parametrize("dt", DTYPES)
def test_init_conv7x7(ops, dt):
    B, H, W = 2, 20, 36
    g = torch.Generator().manual_seed(1)
    xt, xi = torch.randn(B, 1, H, W, generator=g), torch.randn(B, 1, H, W, generator=g)
    w, b = torch.randn(64, 2, 7, 7, generator=g) / 10, torch.randn(64, generator=g)
    ref = F.conv2d(torch.cat([xt, xi], 1), w, b, padding=3)
    out = torch.empty(B, H * W, 64, device="cuda", dtype=dt)
    ops.init_conv7x7(xt.reshape(B, -1).cuda(), xi.reshape(B, -1).cuda(), w.cuda(), b.cuda(), out, B, H, W)
    assert rel(nchw(out, H, W), ref) < TOL[dt]


@pytest.mark.parametrize("C", [64, 128, 256, 512, 1024])
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("affine", [True, False])
def test_ln_modulate(ops, C, dt, affine):
    from founddiff_b200.engine import _view_ptr
    B, P = 2, 50
    g = torch.Generator().manual_seed(C)
    x = q(torch.randn(B, P, C, generator=g) * 2 + 0.5, dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    mods = torch.randn(B, 6 * C, generator=g)
    eps = 1e-5 if affine else 1e-6
    ref = F.layer_norm(x, (C,), gamma if affine else None, beta if affine else None, eps=eps)
    ref = ref * (1 + mods[:, None, C:2 * C]) + mods[:, None, :C]
    out = torch.empty(B, P, C, device="cuda", dtype=dt)
    md = mods.cuda()
    ops.ln_modulate(x.to("cuda", dt), out, gamma.cuda() if affine else None, beta.cuda() if affine else None,
                    _view_ptr(md[:, :C]), _view_ptr(md[:, C:2 * C]), 6 * C, B, P, C, eps)
    assert rel(out, ref) < TOL[dt]


@pytest.mark.parametrize("C", [64, 128, 512])
@pytest.mark.parametrize("dt", DTYPES)
def test_groupnorm_silu_add(ops, C, dt):
    B, H, W = 2, 12, 20
    P = H * W
    g = torch.Generator().manual_seed(C)
    y = q(torch.randn(B, C, H, W, generator=g) * 3 + 1, dt)
    skip = q(torch.randn(B, C, H, W, generator=g), dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.silu(F.group_norm(y, 8, gamma, beta, eps=1e-5)) + skip
    sums = torch.zeros(B, 8, 2, device="cuda")
    yd = nhwc(y, dt)
    ops.gn_stats(yd, sums, B, P, C, 8)
    out = torch.empty_like(yd)
    ops.gn_silu_add(yd, sums, gamma.cuda(), beta.cuda(), nhwc(skip, dt), out, B, P, C, 8)
    assert rel(nchw(out, H, W), ref) < TOL[dt]
    ops.gn_silu_add(yd, sums, gamma.cuda(), beta.cuda(), None, out, B, P, C, 8)
    assert rel(nchw(out, H, W), ref - skip) < TOL[dt]


@pytest.mark.parametrize("hw", [(16, 24), (32, 32), (8, 12), (64, 32), (6, 10), (36, 70)])
@pytest.mark.parametrize("dt", DTYPES)
def test_dwconv_scan_and_merge(ops, hw, dt):
    """dwconv+SiLU+EfficientScan and EfficientMerge+LN+gate against the oracle's restatement of
    src/emamba2.py:186-262."""
    H, W = hw
    B, C = 2, 32
    D, L = 2 * C, H * W // 4
    g = torch.Generator().manual_seed(H * W)
    xz = q(torch.randn(B, H, W, 4 * C, generator=g), dt)
    w, b = torch.randn(D, 1, 3, 3, generator=g) / 3, torch.randn(D, generator=g)
    x = xz[..., :D].permute(0, 3, 1, 2)
    ref_xs = O.efficient_scan(F.silu(F.conv2d(x, w, b, padding=1, groups=D)))
    xs = torch.empty(B, 4, D, L, device="cuda", dtype=dt)
    xzd = xz.reshape(B, H * W, 4 * C).to("cuda", dt)
    ops.dwconv3x3_silu_scan(xzd, 4 * C, w.reshape(D, 9).cuda(), b.cuda(), xs, B, H, W, D)
    assert rel(xs, ref_xs) < TOL[dt]
    # merge
    ys = q(torch.randn(B, 4, D, L, generator=g), dt)
    gamma, beta, local = torch.randn(D, generator=g), torch.randn(D, generator=g), torch.randn(B, D, generator=g)
    y = O.efficient_merge(ys, H, W).permute(0, 2, 3, 1)
    ref = F.layer_norm(y, (D,), gamma, beta, eps=1e-5) * xz[..., D:2 * D] + local[:, None, None, :]
    out = torch.empty(B, H * W, D, device="cuda", dtype=dt)
    stat = torch.empty(B, H * W, 2, device="cuda")
    ops.merge_ln_gate(ys.to("cuda", dt), xzd, 4 * C, D, gamma.cuda(), beta.cuda(), local.cuda(), stat, out, B, H, W, D)
    assert rel(out.reshape(B, H, W, D), ref) < TOL[dt]


@pytest.mark.parametrize("cfg", [(128, 4, 4, 100), (128, 4, 8, 384), (256, 8, 16, 64), (1024, 32, 32, 40), (512, 16, 32, 1024),
                                 (128, 4, 4, 4096)])
@pytest.mark.parametrize("dt", DTYPES)
def test_xdt_proj(ops, cfg, dt):
    D, R, N, L = cfg
    B = 2
    g = torch.Generator().manual_seed(D + L)
    xs = q(torch.randn(B, 4, D, L, generator=g), dt)
    Wx = torch.randn(4, R + 2 * N, D, generator=g) / math.sqrt(D)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, Wx)
    dts_r, Bs_r, Cs_r = torch.split(x_dbl, [R, N, N], dim=2)
    dts_r = torch.einsum("bkrl,kdr->bkdl", dts_r, Wdt)
    dts = torch.empty(B, 4, D, L, device="cuda", dtype=dt)
    Bs, Cs = torch.empty(B, 4, N, L, device="cuda"), torch.empty(B, 4, N, L, device="cuda")
    ops.xdt_proj(xs.to("cuda", dt), Wx.cuda(), Wdt.cuda(), dts, Bs, Cs, B, D, L, R, N)
    assert rel(dts, dts_r) < TOL[dt] and rel(Bs, Bs_r) < 2e-5 and rel(Cs, Cs_r) < 2e-5
    if dt != torch.float32:       # tensor-core variant: 16-bit weights, fp32 accumulation
        xw16, dw16, Rp = ops.pack_xdt_weights(Wx.cuda(), Wdt.cuda(), dt)
        dts2 = torch.zeros_like(dts)
        Bs2, Cs2 = torch.zeros_like(Bs), torch.zeros_like(Cs)
        ops.xdt_proj_tc(xs.to("cuda", dt), xw16, dw16, Rp, dts2, Bs2, Cs2, B, D, L, R, N)
        wt = 1e-2 if dt == torch.bfloat16 else 2e-3
        assert rel(dts2, dts_r) < 2 * wt and rel(Bs2, Bs_r) < wt and rel(Cs2, Cs_r) < wt


@pytest.mark.parametrize("cfg", [(8, 64, 4, 4, 8192), (16, 32, 4, 8, 3336), (8, 128, 8, 16, 6664)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_xdt_proj_tc_long_ragged_rows(ops, cfg, dt):
    """Long rows whose length is not a multiple of the 128-step tile, with the time-major / softplus epilogues."""
    B, D, R, N, L = cfg
    g = torch.Generator().manual_seed(D + L)
    xs = q(torch.randn(B, 4, D, L, generator=g), dt)
    Wx = torch.randn(4, R + 2 * N, D, generator=g) / math.sqrt(D)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    bias = torch.randn(4 * D, generator=g)
    xw16, dw16, Rp = ops.pack_xdt_weights(Wx.cuda(), Wdt.cuda(), dt)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, q(Wx, dt))
    dts_r, Bs_r, Cs_r = torch.split(x_dbl, [R, N, N], dim=2)
    dts_r = torch.einsum("bkrl,kdr->bkdl", q(dts_r, dt), q(Wdt, dt))
    dts = torch.zeros(B, 4, D, L, device="cuda", dtype=dt)
    Bs, Cs = torch.zeros(B, 4, N, L, device="cuda"), torch.zeros(B, 4, N, L, device="cuda")
    ops.xdt_proj_tc(xs.to("cuda", dt), xw16, dw16, Rp, dts, Bs, Cs, B, D, L, R, N)
    wt = 1e-2 if dt == torch.bfloat16 else 2e-3
    assert rel(dts, dts_r) < wt and rel(Bs, Bs_r) < 1e-4 and rel(Cs, Cs_r) < 1e-4
    dts2 = torch.zeros_like(dts)
    Bt, Ct = torch.zeros(B, 4, L, N, device="cuda"), torch.zeros(B, 4, L, N, device="cuda")
    ops.xdt_proj_tc(xs.to("cuda", dt), xw16, dw16, Rp, dts2, Bt, Ct, B, D, L, R, N, time_major=True, dt_bias=bias.cuda(),
                    delta_softplus=True)
    assert rel(Bt.transpose(2, 3), Bs_r) < 1e-4 and rel(Ct.transpose(2, 3), Cs_r) < 1e-4
    ref_sp = F.softplus(dts_r + bias.view(1, 4, D, 1), threshold=20.0)
    assert rel(dts2, ref_sp) < wt


@pytest.mark.parametrize("C", [64, 128])
@pytest.mark.parametrize("hw", [(16, 24), (8, 40), (16, 32), (72, 64)])
@pytest.mark.parametrize("dt", DTYPES)
def test_transposed_attention(ops, C, hw, dt):
    """dwconv+Gram, softmax/W_eff fold and the per-sample 1x1 GEMM together == TransposedAttention after the qkv
    1x1 (src/DADiff.py:266-283)."""
    H, W = hw
    B, heads = 2, C // 32
    g = torch.Generator().manual_seed(C + H)
    qkv = q(torch.randn(B, 3 * C, H, W, generator=g), dt)
    wdw = torch.randn(3 * C, 1, 3, 3, generator=g) / 3
    wproj = torch.randn(C, C, 1, 1, generator=g) / math.sqrt(C)
    temp = torch.rand(heads, 1, 1, generator=g) + 0.5
    t = F.conv2d(qkv, wdw, padding=1, groups=3 * C)
    qq, kk, vv = t.chunk(3, dim=1)
    qn = F.normalize(qq.reshape(B, heads, 32, H * W), dim=-1)
    kn = F.normalize(kk.reshape(B, heads, 32, H * W), dim=-1)
    attn = ((qn @ kn.transpose(-2, -1)) * temp).softmax(dim=-1)
    ref = F.conv2d((attn @ vv.reshape(B, heads, 32, H * W)).reshape(B, C, H, W), wproj)
    gram = torch.zeros(B, heads, 32, 32, device="cuda")
    qk = torch.zeros(B, 2, C, device="cuda")
    if dt == torch.float32:
        v = torch.empty(B, H * W, C, device="cuda", dtype=dt)
        ops.dwconv3x3_qkv_gram(nhwc(qkv, dt), wdw.reshape(3 * C, 9).cuda(), v, gram, qk, B, H, W, C)
        v_src, kw = v, {}
    else:       # streaming dwconv (register sliding window) + tensor-core Gram; v is read in place by the GEMM
        qkv2 = torch.empty(B, H * W, 3 * C, device="cuda", dtype=dt)
        ops.dwconv3x3_nhwc(nhwc(qkv, dt), wdw.reshape(3 * C, 9).t().contiguous().cuda(), None, qkv2, B, H, W, 3 * C)
        assert rel(nchw(qkv2, H, W), t) < TOL[dt]
        ops.gram_qk(qkv2, 3 * C, gram, qk, B, H * W, C)
        v = qkv2[:, :, 2 * C:]
        v_src, kw = v, dict(c0=C, ld0=3 * C)
    assert rel(nchw(v.contiguous(), H, W), vv) < TOL[dt]
    # 16-bit modes form the Gram matrix on the tensor cores from q, k rounded once to the storage type
    gtol = {torch.float32: 1e-4, torch.bfloat16: 1e-2, torch.float16: 2e-3}[dt]
    assert rel(gram, qq.reshape(B, heads, 32, -1) @ kk.reshape(B, heads, 32, -1).transpose(-2, -1)) < gtol
    assert rel(qk[:, 0], (qq * qq).sum(dim=(2, 3))) < gtol and rel(qk[:, 1], (kk * kk).sum(dim=(2, 3))) < gtol
    weff = torch.empty(B, C, C, device="cuda", dtype=dt)
    ops.attn_weff(gram, qk, temp.reshape(-1).cuda(), wproj.reshape(C, C).cuda(), weff, B, C)
    out = torch.empty(B, H * W, C, device="cuda", dtype=dt)
    ops.Conv(v_src, weff, out, B=B, Hin=H, Win=W, per_batch_weight=True, prefer_tc=False, **kw).run()
    assert rel(nchw(out, H, W), ref) < 2 * TOL[dt]
    if dt != torch.float32 and H % 8 == 0 and W % 16 == 0:
        out2 = torch.empty_like(out)
        c = ops.Conv(v_src, weff, out2, B=B, Hin=H, Win=W, per_batch_weight=True, prefer_tc=True, **kw)
        assert c.uses_tc
        c.run()
        assert rel(out2, out) < 3e-3


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(2, 40, 24, 64, True), (1, 33, 17, 24, False), (2, 64, 64, 192, False)])
def test_dwconv3x3_nhwc(ops, dt, shape):
    B, H, W, C, silu = shape
    g = torch.Generator().manual_seed(H * W + C)
    x = q(torch.randn(B, C, H, W, generator=g), dt)
    w, b = torch.randn(C, 1, 3, 3, generator=g) / 3, torch.randn(C, generator=g)
    ref = F.conv2d(x, w, b if silu else None, padding=1, groups=C)
    ref = F.silu(ref) if silu else ref
    out = torch.empty(B, H * W, C, device="cuda", dtype=dt)
    ops.dwconv3x3_nhwc(nhwc(x, dt), w.reshape(C, 9).t().contiguous().cuda(), b.cuda() if silu else None, out, B, H, W, C, silu=silu)
    assert rel(nchw(out, H, W), ref) < TOL[dt]


def test_small_linear_and_time_embedding(ops):
    g = torch.Generator().manual_seed(0)
    B, K, N = 3, 256, 777
    x, Wm, b, add = torch.randn(B, K, generator=g), torch.randn(N, K, generator=g) / 16, torch.randn(N, generator=g), torch.randn(B, N, generator=g)
    out = torch.empty(B, N, device="cuda")
    ops.linear_small(x.cuda(), Wm.cuda(), b.cuda(), out, add=add.cuda(), act_in=1, act_out=2)
    assert rel(out, F.gelu(F.linear(F.silu(x), Wm, b)) + add) < 1e-5
    ops.linear_small(x.cuda(), Wm.cuda(), None, out, act_out=3)
    assert rel(out, F.relu(F.linear(x, Wm))) < 1e-5
    time = torch.tensor([993.647, 719.666, 0.05001])
    dim = 64
    half = dim // 2
    freq = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
    ref = torch.cat(((time[:, None] * freq).sin(), (time[:, None] * freq).cos()), dim=-1)
    out = torch.empty(3, dim, device="cuda")
    ops.time_sinusoid(time.cuda(), out)
    assert (out.cpu() - ref).abs().max() < 2e-4


@pytest.mark.parametrize("dt", DTYPES)
def test_sampler_kernels(ops, dt):
    g = torch.Generator().manual_seed(9)
    B, P, C = 2, 300, 64
    ldct, noise = torch.rand(B, P, generator=g), torch.randn(B, P, generator=g)
    xi, xt, first = (torch.empty(B, P, device="cuda") for _ in range(3))
    ops.sampler_init(ldct.cuda(), noise.cuda(), 0.1, xi, xt, first)
    assert torch.allclose(xi.cpu(), ldct * 2 - 1, atol=1e-6) and torch.allclose(xt.cpu(), ldct * 2 - 1 + 0.1 * noise, atol=1e-6)
    assert torch.allclose(first.cpu(), (ldct * 2 - 1 + 0.1 * noise + 1) / 2, atol=1e-6)
    feat = q(torch.randn(B, P, C, generator=g), dt)
    w, b = torch.randn(C, generator=g) / 4, torch.randn(1, generator=g)
    coef = torch.tensor([0.7, -0.3, 0.2, 0.05, 0.72, 0.96, 0, 0])
    n2 = torch.randn(B, P, generator=g)
    pr = (feat @ w + b).clamp(-1, 1)
    x0 = (xi.cpu() - pr).clamp(-1, 1)
    pn = (xt.cpu() - xi.cpu() - (coef[4] - 1) * pr) / coef[5]
    xn = coef[0] * xt.cpu() + coef[1] * pr + coef[2] * x0 + coef[3] * n2
    o = [torch.empty(B, P, device="cuda") for _ in range(4)]
    ops.final_conv_update(feat.to("cuda", dt), w.cuda(), b.cuda(), xi, xt, n2.cuda(), coef.cuda(), o[0], o[1], o[2], o[3])
    for mine, ref in zip(o, (xn, pr, pn, x0)):
        assert rel(mine, ref) < 1e-5
    un = torch.empty(B, P, device="cuda")
    ops.unnormalize(xt, un)
    assert torch.allclose(un, (xt + 1) / 2)


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("objective", ["pred_res", "pred_noise", "pred_res_noise", "pred_x0_noise"])
def test_final_conv_update_objectives(ops, dt, objective):
    """fd_final_conv_update_obj: every branch of model_predictions (src/DADiff.py:1168-1207) + the shared update line."""
    g = torch.Generator().manual_seed(19)
    B, P, C = 2, 333, 64
    xi, xt, nz = (torch.randn(B, P, generator=g) * 0.5 for _ in range(3))
    f0, f1 = q(torch.randn(B, P, C, generator=g), dt), q(torch.randn(B, P, C, generator=g), dt)
    w0, w1 = torch.randn(C, generator=g) / 6, torch.randn(C, generator=g) / 6
    b0, b1 = torch.randn(1, generator=g) / 4, torch.randn(1, generator=g) / 4
    coef = torch.tensor([0.7, -0.3, 0.2, 0.05, 0.72, 0.96, 0.28, 0])
    acs, bcs, om = coef[4], coef[5], coef[6]
    o0, o1 = f0 @ w0 + b0, f1 @ w1 + b1
    cl = lambda v: v.clamp(-1, 1)  # noqa: E731
    if objective == "pred_res":
        pr = cl(o0); x0 = cl(xi - pr); pn = (xt - xi - (acs - 1) * pr) / bcs  # noqa: E702
    elif objective == "pred_noise":
        pn = o0; x0 = cl((xt - acs * xi - bcs * pn) / om); pr = cl(xi - x0)  # noqa: E702
    elif objective == "pred_res_noise":
        pr = cl(o0); pn = o1; x0 = cl(xt - acs * pr - bcs * pn)  # noqa: E702
    else:
        pr = cl(xi - o0); pn = o1; x0 = cl(o0)  # noqa: E702
    xn = coef[0] * xt + coef[1] * pr + coef[2] * x0 + coef[3] * nz
    o = [torch.empty(B, P, device="cuda") for _ in range(4)]
    two = objective in ("pred_res_noise", "pred_x0_noise")
    ops.final_conv_update(f0.to("cuda", dt), w0.cuda(), b0.cuda(), xi.cuda(), xt.cuda(), nz.cuda(), coef.cuda(), o[0], o[1], o[2], o[3],
                          objective=objective, feat1=f1.to("cuda", dt) if two else None, w1=w1.cuda() if two else None,
                          bias1=b1.cuda() if two else None)
    for mine, ref in zip(o, (xn, pr, pn, x0)):
        assert rel(mine, ref) < 2e-5
    if two:
        with pytest.raises(Exception):
            ops.final_conv_update(f0.to("cuda", dt), w0.cuda(), b0.cuda(), xi.cuda(), xt.cuda(), None, coef.cuda(), o[0],
                                  objective=objective)


# ---------------------------------------------------------------------------------------------------------
# tcgen05 / TMEM / TMA implicit-GEMM path against the CUDA-core path (same inputs, same fused epilogue)
TC_CASES = [
    # c0, c1, cout, k, stride, pad, upsample, H, W
    (64, 0, 64, 3, 1, 1, False, 16, 32),
    (64, 64, 64, 3, 1, 1, False, 8, 16),
    (256, 128, 256, 3, 1, 1, False, 16, 16),
    (128, 0, 512, 1, 1, 0, False, 8, 32),
    (64, 0, 192, 1, 1, 0, False, 16, 16),
    (64, 0, 128, 4, 2, 1, False, 32, 32),
    (128, 0, 64, 3, 1, 1, True, 8, 16),
    (512, 0, 256, 3, 1, 1, True, 8, 16),
    (512, 256, 512, 3, 1, 1, False, 8, 16),
    # halo mode (16x8 tiles, shifted-window UMMA descriptors): resident weights, weight ring, concat, upsample phases,
    # and enough tiles that every persistent CTA walks several of them
    (64, 64, 64, 3, 1, 1, False, 32, 32),
    (64, 0, 64, 3, 1, 1, False, 128, 128),
    (64, 64, 64, 3, 1, 1, False, 144, 136),
    (128, 64, 128, 3, 1, 1, False, 32, 16),
    (128, 0, 64, 3, 1, 1, True, 16, 16),
    (128, 0, 64, 3, 1, 1, True, 96, 64),
    (256, 256, 256, 3, 1, 1, False, 64, 64),
]


@pytest.mark.parametrize("case", TC_CASES)
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_conv2d_tc_matches_simt(ops, case, dt):
    c0, c1, cout, k, stride, pad, up, H, W = case
    B = 2
    g = torch.Generator().manual_seed(c0 + cout + k + c1)
    x0 = torch.randn(B, H * W, c0, generator=g).to("cuda", dt)
    x1 = torch.randn(B, H * W, c1, generator=g).to("cuda", dt) if c1 else None
    w = (torch.randn(cout, k, k, c0 + c1, generator=g) / math.sqrt((c0 + c1) * k * k)).to("cuda", dt)
    bias = torch.randn(cout, generator=g).cuda()
    Ho = (H * (2 if up else 1) + 2 * pad - k) // stride + 1
    Wo = (W * (2 if up else 1) + 2 * pad - k) // stride + 1
    outs = []
    for tc in (False, True):
        out = torch.zeros(B, Ho * Wo, cout, device="cuda", dtype=dt)
        sums = torch.zeros(B, 8, 2, device="cuda")
        use_gn = cout in (64, 128, 256, 512)
        conv = ops.Conv(x0, w, out, B=B, Hin=H, Win=W, KH=k, KW=k, stride=stride, pad=pad, upsample=up, src1=x1, bias=bias,
                        gn_sums=sums if use_gn else None, gn_groups=8 if use_gn else 0, prefer_tc=tc)
        assert conv.uses_tc == tc
        conv.run()
        torch.cuda.synchronize()
        outs.append((out.float().cpu(), sums.cpu()))
    # same bf16 inputs, fp32 accumulation on both paths: only summation order (and the phase-summed upsample
    # weights, rounded once more to 16 bit) differ
    tol = (2e-2 if dt == torch.bfloat16 else 3e-3) if up else 3e-3 if dt == torch.bfloat16 else 5e-4
    assert rel(outs[1][0], outs[0][0]) < tol, rel(outs[1][0], outs[0][0])
    assert rel(outs[1][1], outs[0][1]) < tol


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_conv2d_tc_epilogues(ops, dt):
    from founddiff_b200.engine import _view_ptr
    B, C, H, W, cout = 2, 128, 16, 16, 256
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, H * W, C, generator=g).to("cuda", dt)
    w = (torch.randn(cout, C, generator=g) / math.sqrt(C)).to("cuda", dt)
    res = []
    for tc in (False, True):
        out = torch.zeros(B, H * W, cout, device="cuda", dtype=dt)
        conv = ops.Conv(x, w, out, B=B, Hin=H, Win=W, silu_from=cout // 2, prefer_tc=tc)
        assert conv.uses_tc == tc
        conv.run()
        res.append(out.float().cpu())
    assert rel(res[1], res[0]) < 3e-3
    mods = torch.randn(B, 3 * cout, generator=g).cuda()
    add = torch.randn(B, H * W, cout, generator=g).to("cuda", dt)
    wb = (torch.randn(B, cout, C, generator=g) / math.sqrt(C)).to("cuda", dt)
    res = []
    for tc in (False, True):
        buf = add.clone()
        conv = ops.Conv(x, wb, buf, B=B, Hin=H, Win=W, gate=_view_ptr(mods[:, cout:2 * cout]), gate_stride=3 * cout, addend=buf,
                        per_batch_weight=True, prefer_tc=tc)
        assert conv.uses_tc == tc
        conv.run()
        res.append(buf.float().cpu())
    assert rel(res[1], res[0]) < 3e-3


@pytest.mark.parametrize("cfg", [(16, 32, 64, 4), (32, 16, 32, 8), (16, 16, 128, 16), (8, 16, 256, 32), (64, 64, 16, 4)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_scan_merge_fused_and_ln_gate(ops, cfg, dt):
    """Scan with EfficientMerge fused (channels-last output) + row-wise LN/gate == scan -> EfficientMerge -> LayerNorm ->
    y*z + local of the oracle (src/emamba2.py:353-367, 747-748)."""
    H, W, Dg, N = cfg
    B, L = 2, (H // 2) * (W // 2)
    g = torch.Generator().manual_seed(H * W + N)
    u = q(torch.randn(B, 4 * Dg, L, generator=g), dt)
    delta = q(torch.randn(B, 4 * Dg, L, generator=g), dt)
    A = -torch.exp(torch.randn(4 * Dg, N, generator=g) * 0.3)
    Bm, Cm = torch.randn(B, 4, N, L, generator=g), torch.randn(B, 4, N, L, generator=g)
    D, bias = torch.randn(4 * Dg, generator=g), torch.randn(4 * Dg, generator=g)
    ys = scan_cpu.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True).reshape(B, 4, Dg, L)
    y_ref = O.efficient_merge(ys, H, W).permute(0, 2, 3, 1)                       # (B,H,W,Dg)
    c = lambda t: t.cuda()
    y = torch.zeros(B, H * W, Dg, device="cuda", dtype=dt)
    ops.selective_scan_fwd_merge(c(u).to(dt), c(delta).to(dt), c(A), c(Bm), c(Cm), c(D), c(bias), True, y, H, W)
    assert rel(y.reshape(B, H, W, Dg), y_ref) < TOL[dt]
    xz = q(torch.randn(B, H, W, 2 * Dg, generator=g), dt)
    gamma, beta, local = torch.randn(Dg, generator=g), torch.randn(Dg, generator=g), torch.randn(B, Dg, generator=g)
    ref = F.layer_norm(y.float().cpu().reshape(B, H, W, Dg), (Dg,), gamma, beta, eps=1e-5) * xz[..., Dg:] + local[:, None, None, :]
    out = torch.empty(B, H * W, Dg, device="cuda", dtype=dt)
    ops.ln_gate(y, xz.reshape(B, H * W, 2 * Dg).to("cuda", dt), 2 * Dg, Dg, gamma.cuda(), beta.cuda(), local.cuda(), out, B, H * W, Dg)
    assert rel(out.reshape(B, H, W, Dg), ref) < TOL[dt]


# ---------------------------------------------------------------------------------------------------------
# Secondary path kernels (lucidrains Unet, src/denoising_diffusion_pytorch.py)
@pytest.mark.parametrize("dt", DTYPES)
def test_gn_scale_shift_silu(ops, dt):
    """Block.forward with scale_shift (:183-199) + the ResnetBlock skip (:225)."""
    B, C, H, W = 2, 128, 12, 20
    P = H * W
    g = torch.Generator().manual_seed(7)
    y = q(torch.randn(B, C, H, W, generator=g) * 2 + 0.3, dt)
    skip = q(torch.randn(B, C, H, W, generator=g), dt)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ss = torch.randn(B, 2 * C, generator=g)                       # mlp(time_emb).chunk(2): scale | shift
    scale, shift = ss[:, :C], ss[:, C:]
    ref = F.silu(F.group_norm(y, 8, gamma, beta, eps=1e-5) * (scale[:, :, None, None] + 1) + shift[:, :, None, None]) + skip
    sums = torch.zeros(B, 8, 2, device="cuda")
    yd = nhwc(y, dt)
    ops.gn_stats(yd, sums, B, P, C, 8)
    out = torch.empty_like(yd)
    from founddiff_b200.engine import _view_ptr
    ssd = ss.cuda()
    ops.gn_scale_shift_silu(yd, sums, gamma.cuda(), beta.cuda(), _view_ptr(ssd[:, :C]), _view_ptr(ssd[:, C:]), 2 * C, nhwc(skip, dt),
                            out, B, P, C, 8)
    assert rel(nchw(out, H, W), ref) < TOL[dt]


@pytest.mark.parametrize("cfg", [(2, 4, 1024), (1, 4, 4096), (2, 2, 200), (1, 1, 64), (3, 2, 128), (1, 3, 392)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("impl", ["tc", "mma"])
def test_flash_attention_d32(ops, cfg, dt, impl):
    """Attention.forward between to_qkv and to_out (:266-277): softmax(q^T k * 32^-0.5) v per head — the tcgen05 / tensor-memory
    kernel (128-query CTAs; ragged last key and query tiles at N = 200, 392, 64) and the mma.sync kernel it replaced."""
    B, heads, N = cfg
    HC = heads * 32
    g = torch.Generator().manual_seed(N + heads)
    qkv = q(torch.randn(B, N, 3 * HC, generator=g), dt)
    qq, kk, vv = [t.reshape(B, N, heads, 32).permute(0, 2, 1, 3) for t in qkv.chunk(3, dim=-1)]      # (B, h, N, d)
    attn = ((qq * 32 ** -0.5) @ kk.transpose(-2, -1)).softmax(dim=-1)
    ref = (attn @ vv).permute(0, 2, 1, 3).reshape(B, N, HC)
    out = torch.zeros(B, N, HC, device="cuda", dtype=dt)
    ops.flash_attn_d32(qkv.to("cuda", dt), out, B, N, heads, 32 ** -0.5, impl=impl)
    # P is rounded to the storage type before the P.V tensor-core product
    r = rel(out, ref)
    print(f"flash_attn_d32 {impl} {cfg} {dt}: rel-L2 {r:.3e}")
    assert torch.isfinite(out.float()).all()
    assert r < (1e-2 if dt == torch.bfloat16 else 2e-3), r


@pytest.mark.parametrize("cfg", [(2, 64, 32, 32), (1, 128, 64, 48), (2, 64, 16, 16)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_linear_attention(ops, cfg, dt):
    """LinearAttention.forward after to_qkv (:238-255): softmax_d(q)*scale, softmax_n(k), v/N, context, out, to_out conv + LayerNorm."""
    B, dim, H, W = cfg
    heads, HC, N = 4, 128, H * W
    g = torch.Generator().manual_seed(dim + H)
    qkv = q(torch.randn(B, 3 * HC, H, W, generator=g), dt)
    wout = torch.randn(dim, HC, generator=g) / math.sqrt(HC)
    bias, gam = torch.randn(dim, generator=g), torch.randn(dim, generator=g)
    qq, kk, vv = [t.reshape(B, heads, 32, N) for t in qkv.chunk(3, dim=1)]
    qs = qq.softmax(dim=-2) * 32 ** -0.5
    ks = kk.softmax(dim=-1)
    ctx = torch.einsum("bhdn,bhen->bhde", ks, vv / N)
    o = torch.einsum("bhde,bhdn->bhen", ctx, qs).reshape(B, HC, H, W)
    y = F.conv2d(o, wout.reshape(dim, HC, 1, 1), bias)
    var, mean = torch.var(y, dim=1, unbiased=False, keepdim=True), torch.mean(y, dim=1, keepdim=True)
    ref = (y - mean) * (var + 1e-5).rsqrt() * gam.reshape(1, dim, 1, 1)
    out = torch.empty(B, N, dim, device="cuda", dtype=dt)
    ops.linear_attention(nhwc(qkv, dt), wout.cuda(), bias.cuda(), gam.cuda(), out, B, H, W, heads, dim)
    assert rel(nchw(out, H, W), ref) < (2e-2 if dt == torch.bfloat16 else 4e-3)


def test_ddpm_update_matches_gaussian_diffusion_formulas(ops):
    """p_sample (:588-595) and one DDIM step (:629-642) of the epsilon-prediction GaussianDiffusion, cosine schedule."""
    T = 1000
    steps = T + 1
    x = torch.linspace(0, T, steps, dtype=torch.float64)
    ac = torch.cos(((x / T) + 0.008) / 1.008 * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    alphas = 1. - betas
    abar = torch.cumprod(alphas, dim=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.)
    pv = betas * (1. - abar_prev) / (1. - abar)
    c1 = betas * torch.sqrt(abar_prev) / (1. - abar)
    c2 = (1. - abar_prev) * torch.sqrt(alphas) / (1. - abar)
    g = torch.Generator().manual_seed(3)
    xt, eps, nz = (torch.randn(2, 4096, generator=g) for _ in range(3))
    t, tn = 700, 500
    sr, srm1 = float(torch.sqrt(1. / abar[t])), float(torch.sqrt(1. / abar[t] - 1))
    x0 = (sr * xt - srm1 * eps).clamp(-1, 1)
    ref_anc = float(c1[t]) * x0 + float(c2[t]) * xt + float((0.5 * torch.log(pv[t].clamp(min=1e-20))).exp()) * nz
    ref_ddim = x0 * float(abar[tn].sqrt()) + float((1 - abar[tn]).sqrt()) * eps       # eta = 0
    out, xs = torch.empty(2, 4096, device="cuda"), torch.empty(2, 4096, device="cuda")
    coef = torch.tensor([sr, srm1, float(c1[t]), float(c2[t]), 0., float((0.5 * torch.log(pv[t].clamp(min=1e-20))).exp()), 1., 0.])
    ops.ddpm_update(xt.cuda(), eps.cuda(), nz.cuda(), coef.cuda(), out, xs)
    assert rel(out, ref_anc) < 1e-6 and rel(xs, x0) < 1e-6
    coef = torch.tensor([sr, srm1, float(abar[tn].sqrt()), 0., float((1 - abar[tn]).sqrt()), 0., 1., 0.])
    ops.ddpm_update(xt.cuda(), eps.cuda(), None, coef.cuda(), out)
    assert rel(out, ref_ddim) < 1e-6


@pytest.mark.parametrize("cfg", [(32, 64, 128, 4, 4), (16, 48, 64, 8, 8), (64, 64, 32, 4, 4), (24, 40, 96, 8, 8)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_x_proj_and_scan_with_fused_dt_proj(ops, cfg, dt):
    """x_proj alone (tensor cores, fp32 x_dbl) + scan with dt_proj and EfficientMerge fused == the oracle's x_proj ->
    dt_proj -> selective scan -> EfficientMerge (src/emamba2.py:334-367)."""
    H, W, D, N, R = cfg
    B, L = 2, (H // 2) * (W // 2)
    g = torch.Generator().manual_seed(H * W + D)
    xs = q(torch.randn(B, 4, D, L, generator=g), dt)
    Wx = torch.randn(4, R + 2 * N, D, generator=g) / math.sqrt(D)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    A = -torch.exp(torch.randn(4 * D, N, generator=g) * 0.3)
    Dp, bias = torch.randn(4 * D, generator=g), torch.randn(4 * D, generator=g) * 0.5
    # reference (fp32 on the 16-bit-rounded xs / x_proj weights, as the kernel sees them)
    x_dbl_ref = torch.einsum("bkdl,kcd->bkcl", xs, q(Wx, dt))
    dts_r, Bs_r, Cs_r = torch.split(x_dbl_ref, [R, N, N], dim=2)
    delta = torch.einsum("bkrl,kdr->bkdl", dts_r, Wdt).reshape(B, 4 * D, L)
    y_ref = scan_cpu.selective_scan_fwd(xs.reshape(B, 4 * D, L).contiguous(), delta.contiguous(), A, Bs_r.contiguous(), Cs_r.contiguous(),
                                        Dp, bias, True)
    y_ref = O.efficient_merge(y_ref.view(B, 4, D, L), H, W).permute(0, 2, 3, 1)          # (B, H, W, D)
    xw16, _dw16, _Rp = ops.pack_xdt_weights(Wx.cuda(), Wdt.cuda(), dt)
    x_dbl = torch.zeros(B, 4, R + 2 * N, L, device="cuda")
    ops.x_proj_tc(xs.to("cuda", dt), xw16, x_dbl, B, D, L, R, N)
    assert rel(x_dbl, x_dbl_ref) < 2e-5
    y = torch.empty(B, H * W, D, device="cuda", dtype=dt)
    ops.selective_scan_fwd_merge_xdbl(xs.reshape(B, 4 * D, L).to("cuda", dt), x_dbl, Wdt.reshape(4 * D, R).cuda(), A.cuda(), Dp.cuda(),
                                      bias.cuda(), True, y, H, W)
    assert rel(y.reshape(B, H, W, D), y_ref) < TOL[dt]


@pytest.mark.parametrize("cfg", [(16, 16, 32, 16, 16), (16, 32, 64, 32, 32), (8, 24, 96, 16, 16), (32, 16, 128, 8, 8)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_channel_per_lane_scan_with_time_major_bc(ops, cfg, dt):
    """xdt_proj with time-major B/C + the channel-per-lane scan with EfficientMerge fused == oracle x_proj -> dt_proj ->
    selective scan -> EfficientMerge (src/emamba2.py:334-367); also == the warp-shuffle kernel on the same inputs."""
    H, W, D, N, R = cfg
    B, L = 2, (H // 2) * (W // 2)
    g = torch.Generator().manual_seed(H * W + D + N)
    xs = q(torch.randn(B, 4, D, L, generator=g), dt)
    Wx = torch.randn(4, R + 2 * N, D, generator=g) / math.sqrt(D)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    A = -torch.exp(torch.randn(4 * D, N, generator=g) * 0.3)
    Dp, bias = torch.randn(4 * D, generator=g), torch.randn(4 * D, generator=g) * 0.5
    xw16, dw16, Rp = ops.pack_xdt_weights(Wx.cuda(), Wdt.cuda(), dt)
    xs_d = xs.to("cuda", dt)
    dts = torch.empty(B, 4, D, L, device="cuda", dtype=dt)
    Bs, Cs = torch.empty(B, 4, N, L, device="cuda"), torch.empty(B, 4, N, L, device="cuda")
    Bt, Ct = torch.empty(B, 4, L, N, device="cuda"), torch.empty(B, 4, L, N, device="cuda")
    dts_t = torch.empty_like(dts)
    ops.xdt_proj_tc(xs_d, xw16, dw16, Rp, dts, Bs, Cs, B, D, L, R, N)
    ops.xdt_proj_tc(xs_d, xw16, dw16, Rp, dts_t, Bt, Ct, B, D, L, R, N, time_major=True)
    assert torch.equal(Bt.permute(0, 1, 3, 2), Bs) and torch.equal(Ct.permute(0, 1, 3, 2), Cs) and torch.equal(dts, dts_t)
    # bias + softplus finished by the producer: compare with the fp32 formula on the un-rounded projection
    dts_sp = torch.empty_like(dts)
    ops.xdt_proj_tc(xs_d, xw16, dw16, Rp, dts_sp, Bs, Cs, B, D, L, R, N, dt_bias=bias.cuda(), delta_softplus=True)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, q(Wx, dt))
    dt_ref = torch.einsum("bkrl,kdr->bkdl", q(x_dbl[:, :, :R], dt), q(Wdt, dt)) + bias.view(1, 4, D, 1)
    assert rel(dts_sp, F.softplus(dt_ref)) < TOL[dt]
    # oracle scan on exactly the tensors the kernel reads
    y_ref = scan_cpu.selective_scan_fwd(xs.reshape(B, 4 * D, L), dts.float().cpu().reshape(B, 4 * D, L), A, Bs.cpu(), Cs.cpu(), Dp, bias, True)
    y_ref = O.efficient_merge(y_ref.view(B, 4, D, L), H, W).permute(0, 2, 3, 1)
    y = torch.empty(B, H * W, D, device="cuda", dtype=dt)
    ops.selective_scan_fwd_merge_cl(xs_d.view(B, 4 * D, L), dts.view(B, 4 * D, L), A.cuda(), Bt, Ct, Dp.cuda(), bias.cuda(), True, y, H, W)
    assert rel(y.reshape(B, H, W, D), y_ref) < TOL[dt]
    y2 = torch.empty_like(y)
    ops.selective_scan_fwd_merge(xs_d.view(B, 4 * D, L), dts.view(B, 4 * D, L), A.cuda(), Bs, Cs, Dp.cuda(), bias.cuda(), True, y2, H, W)
    assert rel(y, y2) < TOL[dt]


@pytest.mark.parametrize("hw", [(16, 32), (64, 48), (8, 16)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_init_conv7x7_tensor_core(ops, hw, dt):
    """init_conv on the tensor cores (fp32 images split into fp16 hi + lo, fp16 weights) == F.conv2d in fp32 up to the
    weight rounding and the output storage type (src/DADiff.py:558, 700)."""
    H, W = hw
    B = 2
    g = torch.Generator().manual_seed(H + W)
    x_t, x_in = torch.randn(B, H * W, generator=g) * 1.5, torch.rand(B, H * W, generator=g) * 2 - 1
    w = torch.randn(64, 2, 7, 7, generator=g) / math.sqrt(98)
    bias = torch.randn(64, generator=g)
    ref = F.conv2d(torch.stack([x_t.view(B, H, W), x_in.view(B, H, W)], dim=1), w.half().float(), bias, padding=3)
    out = torch.empty(B, H * W, 64, device="cuda", dtype=dt)
    ops.init_conv7x7_tc(x_t.cuda(), x_in.cuda(), ops.pack_init_conv_weights(w.cuda()), bias.cuda(), out, B, H, W)
    got = out.float().cpu().view(B, H, W, 64).permute(0, 3, 1, 2)
    assert rel(got, ref) < (3e-3 if dt == torch.bfloat16 else 4e-4), rel(got, ref)
    # the hi/lo split keeps the inputs exact: against fp32 weights only the fp16 weight rounding (2^-12) remains
    ref32 = F.conv2d(torch.stack([x_t.view(B, H, W), x_in.view(B, H, W)], dim=1), w, bias, padding=3)
    assert rel(got, ref32) < (3e-3 if dt == torch.bfloat16 else 6e-4)


@pytest.mark.parametrize("types", [(torch.bfloat16, torch.float16), (torch.float16, torch.bfloat16)])
def test_conv2d_mixed_operand_and_output_types(ops, types):
    """fd_conv_params.ab_dtype_p1: operands (src / weight) in one 16-bit type, output + addend in the other, plus relu_out —
    tcgen05 path against the CUDA-core path and against F.conv2d (the bf16 sampling mode's fp16 residual stream and the
    DA-CLIP blocks use exactly this)."""
    ab, od = types
    B, C, H, W, cout = 2, 128, 16, 32, 64
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, H * W, C, generator=g).to("cuda", ab)
    w = (torch.randn(cout, 3, 3, C, generator=g) / math.sqrt(9 * C)).to("cuda", ab)
    bias = torch.randn(cout, generator=g).cuda()
    add = torch.randn(B, H * W, cout, generator=g).to("cuda", od)
    ref = F.conv2d(x.float().cpu().view(B, H, W, C).permute(0, 3, 1, 2), w.float().cpu().permute(0, 3, 1, 2), bias.cpu(), padding=1)
    ref = F.relu(ref + add.float().cpu().view(B, H, W, cout).permute(0, 3, 1, 2))
    res = []
    for tc in (False, True):
        out = torch.zeros(B, H * W, cout, device="cuda", dtype=od)
        conv = ops.Conv(x, w, out, B=B, Hin=H, Win=W, KH=3, KW=3, pad=1, bias=bias, addend=add, relu_out=True, prefer_tc=tc)
        assert conv.uses_tc == tc
        conv.run()
        res.append(out.float().cpu().view(B, H, W, cout).permute(0, 3, 1, 2))
    tol = 1e-2 if od == torch.bfloat16 else 2e-3
    assert rel(res[0], ref) < tol and rel(res[1], ref) < tol and rel(res[1], res[0]) < tol


def test_ln_modulate_separate_in_out_types(ops):
    B, P, C = 2, 96, 128
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, P, C, generator=g) * 3 + 1
    mods = torch.randn(B, 2 * C, generator=g)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    xh = x.to(torch.float16)
    ref = F.layer_norm(xh.float(), (C,), gamma, beta, eps=1e-5) * (1 + mods[:, None, C:]) + mods[:, None, :C]
    out = torch.empty(B, P, C, device="cuda", dtype=torch.bfloat16)
    from founddiff_b200.engine import _view_ptr
    md = mods.cuda()
    ops.ln_modulate(xh.cuda(), out, gamma.cuda(), beta.cuda(), _view_ptr(md[:, :C]), _view_ptr(md[:, C:]), 2 * C, B, P, C, 1e-5)
    assert rel(out, ref) < TOL[torch.bfloat16]


@pytest.mark.parametrize("dt", DTYPES)
def test_avgpool2x2_nhwc(ops, dt):
    B, H, W, C = 2, 12, 20, 64
    x = q(torch.randn(B, H, W, C, generator=torch.Generator().manual_seed(4)), dt)
    out = torch.empty(B, (H // 2) * (W // 2), C, device="cuda", dtype=dt)
    ops.avgpool2x2_nhwc(x.reshape(B, H * W, C).to("cuda", dt), out, B, H, W, C)
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).reshape(B, -1, C)
    assert rel(out, ref) < TOL[dt]


@pytest.mark.parametrize("hw", [(64, 96), (33, 50), (512, 512)])
def test_device_metrics_match_reference_formulas(ops, hw):
    """fd_slice_metrics / founddiff_b200.metrics == compute_psnr / compute_ssim / compute_rmse (src/util.py:188-236)."""
    from founddiff_b200 import metrics
    from oracle import metrics_oracle as MO
    H, W = hw
    g = torch.Generator().manual_seed(H + W)
    y = torch.rand(3, 1, H, W, generator=g)
    y = F.avg_pool2d(F.pad(y, (2, 2, 2, 2), mode="reflect"), 5, stride=1)          # low-frequency content, like CT
    p = (y + 0.05 * torch.randn(3, 1, H, W, generator=g)).clamp(0, 1)
    yd, pd = y.cuda(), p.cuda()
    assert abs(float(metrics.compute_psnr(pd, yd)) - float(MO.compute_psnr(p, y))) < 1e-3
    assert abs(float(metrics.compute_rmse(pd, yd)) - float(MO.compute_rmse(p, y))) < 1e-6
    assert abs(float(metrics.compute_ssim(pd, yd)) - float(MO.compute_ssim(p, y))) < 2e-5
    ps, ss, rs = metrics.slice_metrics(pd, yd)
    for i in range(3):
        assert abs(float(ps[i]) - float(MO.compute_psnr(p[i:i + 1], y[i:i + 1]))) < 1e-3
        assert abs(float(ss[i]) - float(MO.compute_ssim(p[i:i + 1], y[i:i + 1]))) < 2e-5
        assert abs(float(rs[i]) - float(MO.compute_rmse(p[i:i + 1], y[i:i + 1]))) < 1e-6


def test_device_metrics_vs_reference_util_fixture(ops):
    """Device PSNR / SSIM / RMSE against tests/golden/metrics.npz (the reference's src/util.py functions, kornia's two
    primitives supplied by OpenCV; oracle/gen_golden_metrics.py)."""
    from founddiff_b200 import metrics
    g = load_golden("metrics.npz")
    for tag in "abc":
        pred, tgt = g[f"{tag}.pred"].cuda(), g[f"{tag}.target"].cuda()
        ps, ss, rs = metrics.slice_metrics(pred, tgt)
        for i in range(pred.shape[0]):
            assert abs(float(ss[i]) - float(g[f"{tag}.{i}.ssim"])) < 2e-5, tag
            assert abs(float(ps[i]) - float(g[f"{tag}.{i}.psnr"])) < 1e-3, tag
            assert abs(float(rs[i]) - float(g[f"{tag}.{i}.rmse"])) < 1e-6, tag


@pytest.mark.parametrize("cfg", [(16, 32, 64, 256, True, 128), (16, 16, 128, 384, False, None), (8, 32, 256, 1024, True, 512), (24, 48, 64, 192, False, None),
                                 (16, 16, 512, 1024, True, 512)])
@pytest.mark.parametrize("types", [(torch.float16, torch.bfloat16), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16)])
def test_layernorm_folded_into_1x1_gemm(ops, cfg, types):
    """LayerNorm (+ affine) + adaLN modulate + 1x1 projection (src/DADiff.py:486-487 -> src/emamba2.py:717 in_proj with SiLU on the
    z half, src/DADiff.py:258 qkv) as ONE tcgen05 GEMM on the raw residual stream: fd_ln_fold makes the per-sample weights, the
    kernel's statistics warps take mean / rstd from the staged operand tile.  Against torch: F.layer_norm -> modulate -> F.linear."""
    H, W, C, Cout, affine, silu_from = cfg
    in_dt, out_dt = types
    B, P = 3, H * W
    g = torch.Generator().manual_seed(C + Cout)
    x = q(torch.randn(B, P, C, generator=g) * 1.5 + 0.7, in_dt)                       # non-zero mean: the mean term must cancel
    Wt = torch.randn(Cout, C, generator=g) / math.sqrt(C)
    gamma = 1 + 0.2 * torch.randn(C, generator=g) if affine else None
    beta = 0.2 * torch.randn(C, generator=g) if affine else None
    mods = torch.randn(B, 2 * C + 5, generator=g) * 0.3                                # strided modulation rows, as in the engine
    shift, scale = mods[:, :C], mods[:, C:2 * C]
    eps = 1e-5 if affine else 1e-6
    xn = F.layer_norm(x, (C,), gamma, beta, eps) * (1 + scale[:, None]) + shift[:, None]
    ref = torch.einsum("bpc,oc->bpo", xn, Wt)
    if silu_from is not None:
        ref = torch.cat([ref[..., :silu_from], F.silu(ref[..., silu_from:])], dim=-1)

    class V:                                    # strided fp32 view handed over as (pointer, row stride), like engine._view_ptr
        def __init__(self, t):
            self.t, self.dtype, self.is_cuda = t, torch.float32, True

        def is_contiguous(self):
            return True

        def data_ptr(self):
            return self.t.data_ptr()
    md = mods.cuda()
    wf = torch.empty(B, Cout, C, device="cuda", dtype=in_dt)
    v = torch.empty(B, Cout, device="cuda")
    ops.ln_fold(Wt.cuda(), gamma.cuda() if affine else None, beta.cuda() if affine else None, V(md[:, :C]), V(md[:, C:2 * C]),
                mods.shape[1], wf, v, B, Cout, C)
    assert float(wf.float().sum(dim=-1).abs().max()) < 0.05         # rows sum to zero up to the 16-bit rounding of the entries
    out = torch.full((B, P, Cout), float("nan"), device="cuda", dtype=out_dt)
    conv = ops.Conv(x.to("cuda", in_dt), wf, out, B=B, Hin=H, Win=W, per_batch_weight=True, ln_v=v, ln_eps=eps,
                    silu_from=silu_from)
    assert conv.uses_tc
    conv.run()
    conv.run()                                  # second run: barrier phases / buffers are reusable
    r = rel(out, ref)
    print(f"ln-folded 1x1 {cfg} {types}: rel-L2 {r:.3e}")
    assert torch.isfinite(out.float()).all() and r < TOL[out_dt] * (1.5 if in_dt == torch.bfloat16 else 1.0), r
    # the same GEMM with the row statistics from fd_row_rstd (no statistics warps in the kernel): the engine's form
    rstd = torch.full((B, P), float("nan"), device="cuda")
    ops.row_rstd(x.to("cuda", in_dt), rstd, B * P, C, eps)
    assert rel(rstd, 1.0 / torch.sqrt(x.var(dim=-1, unbiased=False) + eps)) < 1e-5
    out2 = torch.full((B, P, Cout), float("nan"), device="cuda", dtype=out_dt)
    conv2 = ops.Conv(x.to("cuda", in_dt), wf, out2, B=B, Hin=H, Win=W, per_batch_weight=True, ln_v=v, ln_eps=eps, silu_from=silu_from,
                     ln_rstd=rstd)
    assert conv2.uses_tc
    conv2.run()
    conv2.run()
    r2 = rel(out2, ref)
    assert torch.isfinite(out2.float()).all() and r2 < TOL[out_dt] * (1.5 if in_dt == torch.bfloat16 else 1.0), r2
    assert rel(out2, out.float().cpu()) < 2e-3          # both forms: same GEMM, statistics one-pass (in kernel) vs two-pass


@pytest.mark.parametrize("dts", [(torch.bfloat16, torch.float16), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16)])
@pytest.mark.parametrize("BP", [(2, 48), (3, 4096), (1, 16)])
def test_ln_gate_out_proj_fused_tail(dts, BP):
    """fd_ln_gate_out_proj (out_norm -> * silu(z) + local -> out_proj -> gated residual, src/emamba2.py:365, 747-748 and
    src/DADiff.py:486) against the same chain in fp32 torch, and against the two-launch form it replaces (fd_ln_gate + 1x1
    conv with gate / addend epilogue): same roundings (the gated row is rounded to the operand type in both), so they agree
    to accumulation order."""
    from founddiff_b200 import ops
    dti, dto = dts
    B, P = BP
    D, C = 128, 64
    g = torch.Generator().manual_seed(B * P)
    y = (torch.randn(B, P, D, generator=g) * 3 + 0.5).to(dti)
    xz = torch.randn(B, P, 4 * C, generator=g).to(dti)
    gamma, beta = torch.randn(D, generator=g) * 0.3 + 1, torch.randn(D, generator=g) * 0.1
    local = torch.randn(B, D, generator=g) * 0.2
    w = (torch.randn(C, D, generator=g) / D ** 0.5).to(dti)
    mods = torch.randn(B, 6 * C, generator=g) * 0.5
    addend = torch.randn(B, P, C, generator=g).to(dto)
    assert ops.ln_gate_out_proj_supported(P, D, C, 4 * C, 2 * C, dti, dto)
    yf, zf = y.float(), xz.float()[..., 2 * C:]
    row = (F.layer_norm(yf, (D,), gamma, beta, 1e-5) * zf + local[:, None, :]).to(dti).float()
    gate = mods[:, C:2 * C]
    ref = addend.float() + gate[:, None, :] * (row @ w.float().t())
    out = torch.full((B, P, C), float("nan"), device="cuda", dtype=dto)
    mods_d = mods.cuda()

    class View:                          # a column block of the modulation table, as the engine passes it (pointer + pitch)
        dtype, is_cuda = torch.float32, True

        def is_contiguous(self):
            return True

        def data_ptr(self):
            return mods_d.data_ptr() + C * 4
    ops.ln_gate_out_proj(y.cuda(), xz.cuda(), 4 * C, 2 * C, gamma.cuda(), beta.cuda(), local.cuda(), w.cuda(), View(), 6 * C,
                         addend.cuda(), out, B, P, D, C)
    assert torch.isfinite(out.float()).all()
    tol = 6e-3 if dto == torch.bfloat16 else 1.5e-3
    assert rel(out, ref) < tol, rel(out, ref)
    # two-launch form
    gbuf = torch.empty(B, P, D, device="cuda", dtype=dti)
    ops.ln_gate(y.cuda(), xz.cuda(), 4 * C, 2 * C, gamma.cuda(), beta.cuda(), local.cuda(), gbuf, B, P, D)
    two = addend.float() + gate[:, None, :] * (gbuf.float().cpu() @ w.float().t())
    assert rel(out, two) < tol
    # in place (the engine's trunk may alias: out == addend)
    io = addend.cuda().clone()
    ops.ln_gate_out_proj(y.cuda(), xz.cuda(), 4 * C, 2 * C, gamma.cuda(), beta.cuda(), local.cuda(), w.cuda(), View(), 6 * C, io, io, B, P, D, C)
    assert torch.equal(io, out)
