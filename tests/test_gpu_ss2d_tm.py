"""Time-major SS2D core (founddiff_b200/csrc/fd_ss2d_tm.cu) against the oracle, kernel by kernel and end to end:
depthwise conv + SiLU + EfficientScan (src/emamba2.py:480-488, 186-213), x_proj / dt_proj (:335-340), the segmented
channel-per-lane selective scan with its exact carry pass (:124-157) and EfficientMerge (:238-262)."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import founddiff_oracle as O
from oracle import scan_cpu

pytestmark = pytest.mark.gpu
TOL = {torch.bfloat16: 1e-2, torch.float16: 3e-3}


def rel(a, b):
    return O.rel_l2(a.detach().float().cpu(), b.detach().float().cpu())


def q(x, dt):
    return x.to(dt).float()


@pytest.fixture(scope="module")
def ops():
    from founddiff_b200 import ops as _ops
    return _ops


def scan_order(x):
    """(B, D, H, W) -> (B, 4, D, L) in the reference's EfficientScan order (src/emamba2.py:207-210)."""
    B, D, H, W = x.shape
    return torch.stack([x[:, :, ::2, ::2].reshape(B, D, -1), x[:, :, 1::2, ::2].transpose(2, 3).reshape(B, D, -1),
                        x[:, :, ::2, 1::2].reshape(B, D, -1), x[:, :, 1::2, 1::2].transpose(2, 3).reshape(B, D, -1)], dim=1)


@pytest.mark.parametrize("hw", [(16, 24), (32, 32), (8, 12), (64, 32), (6, 10), (36, 70), (2, 2)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_dwconv_silu_time_major(ops, hw, dt):
    H, W = hw
    B, D = 2, 128
    g = torch.Generator().manual_seed(H * W)
    xz = q(torch.randn(B, H, W, 2 * D, generator=g), dt)                 # [x | z] rows; only the x half is convolved
    w, bias = torch.randn(D, 1, 3, 3, generator=g) * 0.3, torch.randn(D, generator=g) * 0.1
    ref = F.silu(F.conv2d(xz[..., :D].permute(0, 3, 1, 2), w, bias, padding=1, groups=D))
    ref_tm = scan_order(ref).permute(0, 1, 3, 2)                           # (B, 4, L, D)
    assert torch.equal(scan_order(ref), O.efficient_scan(ref)) if hasattr(O, "efficient_scan") else True
    out = torch.full((B, 4, (H // 2) * (W // 2), D), float("nan"), device="cuda", dtype=dt)
    ops.dwconv3x3_silu_tm(xz.to("cuda", dt), 2 * D, w.reshape(D, 9).t().contiguous().cuda(), bias.cuda(), out, B, H, W, D)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref_tm) < TOL[dt]


@pytest.mark.parametrize("cfg", [(128, 4, 4, 100, True), (128, 4, 8, 384, True), (256, 8, 8, 130, True), (256, 8, 16, 64, True),
                                 (512, 16, 16, 40, False), (1024, 32, 32, 24, False), (512, 16, 32, 1000, False), (128, 4, 4, 1, True)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_x_proj_time_major(ops, cfg, dt):
    D, R, N, L, fuse = cfg
    B = 2
    g = torch.Generator().manual_seed(D + L)
    xs_tm = q(torch.randn(B, 4, L, D, generator=g), dt)
    Wx = torch.randn(4, R + 2 * N, D, generator=g) / math.sqrt(D)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    bias = torch.randn(4 * D, generator=g) * 0.5
    x_dbl = torch.einsum("bkld,kcd->bklc", xs_tm, q(Wx, dt))
    xw16, dw16, Rp = ops.pack_xdt_weights(Wx.cuda(), Wdt.cuda(), dt)
    if fuse:
        out = torch.full((B, 4, L, R + 2 * N), float("nan"), device="cuda")
        ops.x_proj_tm(xs_tm.to("cuda", dt), xw16, out, None, None, None, B, D, L, R, N, Rp, True)
        assert rel(out, x_dbl) < 2e-5
    else:
        out = torch.full((B, 4, L, 2 * N), float("nan"), device="cuda")
        dts = torch.full((B, 4, L, D), float("nan"), device="cuda", dtype=dt)
        ops.x_proj_tm(xs_tm.to("cuda", dt), xw16, out, dw16, dts, bias.cuda(), B, D, L, R, N, Rp, False)
        assert rel(out, x_dbl[..., R:]) < 2e-5
        dt_ref = torch.einsum("bklr,kdr->bkld", q(x_dbl[..., :R], dt), q(Wdt, dt)) + bias.view(1, 4, 1, D)
        assert rel(dts, F.softplus(dt_ref)) < TOL[dt]


def _scan_case(ops, H, W, D, N, R, fuse, dt, segments, slow=False, B=2, seed=0):
    L = (H // 2) * (W // 2)
    g = torch.Generator().manual_seed(H * W + D + N + seed)
    u_tm = q(torch.randn(B, 4, L, D, generator=g), dt)
    A = -torch.exp(torch.randn(4 * D, N, generator=g) * 0.3)
    if slow:                                   # channels that remember far longer than a segment: the carry pass must walk it all
        A[::7] *= 1e-3
    Dp, bias = torch.randn(4 * D, generator=g), torch.randn(4 * D, generator=g) * 0.5 - (3.0 if slow else 0.0)
    Bm, Cm = torch.randn(B, 4, L, N, generator=g), torch.randn(B, 4, L, N, generator=g)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    S = segments if segments > 0 else ops.scan_tm_segments(B, D, H, W)
    carry = torch.full((B * 4 * max(S, 1) * 2 * N * D,), float("nan"), device="cuda")
    y = torch.full((B, H * W, D), float("nan"), device="cuda", dtype=dt)
    if fuse:
        dtin = torch.randn(B, 4, L, R, generator=g)
        xdbl = torch.cat([dtin, Bm, Cm], dim=-1).contiguous()
        delta_raw = torch.einsum("bklr,kdr->bkdl", dtin, Wdt).reshape(B, 4 * D, L)            # + bias, softplus inside the oracle
        y_ref = scan_cpu.selective_scan_fwd(u_tm.permute(0, 1, 3, 2).reshape(B, 4 * D, L).contiguous(), delta_raw.contiguous(), A,
                                            Bm.permute(0, 1, 3, 2).contiguous(), Cm.permute(0, 1, 3, 2).contiguous(), Dp, bias, True)
        ops.selective_scan_tm(u_tm.to("cuda", dt), None, xdbl.cuda(), A.cuda(), Wdt.reshape(4 * D, R).contiguous().cuda(), bias.cuda(),
                              Dp.cuda(), carry, y, B, D, H, W, N, R, segments)
    else:
        delta = q(F.softplus(torch.randn(B, 4, L, D, generator=g) + bias.view(1, 4, 1, D)), dt)   # as x_proj_tm stores it
        xdbl = torch.cat([Bm, Cm], dim=-1).contiguous()
        y_ref = scan_cpu.selective_scan_fwd(u_tm.permute(0, 1, 3, 2).reshape(B, 4 * D, L).contiguous(),
                                            delta.permute(0, 1, 3, 2).reshape(B, 4 * D, L).contiguous(), A,
                                            Bm.permute(0, 1, 3, 2).contiguous(), Cm.permute(0, 1, 3, 2).contiguous(), Dp, None, False)
        ops.selective_scan_tm(u_tm.to("cuda", dt), delta.to("cuda", dt), xdbl.cuda(), A.cuda(), None, None, Dp.cuda(), carry, y,
                              B, D, H, W, N, 0, segments)
    y_ref = O.efficient_merge(y_ref.view(B, 4, D, L), H, W).permute(0, 2, 3, 1)
    assert torch.isfinite(y.float()).all()
    return rel(y.reshape(B, H, W, D), y_ref)


@pytest.mark.parametrize("cfg", [(16, 24, 128, 4, 4, True), (64, 64, 128, 8, 4, True), (32, 48, 256, 8, 8, True), (24, 40, 256, 16, 8, True),
                                 (16, 16, 512, 16, 16, False), (8, 12, 1024, 32, 32, False), (6, 10, 128, 4, 4, True),
                                 (32, 32, 128, 8, 8, False), (2, 2, 128, 4, 4, True)])
@pytest.mark.parametrize("segments", [1, 3, 0, -8, -4, -1008, -1004])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_scan_time_major_vs_oracle(ops, cfg, segments, dt):
    H, W, D, N, R, fuse = cfg
    if segments < 0 and (not fuse or N > 8):
        pytest.skip("the time-sliced kernel exists for the fused-dt, d_state <= 8 levels")
    r = _scan_case(ops, H, W, D, N, R, fuse, dt, segments)
    assert r < TOL[dt], r


@pytest.mark.parametrize("slow", [False, True])
@pytest.mark.parametrize("segments", [1, 2, 8, 16, -8, -4, -1008, 0])
def test_scan_time_major_segments_are_exact(ops, segments, slow):
    """Long rows (L = 16384) cut into 1 / 2 / 8 / 16 segments give the same answer as the sequential recurrence — including
    channels whose memory is far longer than a segment (A scaled by 1e-3, small delta), where the carry pass cannot stop early
    and the state must be handed across several segment boundaries."""
    r = _scan_case(ops, 256, 256, 128, 4, 4, True, torch.float16, segments, slow=slow, B=1)
    print(f"segments={segments} slow={slow}: rel-L2 {r:.3e}")
    assert r < 1.5e-3, r                        # fp16 output rounding alone is ~3e-4


@pytest.mark.parametrize("cfg", [(128, 128, 128, 4, 4, 2), (256, 256, 128, 4, 4, 1), (128, 128, 128, 8, 4, 2), (128, 128, 256, 8, 8, 2),
                                 (256, 128, 256, 8, 8, 4)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_chained_time_sliced_scan_is_bit_identical(ops, cfg, dt, monkeypatch):
    """fd_selective_scan_tm_chained (rows cut into segments that run as separate blocks, the state handed over through global
    memory) against the one-block-per-row kernel on the same inputs: the SAME bits, launch after launch on one workspace (every
    launch must leave the ticket counter and the hand-over flags zeroed), with channels whose memory spans many segments; and
    against the fp32 recurrence of the oracle.  Reference op: src/emamba2.py:124-157, merge :238-262."""
    H, W, D, N, R, B = cfg
    L = (H // 2) * (W // 2)
    nseg, ws_floats = ops.scan_tm_chain_plan(B, D, H, W, N, R)
    if nseg <= 1:                               # short rows are not chained by default: force it (the variable is read per call)
        monkeypatch.setenv("FD_SCAN_CHAIN", "4")
        nseg, ws_floats = ops.scan_tm_chain_plan(B, D, H, W, N, R)
    if nseg <= 1:
        pytest.skip(f"geometry {cfg} is not chained on this device (plan {ops.scan_tm_plan(B, D, H, W, N, R)})")
    g = torch.Generator().manual_seed(H + W + D + N)
    u_tm = q(torch.randn(B, 4, L, D, generator=g), dt).to("cuda", dt)
    A = -torch.exp(torch.randn(4 * D, N, generator=g) * 0.3)
    A[::5] *= 1e-3                              # memory far longer than a segment
    Dp, bias = torch.randn(4 * D, generator=g), torch.randn(4 * D, generator=g) * 0.5 - 1.0
    xdbl = torch.randn(B, 4, L, R + 2 * N, generator=g)
    Wdt = (torch.randn(4, D, R, generator=g) / math.sqrt(R)).reshape(4 * D, R).contiguous()
    args = (xdbl.cuda(), A.cuda(), Wdt.cuda(), bias.cuda(), Dp.cuda())
    y0 = torch.full((B, H * W, D), float("nan"), device="cuda", dtype=dt)
    ops.selective_scan_tm(u_tm, None, args[0], args[1], args[2], args[3], args[4], None, y0, B, D, H, W, N, R,
                          ops.scan_tm_plan(B, D, H, W, N, R))
    ws = torch.zeros(ws_floats, device="cuda")
    links = B * 4 * (D // 32) * (nseg - 1)
    for it in range(3):
        y1 = torch.full((B, H * W, D), float("nan"), device="cuda", dtype=dt)
        ops.selective_scan_tm_chained(u_tm, *args, ws, y1, B, D, H, W, N, R)
        torch.cuda.synchronize()
        assert torch.equal(y0.view(torch.int16), y1.view(torch.int16)), f"launch {it}: chained scan differs from the unchained kernel"
        assert int(ws[: 1 + links].view(torch.int32).abs().sum()) == 0, f"launch {it}: ticket counter / flags not left zeroed"
    delta_raw = torch.einsum("bklr,kdr->bkdl", xdbl[..., :R], Wdt.view(4, D, R)).reshape(B, 4 * D, L)
    y_ref = scan_cpu.selective_scan_fwd(u_tm.float().cpu().permute(0, 1, 3, 2).reshape(B, 4 * D, L).contiguous(), delta_raw.contiguous(), A,
                                        xdbl[..., R:R + N].permute(0, 1, 3, 2).contiguous(), xdbl[..., R + N:].permute(0, 1, 3, 2).contiguous(),
                                        Dp, bias, True)
    y_ref = O.efficient_merge(y_ref.view(B, 4, D, L), H, W).permute(0, 2, 3, 1)
    r = rel(y1.reshape(B, H, W, D), y_ref)
    print(f"chained scan {cfg} {dt}: {nseg} segments, rel-L2 vs oracle {r:.3e}")
    assert r < TOL[dt], r


@pytest.mark.parametrize("cfg", [(32, 48, 64, 4, 4), (64, 32, 64, 8, 4), (32, 32, 128, 8, 8), (16, 48, 128, 16, 8), (16, 16, 256, 16, 16),
                                 (16, 16, 512, 32, 32), (12, 20, 64, 4, 4)])
@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_time_major_pipeline_vs_oracle(ops, cfg, dt):
    """dwconv -> x_proj (-> dt_proj) -> scan -> merge chained exactly as the engine does, against the oracle's
    conv2d + SiLU -> cross_selective_scan (src/emamba2.py:295-367) on the same 16-bit-rounded input."""
    H, W, C, N, R = cfg
    B, D, L = 2, 2 * C, (H // 2) * (W // 2)
    fuse = R <= 8
    g = torch.Generator().manual_seed(H + W + C)
    xz = q(torch.randn(B, H, W, 2 * D, generator=g), dt)
    wc, bc = torch.randn(D, 1, 3, 3, generator=g) * 0.3, torch.randn(D, generator=g) * 0.1
    Wx = torch.randn(4, R + 2 * N, D, generator=g) / math.sqrt(D)
    Wdt = torch.randn(4, D, R, generator=g) / math.sqrt(R)
    A = -torch.exp(torch.randn(4 * D, N, generator=g) * 0.3)
    Dp, bias = torch.randn(4 * D, generator=g), torch.randn(4 * D, generator=g) * 0.5
    x = F.silu(F.conv2d(xz[..., :D].permute(0, 3, 1, 2), wc, bc, padding=1, groups=D))
    xs = scan_order(x)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, Wx)
    dts_r, Bs_r, Cs_r = torch.split(x_dbl, [R, N, N], dim=2)
    delta = torch.einsum("bkrl,kdr->bkdl", dts_r, Wdt).reshape(B, 4 * D, L)
    y_ref = scan_cpu.selective_scan_fwd(xs.reshape(B, 4 * D, L).contiguous(), delta.contiguous(), A, Bs_r.contiguous(), Cs_r.contiguous(),
                                        Dp, bias, True)
    y_ref = O.efficient_merge(y_ref.view(B, 4, D, L), H, W).permute(0, 2, 3, 1)
    xw16, dw16, Rp = ops.pack_xdt_weights(Wx.cuda(), Wdt.cuda(), dt)
    xs_tm = torch.empty(B, 4, L, D, device="cuda", dtype=dt)
    ops.dwconv3x3_silu_tm(xz.to("cuda", dt), 2 * D, wc.reshape(D, 9).t().contiguous().cuda(), bc.cuda(), xs_tm, B, H, W, D)
    S = ops.scan_tm_segments(B, D, H, W)
    carry = torch.empty(B * 4 * S * 2 * N * D, device="cuda")
    y = torch.empty(B, H * W, D, device="cuda", dtype=dt)
    if fuse:
        xdbl = torch.empty(B, 4, L, R + 2 * N, device="cuda")
        ops.x_proj_tm(xs_tm, xw16, xdbl, None, None, None, B, D, L, R, N, Rp, True)
        ops.selective_scan_tm(xs_tm, None, xdbl, A.cuda(), Wdt.reshape(4 * D, R).contiguous().cuda(), bias.cuda(), Dp.cuda(), carry, y,
                              B, D, H, W, N, R)
    else:
        xdbl = torch.empty(B, 4, L, 2 * N, device="cuda")
        dts = torch.empty(B, 4, L, D, device="cuda", dtype=dt)
        ops.x_proj_tm(xs_tm, xw16, xdbl, dw16, dts, bias.cuda(), B, D, L, R, N, Rp, False)
        ops.selective_scan_tm(xs_tm, dts, xdbl, A.cuda(), None, None, Dp.cuda(), carry, y, B, D, H, W, N, 0)
    r = rel(y.reshape(B, H, W, D), y_ref)
    print(f"time-major pipeline {cfg} {dt}: rel-L2 {r:.3e}")
    assert r < 2 * TOL[dt], r                   # three 16-bit roundings (xs, delta or x_dbl dt rows, y) against an all-fp32 reference
