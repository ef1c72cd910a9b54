#!/usr/bin/env python
"""Micro-benchmarks of single hot-path kernels (BASELINE.json config 5: emamba2 selective scan + attention at 256^2 and
512^2 feature maps vs the HBM roofline).  Each kernel is timed alone with CUDA events (3 warm-ups, then `--iters`
launches; inputs are larger than L2 at batch 16), and reported as algorithmic GB/s and fraction of the measured HBM peak
(MEASURED_PEAKS.json burst figure — kernels timed in isolation).

    python bench_micro.py [--batch 16] [--iters 10] [--only scan|attn|xdt|ss2d|norm|conv] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# (KD, L, N) of the nine selective-scan calls of one Unet evaluation at 512^2
SCAN_SHAPES = [(512, 65536, 4), (512, 16384, 8), (1024, 4096, 16), (2048, 1024, 32), (4096, 1024, 32),
               (2048, 4096, 16), (1024, 16384, 8)]
LEVELS = [(64, 512, 4), (64, 256, 8), (128, 128, 16), (256, 64, 32), (512, 64, 32), (256, 128, 16), (128, 256, 8)]  # (C, H, N)


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--json", default="")
    ap.add_argument("--pick", default="", help="comma-separated indices into the shape list of the selected kernel")
    args = ap.parse_args()
    from founddiff_b200 import ops
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[args.dtype]
    es = 4 if dt == torch.float32 else 2
    B = args.batch
    pick = [int(i) for i in args.pick.split(",")] if args.pick else None
    sel = lambda lst: [x for i, x in enumerate(lst) if pick is None or i in pick]  # noqa: E731
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    rows = []

    def report(kernel, shape, ms, byts, flops=0.0):
        r = dict(kernel=kernel, shape=shape, us=round(ms * 1e3, 1), GBps=round(byts / ms / 1e6, 1), frac_hbm=round(byts / ms / 1e6 / hbm, 3),
                 GFLOP=round(flops / 1e9, 2), TFLOPs=round(flops / ms / 1e9, 2))
        rows.append(r)
        print(f"{kernel:18s} {shape:34s} {r['us']:10.1f} us {r['GBps']:9.1f} GB/s ({r['frac_hbm']:.3f} of {hbm:.0f})  {r['TFLOPs']:8.2f} TFLOP/s", flush=True)

    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s, d=dt: torch.randn(*s, device="cuda", generator=g).to(d)  # noqa: E731

    if args.only in ("", "scan"):
        for KD, L, N in sel(SCAN_SHAPES):
            u, delta = rn(B, KD, L), rn(B, KD, L) * 0.5
            A = -torch.exp(torch.randn(KD, N, device="cuda", generator=g) * 0.3)
            Bm, Cm = rn(B, 4, N, L, d=torch.float32), rn(B, 4, N, L, d=torch.float32)
            D, bias = rn(KD, d=torch.float32), rn(KD, d=torch.float32)
            y = torch.empty_like(u)
            ms = timeit(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True, out=y), args.iters)
            report("selective_scan", f"{B}x{KD}x{L} N{N}", ms, 3.0 * B * KD * L * es + 2.0 * B * 4 * N * L * 4, 9.0 * B * KD * L * N)
            if dt != torch.float32 and N >= 16 and B * KD >= 16384:      # deep levels: channel-per-lane kernel (+ fused merge)
                H2 = int(round(math.sqrt(L)))
                Bt, Ct = Bm.permute(0, 1, 3, 2).contiguous(), Cm.permute(0, 1, 3, 2).contiguous()
                yn = torch.empty(B, 4 * L, KD // 4, device="cuda", dtype=dt)
                ms = timeit(lambda: ops.selective_scan_fwd_merge_cl(u, delta, A, Bt, Ct, D, bias, True, yn, 2 * H2, 2 * H2), args.iters)
                report("selective_scan_cl", f"{B}x{KD}x{L} N{N}", ms, 3.0 * B * KD * L * es + 2.0 * B * 4 * N * L * 4, 9.0 * B * KD * L * N)
                del Bt, Ct, yn
            del u, delta, Bm, Cm, y
    if args.only in ("", "attn"):
        for C, H, _N in sel(LEVELS):
            qkv = rn(B, H * H, 3 * C)
            w = rn(3 * C, 9, d=torch.float32)
            gram = torch.zeros(B, C // 32, 32, 32, device="cuda")
            qk = torch.zeros(B, 2, C, device="cuda")
            if dt == torch.float32:
                v = torch.empty(B, H * H, C, device="cuda", dtype=dt)
                ms = timeit(lambda: ops.dwconv3x3_qkv_gram(qkv, w, v, gram, qk, B, H, H, C), args.iters)
                report("dwconv_qkv_gram", f"{B}x{H}x{H}x{C}", ms, 4.0 * B * H * H * C * es, (54.0 + 64.0) * B * H * H * C)
            else:
                qkv2 = torch.empty_like(qkv)
                wt = w.t().contiguous()
                ms = timeit(lambda: ops.dwconv3x3_nhwc(qkv, wt, None, qkv2, B, H, H, 3 * C), args.iters)
                report("dwconv3x3_nhwc", f"{B}x{H}x{H}x{3 * C}", ms, 6.0 * B * H * H * C * es, 54.0 * B * H * H * C)
                ms = timeit(lambda: ops.gram_qk(qkv2, 3 * C, gram, qk, B, H * H, C), args.iters)
                report("gram_qk", f"{B}x{H * H}x{C}", ms, 2.0 * B * H * H * C * es, 64.0 * 3 * B * H * H * C)
                del qkv2
            del qkv
    if args.only in ("", "xdt"):
        for C, H, N in sel(LEVELS):
            D, L, R = 2 * C, H * H // 4, math.ceil(C / 16)
            xs = rn(B, 4, D, L)
            Wx, Wd = rn(4, R + 2 * N, D, d=torch.float32), rn(4, D, R, d=torch.float32)
            dts = torch.empty_like(xs)
            Bs, Cs = torch.empty(B, 4, N, L, device="cuda"), torch.empty(B, 4, N, L, device="cuda")
            if dt == torch.float32:
                ms = timeit(lambda: ops.xdt_proj(xs, Wx, Wd, dts, Bs, Cs, B, D, L, R, N), args.iters)
                name = "xdt_proj"
            else:           # the 16-bit path: tensor-core x_proj + dt_proj, cp.async ring
                xw16, dw16, Rp = ops.pack_xdt_weights(Wx, Wd, dt)
                ms = timeit(lambda: ops.xdt_proj_tc(xs, xw16, dw16, Rp, dts, Bs, Cs, B, D, L, R, N), args.iters)
                name = "xdt_proj_tc"
            report(name, f"{B}x{D}x{L} R{R} N{N}", ms, 2.0 * B * 4 * D * L * es + 2.0 * B * 4 * N * L * 4, 2.0 * B * 4 * L * D * (2 * R + 2 * N))
            del xs, dts, Bs, Cs
    if args.only in ("", "ss2d"):
        for C, H, N in LEVELS[:3]:
            D, L = 2 * C, H * H // 4
            xz = rn(B, H * H, 4 * C)
            w, b = rn(D, 9, d=torch.float32), rn(D, d=torch.float32)
            xs = torch.empty(B, 4, D, L, device="cuda", dtype=dt)
            ms = timeit(lambda: ops.dwconv3x3_silu_scan(xz, 4 * C, w, b, xs, B, H, H, D), args.iters)
            report("dwconv_scan", f"{B}x{H}x{H}x{D}", ms, 2.0 * B * H * H * D * es)
            gm, bt, loc = rn(D, d=torch.float32), rn(D, d=torch.float32), rn(B, D, d=torch.float32)
            stat = torch.empty(B, H * H, 2, device="cuda")
            out = torch.empty(B, H * H, D, device="cuda", dtype=dt)
            ms = timeit(lambda: ops.merge_ln_gate(xs, xz, 4 * C, D, gm, bt, loc, stat, out, B, H, H, D), args.iters)
            report("merge_ln_gate", f"{B}x{H}x{H}x{D}", ms, 3.0 * B * H * H * D * es)
            ynhwc = rn(B, H * H, D)
            ms = timeit(lambda: ops.ln_gate(ynhwc, xz, 4 * C, 2 * C, gm, bt, loc, out, B, H * H, D), args.iters)
            report("ln_gate", f"{B}x{H * H}x{D}", ms, 3.0 * B * H * H * D * es)
            del ynhwc
            del xz, xs, out
    if args.only in ("", "tm") and dt != torch.float32:
        # the time-major SS2D chain the 16-bit engine runs (fd_ss2d_tm.cu): depthwise conv -> x_proj (-> dt_proj) -> scan + merge, with
        # the reference's parameter initialisation (dt in [1e-3, 1e-1], A = -1 .. -N: src/emamba2.py:548-574), which decides how far
        # a carry has to be walked.  Bytes: SURVEY 8(d) stage figures.
        for C, H, N in sel(LEVELS):
            D, L, R = 2 * C, H * H // 4, math.ceil(C / 16)
            fuse = (N, R) in ((4, 4), (8, 4), (8, 8), (16, 8))
            xz = rn(B, H * H, 4 * C)
            w9, bcv = rn(9, D, d=torch.float32) * 0.3, rn(D, d=torch.float32) * 0.1
            xs = torch.empty(B, 4, L, D, device="cuda", dtype=dt)
            ms = timeit(lambda: ops.dwconv3x3_silu_tm(xz, 4 * C, w9, bcv, xs, B, H, H, D), args.iters)
            report("dwconv_tm", f"{B}x{H}x{H}x{D}", ms, 2.0 * B * H * H * D * es, 18.0 * B * H * H * D)
            Wx, Wd = rn(4, R + 2 * N, D, d=torch.float32) / math.sqrt(D), (torch.rand(4, D, R, device="cuda", generator=g) * 2 - 1) * R ** -0.5
            xw16, dw16, Rp = ops.pack_xdt_weights(Wx, Wd, dt)
            dtv = torch.exp(torch.rand(4 * D, device="cuda", generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
            bias = (dtv + torch.log(-torch.expm1(-dtv))).contiguous()
            A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(4 * D, 1).contiguous()
            Dp = torch.ones(4 * D, device="cuda")
            xdbl = torch.empty(B, 4, L, (R + 2 * N) if fuse else 2 * N, device="cuda")
            dts = None if fuse else torch.empty(B, 4, L, D, device="cuda", dtype=dt)
            ms = timeit(lambda: ops.x_proj_tm(xs, xw16, xdbl, None if fuse else dw16, dts, None if fuse else bias, B, D, L, R, N, Rp, fuse), args.iters)
            report("x_proj_tm" + ("" if fuse else "+dt"), f"{B}x{D}x{L} R{R} N{N}", ms,
                   (B * 4.0 * D * L * es + B * 4.0 * (R + 2 * N) * L * 4) if fuse else (2.0 * B * 4 * D * L * es + 2.0 * B * 4 * N * L * 4),
                   2.0 * B * 4 * L * D * ((R + 2 * N) if fuse else (2 * R + 2 * N)))
            plan = ops.scan_tm_plan(B, D, H, H, N, R if fuse else 0)
            carry = torch.empty(B * 4 * max(plan, 1) * 2 * N * D, device="cuda")
            y = torch.empty(B, H * H, D, device="cuda", dtype=dt)
            dtw = Wd.reshape(4 * D, R).contiguous()
            ms = timeit(lambda: ops.selective_scan_tm(xs, dts, xdbl, A, dtw if fuse else None, bias if fuse else None, Dp, carry, y, B, D, H, H,
                                                      N, R if fuse else 0, plan), args.iters)
            report(f"scan_tm {'TW' + str(-plan) if plan < 0 else 'S' + str(plan)}", f"{B}x{4 * D}x{L} N{N}", ms,
                   3.0 * B * 4 * D * L * es + 2.0 * B * 4 * N * L * 4, 9.0 * B * 4 * D * L * N)
            nseg, ws_floats = ops.scan_tm_chain_plan(B, D, H, H, N, R) if fuse else (0, 0)
            if nseg > 1:                     # what the engine launches at this level: the same kernel with chained segments
                ws = torch.zeros(ws_floats, device="cuda")
                ms = timeit(lambda: ops.selective_scan_tm_chained(xs, xdbl, A, dtw, bias, Dp, ws, y, B, D, H, H, N, R), args.iters)
                report(f"scan_tm TW{-plan} x{nseg} chained", f"{B}x{4 * D}x{L} N{N}", ms,
                       3.0 * B * 4 * D * L * es + 2.0 * B * 4 * N * L * 4, 9.0 * B * 4 * D * L * N)
            del xz, xs, xdbl, dts, carry, y
    if args.only in ("", "norm"):
        from founddiff_b200.engine import _view_ptr
        for C, H, _ in LEVELS[:3]:
            P = H * H
            x, out = rn(B, P, C), torch.empty(B, P, C, device="cuda", dtype=dt)
            mods = rn(B, 6 * C, d=torch.float32)
            gm, bt = rn(C, d=torch.float32), rn(C, d=torch.float32)
            ms = timeit(lambda: ops.ln_modulate(x, out, gm, bt, _view_ptr(mods[:, :C]), _view_ptr(mods[:, C:2 * C]), 6 * C, B, P, C, 1e-5), args.iters)
            report("ln_modulate", f"{B}x{P}x{C}", ms, 2.0 * B * P * C * es)
            sums = torch.zeros(B, 8, 2, device="cuda")
            ops.gn_stats(x, sums, B, P, C, 8)
            skip = rn(B, P, C)               # its own tensor: with skip == x the kernel moved 2 tensors while 3 were counted
            ms = timeit(lambda: ops.gn_silu_add(x, sums, gm, bt, skip, out, B, P, C, 8), args.iters)
            report("gn_silu_add", f"{B}x{P}x{C}", ms, 3.0 * B * P * C * es)
            del skip
    if args.only in ("", "conv"):
        # (c0, c1, cout, k, stride, up, H, epilogue) of representative call sites of one Unet evaluation
        CONVS = [(64, 0, 256, 1, 1, False, 512, "silu_half"), (64, 0, 192, 1, 1, False, 512, "plain"),
                 (128, 0, 64, 1, 1, False, 512, "gate_res"), (64, 64, 64, 3, 1, False, 512, "gn"),
                 (64, 0, 64, 3, 1, False, 512, "gn"), (256, 128, 256, 3, 1, False, 128, "gn"),
                 (512, 0, 256, 3, 1, True, 64, "plain"), (64, 0, 64, 4, 2, False, 512, "plain")]
        for c0, c1, cout, k, stride, up, H, epi in sel(CONVS):
            x0 = rn(B, H * H, c0)
            x1 = rn(B, H * H, c1) if c1 else None
            w = rn(cout, k, k, c0 + c1) * 0.05
            Ho = H * (2 if up else 1) // stride
            out = torch.empty(B, Ho * Ho, cout, device="cuda", dtype=dt)
            kw = {}
            if epi == "silu_half":
                kw = dict(silu_from=cout // 2)
            elif epi == "gate_res":
                gate = rn(B, cout, d=torch.float32)
                kw = dict(gate=gate, gate_stride=cout, addend=out)
            elif epi == "gn":
                kw = dict(bias=rn(cout, d=torch.float32), gn_sums=torch.zeros(B, 8, 2, device="cuda"), gn_groups=8)
            conv = ops.Conv(x0, w, out, B=B, Hin=H, Win=H, KH=k, KW=k, stride=stride, pad=(k - 1) // 2 if k != 4 else 1, upsample=up,
                            src1=x1, **kw)
            ms = timeit(conv.run, args.iters)
            flops = conv.flops()
            byts = (B * H * H * (c0 + c1) + B * Ho * Ho * cout * (2 if epi == "gate_res" else 1)) * es
            report("conv_tc" if conv.uses_tc else "conv_simt", conv.describe() + " " + epi, ms, byts, flops)
            del x0, x1, out, conv
    if args.only in ("", "flash"):
        # lucidrains bottleneck Attention: 4 heads x 32 over (H/8)^2 tokens for 256^2 and 512^2 inputs
        for N in sel([1024, 4096]):
            qkv = rn(B, N, 3 * 128)
            out = torch.empty(B, N, 128, device="cuda", dtype=dt)
            for impl in ("tc", "mma"):
                ms = timeit(lambda: ops.flash_attn_d32(qkv, out, B, N, 4, 32 ** -0.5, impl=impl), args.iters)
                report("flash_attn_d32" + ("_tc (tcgen05)" if impl == "tc" else " (mma.sync)"), f"{B}x{N}x4x32", ms, 4.0 * B * N * 128 * es,
                       4.0 * B * 4 * N * N * 32)
    if args.only in ("", "linattn"):
        # lucidrains LinearAttention (4 heads x 32) at the 256^2 and 512^2 feature maps of the first level (dim 64)
        for H in sel([256, 512]):
            qkv = rn(B, H * H, 3 * 128)
            wout, bias, gam = rn(64, 128, d=torch.float32) * 0.1, rn(64, d=torch.float32), rn(64, d=torch.float32)
            out = torch.empty(B, H * H, 64, device="cuda", dtype=dt)
            ms = timeit(lambda: ops.linear_attention(qkv, wout, bias, gam, out, B, H, H, 4, 64), max(2, args.iters // 3))
            report("linear_attention", f"{B}x{H}x{H} 4x32->64", ms, B * H * H * (2 * 256 + 128 + 128 + 2 * 64 + 2 * 64 + 128) * es / 1.0,
                   2.0 * B * H * H * (128 * 32 + 128 * 64))
            del qkv, out
    if args.json:
        os.makedirs(os.path.dirname(os.path.abspath(args.json)), exist_ok=True)
        json.dump(dict(batch=B, dtype=args.dtype, hbm_peak_gbs=hbm, rows=rows), open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
