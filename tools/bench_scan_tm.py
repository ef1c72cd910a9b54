"""A/B of the time-major scan variants at the benchmark's shapes (B = 16): segmented channel-per-lane scan with S segments vs the
time-sliced cooperative scan, with the reference's parameter initialisation (dt in [1e-3, 1e-1], A = -1 .. -N), CUDA events.
    python tools/bench_scan_tm.py [out.json]"""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import ops  # noqa: E402


def run(B, H, D, N, R, variants, dt=torch.bfloat16, iters=5):
    L = (H // 2) ** 2
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randn(B, 4, L, D, device="cuda", generator=g).to(dt)
    xdbl = torch.randn(B, 4, L, R + 2 * N, device="cuda", generator=g)
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(4 * D, 1).contiguous()
    dtv = torch.exp(torch.rand(4 * D, device="cuda", generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = (dtv + torch.log(-torch.expm1(-dtv))).contiguous()
    Wdt = ((torch.rand(4 * D, R, device="cuda", generator=g) * 2 - 1) * R ** -0.5 * 0.1).contiguous()
    Dp = torch.ones(4 * D, device="cuda")
    y = torch.empty(B, H * H, D, device="cuda", dtype=dt)
    carry = torch.empty(B * 4 * 64 * 2 * N * D, device="cuda")
    out = {}
    ref = None
    for v in variants:
        try:
            ops.selective_scan_tm(u, None, xdbl, A, Wdt, bias, Dp, carry, y, B, D, H, H, N, R, v)
        except Exception as e:
            out[str(v)] = f"unsupported ({e})"
            continue
        torch.cuda.synchronize()
        if ref is None:
            ref = y.float().clone()
        err = float((y.float() - ref).norm() / ref.norm())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            ops.selective_scan_tm(u, None, xdbl, A, Wdt, bias, Dp, carry, y, B, D, H, H, N, R, v)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / iters * 1e3
        byts = B * 4 * L * (2 * D * 2 + (R + 2 * N) * 4)
        out[str(v)] = {"us": round(us, 1), "GBps": round(byts / us / 1e3, 1), "state_updates_per_s_T": round(B * 4 * D * L * N / us / 1e6, 3),
                       "rel_vs_first": err}
        print(f"B{B} H{H} D{D} N{N} R{R} variant {v:>3}: {us:8.1f} us  {byts / us / 1e3:7.1f} GB/s  rel {err:.2e}", flush=True)
    return out


def run_chain(B, H, D, N, R, segs, dt=torch.bfloat16, iters=10):
    """Chained segments (fd_selective_scan_tm_chained) against the one-block-per-row kernel: FD_SCAN_CHAIN is read per call."""
    import os
    L = (H // 2) ** 2
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randn(B, 4, L, D, device="cuda", generator=g).to(dt)
    xdbl = torch.randn(B, 4, L, R + 2 * N, device="cuda", generator=g)
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(4 * D, 1).contiguous()
    dtv = torch.exp(torch.rand(4 * D, device="cuda", generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = (dtv + torch.log(-torch.expm1(-dtv))).contiguous()
    Wdt = ((torch.rand(4 * D, R, device="cuda", generator=g) * 2 - 1) * R ** -0.5 * 0.1).contiguous()
    Dp = torch.ones(4 * D, device="cuda")
    y0 = torch.empty(B, H * H, D, device="cuda", dtype=dt)
    plan = ops.scan_tm_plan(B, D, H, H, N, R)
    out = {}

    def timeit(fn):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters * 1e3

    us0 = timeit(lambda: ops.selective_scan_tm(u, None, xdbl, A, Wdt, bias, Dp, None, y0, B, D, H, H, N, R, plan))
    out["unchained"] = round(us0, 1)
    print(f"B{B} H{H} D{D} N{N} R{R} plan {plan} unchained: {us0:8.1f} us", flush=True)
    for n in segs:
        os.environ["FD_SCAN_CHAIN"] = str(n)
        nseg, floats = ops.scan_tm_chain_plan(B, D, H, H, N, R)
        if nseg <= 1:
            print(f"  FD_SCAN_CHAIN={n}: not chained", flush=True)
            continue
        ws = torch.zeros(floats, device="cuda")
        y = torch.empty_like(y0)
        us = timeit(lambda: ops.selective_scan_tm_chained(u, xdbl, A, Wdt, bias, Dp, ws, y, B, D, H, H, N, R))
        same = bool(torch.equal(y.view(torch.int16), y0.view(torch.int16)))
        out[f"x{nseg}"] = {"us": round(us, 1), "bit_identical": same}
        print(f"  chained x{nseg:<3}: {us:8.1f} us  ({us / us0:.3f} of unchained)  bit-identical {same}", flush=True)
    os.environ.pop("FD_SCAN_CHAIN", None)
    return out


if __name__ == "__main__":
    res = {}
    if "--chain-l1" in sys.argv:                     # with FD_SCAN_TW=8: the 512-row level on 8-warp blocks (needs chaining to balance)
        res["level1 16x256x16384 N8 R8"] = run_chain(16, 256, 256, 8, 8, [4, 8, 16, 32])
        sys.exit(0)
    if "--chain" in sys.argv:
        segs = [2, 4, 8, 16, 32, 64]
        res["level0 16x128x65536 N4 R4"] = run_chain(16, 512, 128, 4, 4, segs)
        res["level1 16x128x16384 N8 R4"] = run_chain(16, 256, 128, 8, 4, segs)
        res["level1 16x256x16384 N8 R8"] = run_chain(16, 256, 256, 8, 8, segs)
        json.dump(res, open("gpurun_out/scan_chain.json", "w"), indent=1)
        sys.exit(0)
    QUICK = [-1008, -8, -1004, -4, 1] if "--tw" in sys.argv else None      # K3b vs K3c only
    res["level0 16x128x65536 N4 R4"] = run(16, 512, 128, 4, 4, QUICK or [1, 4, 16, -8, -4, -1008])
    res["level1 16x128x16384 N8 R4"] = run(16, 256, 128, 8, 4, QUICK or [1, 8, -8, -4, -1008])
    res["level1 16x256x16384 N8 R8"] = run(16, 256, 256, 8, 8, QUICK or [1, 4, 8, -8, -4, -1004])
    res["level2 16x256x4096 N16 R8"] = run(16, 128, 256, 16, 8, [1, 2, 4])
    if len(sys.argv) > 1 and not sys.argv[1].startswith("--"):
        json.dump(res, open(sys.argv[1], "w"), indent=1)
