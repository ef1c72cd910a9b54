"""One launch of each time-major scan variant at the level-0 shape, for ncu (tools/bench_scan_tm.py is the timing script).
    python tools/prof_scan_tm.py -8        # the one-block-per-row time-sliced kernel
    python tools/prof_scan_tm.py chain     # the chained-segments launch the engine uses (default segment count)"""
import sys
sys.path.insert(0, ".")
import tools.bench_scan_tm as b  # noqa: E402

if "chain" in sys.argv:
    import math
    import torch
    from founddiff_b200 import ops
    B, H, D, N, R, dt = 16, 512, 128, 4, 4, torch.bfloat16
    L = (H // 2) ** 2
    g = torch.Generator(device="cuda").manual_seed(0)
    u = torch.randn(B, 4, L, D, device="cuda", generator=g).to(dt)
    xdbl = torch.randn(B, 4, L, R + 2 * N, device="cuda", generator=g)
    A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(4 * D, 1).contiguous()
    dtv = torch.exp(torch.rand(4 * D, device="cuda", generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = (dtv + torch.log(-torch.expm1(-dtv))).contiguous()
    Wdt = ((torch.rand(4 * D, R, device="cuda", generator=g) * 2 - 1) * R ** -0.5 * 0.1).contiguous()
    nseg, floats = ops.scan_tm_chain_plan(B, D, H, H, N, R)
    ws, y = torch.zeros(floats, device="cuda"), torch.empty(B, H * H, D, device="cuda", dtype=dt)
    ops.selective_scan_tm_chained(u, xdbl, A, Wdt, bias, torch.ones(4 * D, device="cuda"), ws, y, B, D, H, H, N, R)
    torch.cuda.synchronize()
    print(f"chained x{nseg}")
else:
    variants = [int(v) for v in sys.argv[1:]] or [16, -8]
    b.run(16, 512, 128, 4, 4, variants, iters=1)
