"""Head-to-head on the box: fd_selective_scan_fwd vs vLLM 0.22's build of the published state-spaces/mamba selective_scan_fwd
kernel (the lineage of the reference's selective_scan_cuda.fwd, src/emamba2.py:152) at the level geometries of BASELINE
config 5 (B = 16 slices of 512^2).  CUDA events, 3 warm-up + 10 timed launches, inputs >> L2.
    python tools/bench_scan_vs_published.py > gpurun_out/scan_vs_published.json"""
import json
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import ops  # noqa: E402
from oracle.gen_golden_scan_vllm import published_scan  # noqa: E402  (test infrastructure: the comparison arm)


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


rows = []
g = torch.Generator(device="cuda").manual_seed(5)
for name, (b, KD, L, N) in {"512^2 map (level 0)": (16, 512, 65536, 4), "256^2 map (level 1)": (16, 1024, 16384, 8),
                            "128^2 map (level 2)": (16, 2048, 4096, 16), "64^2 map (level 3)": (16, 4096, 1024, 32)}.items():
    u = torch.randn(b, KD, L, device="cuda", generator=g)
    delta = torch.randn(b, KD, L, device="cuda", generator=g) * 0.5
    A = -torch.exp(torch.randn(KD, N, device="cuda", generator=g) * 0.5)
    Bm, Cm = torch.randn(b, 4, N, L, device="cuda", generator=g), torch.randn(b, 4, N, L, device="cuda", generator=g)
    D, bias = torch.randn(KD, device="cuda", generator=g), torch.randn(KD, device="cuda", generator=g)
    y = torch.empty_like(u)
    row = dict(shape=name, batch=b, KD=KD, L=L, d_state=N)
    row["ours_fp32_ms"] = timed(lambda: ops.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True, y))
    out = delta.clone()
    states = torch.zeros(b, KD, N, device="cuda")
    from vllm import _custom_ops as vops

    def pub():
        vops.selective_scan_fwd(u, out, A, Bm, Cm, D, None, bias, True, None, None, None, states, -1)
    row["published_fp32_ms"] = timed(pub)
    ref = published_scan(u, delta, A, Bm, Cm, D, bias, True)
    ops.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True, y)
    row["rel_l2_fp32"] = float((y - ref).norm() / ref.norm())
    del out, ref
    ub, db, yb = u.bfloat16(), delta.bfloat16(), torch.empty(b, KD, L, device="cuda", dtype=torch.bfloat16)
    row["ours_bf16_io_ms"] = timed(lambda: ops.selective_scan_fwd(ub, db, A, Bm, Cm, D, bias, True, yb))
    try:
        Bb, Cb, outb = Bm.bfloat16(), Cm.bfloat16(), db.clone()

        def pubb():
            vops.selective_scan_fwd(ub, outb, A, Bb, Cb, D, None, bias, True, None, None, None, states, -1)
        row["published_bf16_io_ms"] = timed(pubb)
    except Exception as e:
        row["published_bf16_io_ms"] = None
        row["published_bf16_error"] = str(e)[:120]
    row["speedup_fp32"] = row["published_fp32_ms"] / row["ours_fp32_ms"]
    rows.append(row)
    del u, delta, y, ub, db, yb, Bm, Cm
    torch.cuda.empty_cache()
print(json.dumps(dict(device=torch.cuda.get_device_name(0), rows=rows), indent=1))
