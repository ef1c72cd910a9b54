"""SS2D tail at the full-resolution shape (B = 16, 512^2, D = 128 -> C = 64): fd_ln_gate + 1x1 conv (gate, addend) against the fused
fd_ln_gate_out_proj.  CUDA events, 10 launches each; bytes = algorithmic traffic of each form.
    python tools/bench_tail.py"""
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import ops  # noqa: E402

B, H, D, C = 16, 512, 128, 64
P = H * H
dti, dto = torch.bfloat16, torch.float16
g = torch.Generator(device="cuda").manual_seed(0)
y = torch.randn(B, P, D, device="cuda", generator=g).to(dti)
xz = torch.randn(B, P, 4 * C, device="cuda", generator=g).to(dti)
gamma, beta = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
local = torch.randn(B, D, device="cuda", generator=g) * 0.1
w = (torch.randn(C, D, device="cuda", generator=g) / D ** 0.5).to(dti)
gate = torch.randn(B, C, device="cuda", generator=g)
x_in = torch.randn(B, P, C, device="cuda", generator=g).to(dto)
x = torch.empty_like(x_in)
gbuf = torch.empty(B, P, D, device="cuda", dtype=dti)
conv = ops.Conv(gbuf, w, x, B=B, Hin=H, Win=H, gate=gate, gate_stride=C, addend=x_in, prefer_tc=True)


def timeit(fn, n=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


t1 = timeit(lambda: ops.ln_gate(y, xz, 4 * C, 2 * C, gamma, beta, local, gbuf, B, P, D))
t2 = timeit(conv.run)
ref = x.float().clone()
t3 = timeit(lambda: ops.ln_gate_out_proj(y, xz, 4 * C, 2 * C, gamma, beta, local, w, gate, C, x_in, x, B, P, D, C))
err = float((x.float() - ref).norm() / ref.norm())
es = 2
print(f"ln_gate {t1:.1f} us ({3 * B * P * D * es / t1 / 1e3:.0f} GB/s)  out_proj {t2:.1f} us ({B * P * (D + 2 * C) * es / t2 / 1e3:.0f} GB/s)  "
      f"sum {t1 + t2:.1f} us | fused {t3:.1f} us ({B * P * (2 * D + 2 * C) * es / t3 / 1e3:.0f} GB/s)  rel diff {err:.2e}")
