"""in_proj-shaped 1x1 GEMM (16 x 512^2, 64 -> 256, SiLU on the upper half) with and without the LayerNorm fold: timing + ncu target."""
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import ops  # noqa: E402

B, H, C, Co = 16, 512, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = H * H
x = torch.randn(B, P, C, device="cuda").to(torch.float16)
a = torch.randn(B, P, C, device="cuda").to(torch.float16)
W = (torch.randn(Co, C, device="cuda") / 8)
out = torch.empty(B, P, Co, device="cuda", dtype=torch.bfloat16)
wf = torch.empty(B, Co, C, device="cuda", dtype=torch.float16)
v = torch.zeros(B, Co, device="cuda")
mods = torch.randn(B, 2 * C, device="cuda") * 0.1


class V:
    def __init__(self, t):
        self.t, self.dtype, self.is_cuda = t, torch.float32, True

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self.t.data_ptr()


ops.ln_fold(W, None, None, V(mods[:, :C]), V(mods[:, C:]), 2 * C, wf, v, B, Co, C)
plain = ops.Conv(a, W.to(torch.float16), out, B=B, Hin=H, Win=H, silu_from=Co // 2)
pbw = ops.Conv(a, wf, out, B=B, Hin=H, Win=H, silu_from=Co // 2, per_batch_weight=True)
fold = ops.Conv(x, wf, out, B=B, Hin=H, Win=H, silu_from=Co // 2, per_batch_weight=True, ln_v=v, ln_eps=1e-5)
rstd = torch.empty(B, P, device="cuda")
ops.row_rstd(x, rstd, B * P, C, 1e-5)
fold2 = ops.Conv(x, wf, out, B=B, Hin=H, Win=H, silu_from=Co // 2, per_batch_weight=True, ln_v=v, ln_eps=1e-5, ln_rstd=rstd)
for name, c in (("plain", plain), ("per-batch weight", pbw), ("ln fold", fold), ("ln fold, external rstd", fold2)):
    for _ in range(2):
        c.run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c.run()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:18s} {e0.elapsed_time(e1) / 5 * 1e3:8.1f} us", flush=True)
