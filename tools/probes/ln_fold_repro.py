"""Is the LayerNorm-folded GEMM reproducible run to run?  Same inputs, 30 launches, bitwise comparison with the first output."""
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import ops  # noqa: E402


class V:
    def __init__(self, t):
        self.t, self.dtype, self.is_cuda = t, torch.float32, True

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self.t.data_ptr()


for (B, H, C, Co) in ((6, 64, 128, 256), (6, 64, 128, 512), (6, 64, 128, 128), (6, 64, 64, 512), (16, 256, 128, 512)):
    P = H * H
    x = (torch.randn(B, P, C, device="cuda") + 0.3).to(torch.float16)
    W = torch.randn(Co, C, device="cuda") / 8
    out = torch.empty(B, P, Co, device="cuda", dtype=torch.bfloat16)
    wf = torch.empty(B, Co, C, device="cuda", dtype=torch.float16)
    v = torch.zeros(B, Co, device="cuda")
    mods = torch.randn(B, 2 * C, device="cuda") * 0.1
    ops.ln_fold(W, None, None, V(mods[:, :C]), V(mods[:, C:]), 2 * C, wf, v, B, Co, C)
    conv = ops.Conv(x, wf, out, B=B, Hin=H, Win=H, silu_from=Co // 2, per_batch_weight=True, ln_v=v, ln_eps=1e-6)
    conv.run()
    torch.cuda.synchronize()
    first = out.clone()
    bad = 0
    for i in range(40):
        out.fill_(float("nan"))
        conv.run()
        torch.cuda.synchronize()
        if not torch.equal(out, first):
            bad += 1
            d = (out.float() - first.float()).abs()
            idx = torch.nonzero(d > 0)
            if bad <= 2: print(f"  run {i}: {idx.shape[0]} elements differ, max {float(d.max()):.3e}; first at (b, p, n) = {idx[0].tolist()}, rows touched: "
                  f"{torch.unique(idx[:, 1] % 128)[:16].tolist()}")
    print(f"B={B} H={H} C={C} Cout={Co}: {bad} of 40 launches differ", flush=True)
