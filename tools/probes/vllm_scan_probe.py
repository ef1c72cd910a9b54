"""Probe: is vLLM's build of the state-spaces/mamba selective_scan_fwd CUDA kernel callable on this box, and does it agree
with fd_selective_scan_fwd?  (Test infrastructure; run on the GPU box.)"""
import sys
import time
import traceback

import torch

sys.path.insert(0, ".")
t0 = time.time()
try:
    from vllm import _custom_ops as vops
    print("vllm._custom_ops imported in %.1fs" % (time.time() - t0), flush=True)
    print("has op:", hasattr(torch.ops._C, "selective_scan_fwd"), flush=True)
except Exception:
    traceback.print_exc()
    sys.exit(0)

from founddiff_b200 import ops  # noqa: E402

g = torch.Generator().manual_seed(11)
for (b, K, Dk, N, L) in [(2, 4, 8, 4, 300), (1, 4, 16, 16, 65), (2, 4, 32, 8, 4096), (2, 1, 64, 16, 2048)]:
    u = torch.randn(b, K * Dk, L, generator=g).cuda()
    delta = (torch.randn(b, K * Dk, L, generator=g) * 2).cuda()
    A = (-torch.exp(torch.randn(K * Dk, N, generator=g) * 0.5)).cuda()
    Bm = torch.randn(b, K, N, L, generator=g).cuda()
    Cm = torch.randn(b, K, N, L, generator=g).cuda()
    D = torch.randn(K * Dk, generator=g).cuda()
    bias = torch.randn(K * Dk, generator=g).cuda()
    mine = torch.empty_like(u)
    ops.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True, mine)
    try:
        out = delta.clone()
        states = torch.zeros(b, K * Dk, N, device="cuda")
        vops.selective_scan_fwd(u.clone(), out, A, Bm, Cm, D, None, bias, True, None, None, None, states, -1)
        torch.cuda.synchronize()
        err = (mine - out).norm() / out.norm()
        print((b, K, Dk, N, L), "rel-L2 ours vs vllm:", float(err), "max abs", float((mine - out).abs().max()), flush=True)
    except Exception:
        traceback.print_exc()
