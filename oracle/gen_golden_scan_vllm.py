"""Pin the selective scan against a BUILD OF THE PUBLISHED THIRD-PARTY KERNEL.

The reference calls `selective_scan_cuda.fwd` / `selective_scan_cuda_core.fwd` (src/emamba2.py:27-34, 152-154): the
state-spaces/mamba CUDA kernel (and VMamba's fork of it), un-vendored and un-pinned.  vLLM 0.22 (in this image) ships
its own build of that kernel — `torch.ops._C.selective_scan_fwd`, csrc/mamba/mamba_ssm/selective_scan_fwd.cu, "adapted
from state-spaces/mamba" — with the same semantics (grouped B / C, delta_bias, softplus with threshold 20, D skip).
It needs a GPU, so this script runs ON THE GPU BOX:

    gpurun -- python -m oracle.gen_golden_scan_vllm          # writes gpurun_out/scan_vllm.npz

and the result is committed as tests/golden/scan_vllm.npz: the inputs of tests/golden/scan.npz (tags a, b, c) plus a
longer grouped case (tag d), each with `y_vllm` = the published kernel's fp32 output.  TEST INFRASTRUCTURE.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def published_scan(u, delta, A, B, C, D, bias, softplus=True):
    """vLLM's build of the mamba selective_scan_fwd kernel; the output is written in place of `delta`."""
    from vllm import _custom_ops as vops
    out = delta.clone().contiguous()
    states = torch.zeros(u.shape[0], u.shape[1], A.shape[1], device=u.device, dtype=torch.float32)
    vops.selective_scan_fwd(u.clone().contiguous(), out, A.contiguous(), B.contiguous(), C.contiguous(), D, None, bias, softplus,
                            None, None, None, states, -1)
    return out


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "scan.npz"))
    fx = {}
    cases = {}
    for tag in "abc":
        cases[tag] = {k: torch.from_numpy(z[f"{tag}.{k}"]) for k in ("u", "delta", "A", "B", "C", "D", "bias")}
    g = torch.Generator().manual_seed(23)
    b, K, Dk, N, L = 1, 4, 8, 8, 1536
    d = dict(u=torch.randn(b, K * Dk, L, generator=g), delta=torch.randn(b, K * Dk, L, generator=g) * 2 - 1,
             A=-torch.exp(torch.randn(K * Dk, N, generator=g) * 0.5), B=torch.randn(b, K, N, L, generator=g),
             C=torch.randn(b, K, N, L, generator=g), D=torch.randn(K * Dk, generator=g), bias=torch.randn(K * Dk, generator=g))
    d["delta"][0, 1, :7] = 30.0                     # softplus threshold branch
    cases["d"] = d
    for tag, c in cases.items():
        cu = {k: v.cuda() for k, v in c.items()}
        y = published_scan(cu["u"], cu["delta"], cu["A"], cu["B"], cu["C"], cu["D"], cu["bias"], True)
        torch.cuda.synchronize()
        fx[f"{tag}.y_vllm"] = y.cpu().numpy()
        if tag == "d":
            fx.update({f"d.{k}": v.numpy() for k, v in c.items()})
        if f"{tag}.y64" in z.files:
            y64 = torch.from_numpy(z[f"{tag}.y64"])
            print(tag, "published kernel vs fp64 sequential recurrence: rel-L2",
                  float((y.cpu().double() - y64).norm() / y64.norm()))
    import vllm
    fx["vllm_version"] = np.array(vllm.__version__)
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "scan_vllm.npz"), **fx)
    print("wrote gpurun_out/scan_vllm.npz", vllm.__version__)


if __name__ == "__main__":
    main()
