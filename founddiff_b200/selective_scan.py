"""Drop-in for the third-party native op the reference imports (src/emamba2.py:23-34):

    from selective_scan_vmamba_pt202 import selective_scan_cuda_core     # VMamba build
    import selective_scan_cuda                                           # mamba_ssm build

Both signatures used at src/emamba2.py:152,154 are provided, backed by fd_selective_scan_fwd (sm_100a):

    selective_scan_cuda_core.fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows) -> (out, x)
    selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus)          -> (out, x)

`install()` registers this module under those names in sys.modules so that the unmodified reference
(`emamba2.SelectiveScan`) runs its scan on the B200 through this kernel.
Forward only: `x` (the per-chunk states the reference saves for backward) is returned as None and `bwd` raises.
"""
from __future__ import annotations

import sys
import types

import torch

from . import ops


def _fwd(u, delta, A, B, C, D, delta_bias, delta_softplus):
    if not u.is_cuda:
        raise RuntimeError("founddiff_b200.selective_scan: CUDA tensors required (no CPU path)")
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    f = lambda t: None if t is None else t.to(torch.float32).contiguous()  # noqa: E731
    io = u.dtype if u.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
    out = ops.selective_scan_fwd(u.to(io).contiguous(), delta.to(io).contiguous(), f(A), f(B), f(C), f(D), f(delta_bias),
                                 bool(delta_softplus))
    return out


class _Core:
    @staticmethod
    def fwd(u, delta, A, B, C, D, delta_bias, delta_softplus, nrows=1):
        return _fwd(u, delta, A, B, C, D, delta_bias, delta_softplus), None

    @staticmethod
    def bwd(*a, **k):
        raise NotImplementedError("founddiff_b200 implements the sampling (forward) path only")


selective_scan_cuda_core = _Core


def fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus):
    """mamba_ssm-style signature (src/emamba2.py:152)."""
    if z is not None:
        raise NotImplementedError("z gating is not used by FoundDiff (src/emamba2.py:152 passes None)")
    return _fwd(u, delta, A, B, C, D, delta_bias, delta_softplus), None


def bwd(*a, **k):
    raise NotImplementedError("founddiff_b200 implements the sampling (forward) path only")


def install():
    me = sys.modules[__name__]
    for name in ("selective_scan_vmamba_pt202", "selective_scan_vmamba"):
        m = types.ModuleType(name)
        m.selective_scan_cuda_core = _Core
        sys.modules[name] = m
    sys.modules["selective_scan_cuda"] = me
    sys.modules["selective_scan_cuda_core"] = me
