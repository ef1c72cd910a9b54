"""ctypes binding of oracle/selective_scan_ref.c (TEST INFRASTRUCTURE, see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfd_oracle_scan.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "selective_scan_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfd_oracle_scan.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.fd_oracle_selective_scan_fwd.restype = ctypes.c_int
        _lib.fd_oracle_selective_scan_fwd.argtypes = [ctypes.c_void_p] * 8 + [ctypes.c_int64] * 5 + [ctypes.c_int] * 3
    return _lib


def selective_scan_fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, acc64=False, nthreads=0):
    """u, delta: (b, KD, L); A: (KD, N); B, C: (b, K, N, L) or (b, N, L); D, delta_bias: (KD,). fp32 CPU."""
    lib = _load()
    f = lambda t: None if t is None else t.detach().to(torch.float32).contiguous().cpu()
    u, delta, A, B, C, D, delta_bias = map(f, (u, delta, A, B, C, D, delta_bias))
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    bt, dtot, L = u.shape
    N, G = A.shape[1], B.shape[1]
    assert delta.shape == u.shape and A.shape[0] == dtot and B.shape == (bt, G, N, L) and C.shape == B.shape
    y = torch.empty_like(u)
    p = lambda t: ctypes.c_void_p(0 if t is None else t.data_ptr())
    rc = lib.fd_oracle_selective_scan_fwd(p(u), p(delta), p(A), p(B), p(C), p(D), p(delta_bias), p(y),
                                          bt, dtot, L, N, G, int(bool(delta_softplus)), int(bool(acc64)), int(nthreads))
    if rc != 0:
        raise ValueError(f"fd_oracle_selective_scan_fwd rc={rc}")
    return y
