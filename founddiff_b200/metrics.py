"""`compute_psnr` / `compute_ssim` / `compute_rmse` of the reference's test loop (src/util.py:188-236, called at
src/DADiff.py:1883-1888) on the device, fused into one kernel (fd_slice_metrics): the denoised slices never leave the GPU
for evaluation.  Same names and argument meaning as the reference functions; `slice_metrics` returns the three per slice.
"""
from __future__ import annotations

import torch

from . import ops


def _check(a: torch.Tensor, b: torch.Tensor):
    if not torch.is_tensor(a) or not torch.is_tensor(b):
        raise TypeError(f"Expected 2 torch tensors but got {type(a)} and {type(b)}")
    if a.shape != b.shape:
        raise TypeError(f"Expected tensors of equal shapes, but got {a.shape} and {b.shape}")
    if not a.is_cuda:
        raise RuntimeError("founddiff_b200 has no CPU path: pass CUDA tensors")


def _raw(pred: torch.Tensor, target: torch.Tensor, max_val: float):
    """(n_images, 2) fp32: [sum of squared errors, sum of the SSIM map] per (batch, channel) image."""
    _check(pred, target)
    H, W = pred.shape[-2:]
    n = pred.numel() // (H * W)
    p = pred.detach().to(torch.float32).reshape(n, H, W).contiguous()
    t = target.detach().to(torch.float32).reshape(n, H, W).contiguous()
    out = torch.empty(n, 2, device=pred.device, dtype=torch.float32)
    ops.slice_metrics(p, t, out, n, H, W, max_val)
    return out, H * W


def compute_psnr(input: torch.Tensor, target: torch.Tensor, max_val: float = 1.0) -> torch.Tensor:
    """src/util.py:223-232: 10 log10(max_val^2 / mse) with the MSE over the WHOLE tensor."""
    raw, hw = _raw(input, target, max_val)
    mse = raw[:, 0].sum() / (raw.shape[0] * hw)
    return 10 * torch.log10(max_val * max_val / mse)


def compute_rmse(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """src/util.py:235-236."""
    raw, hw = _raw(input, target, 1.0)
    return torch.sqrt(raw[:, 0].sum() / (raw.shape[0] * hw))


def compute_ssim(img1: torch.Tensor, img2: torch.Tensor, window_size: int = 11, reduction: str = "mean", max_val: float = 1.0,
                 full: bool = False) -> torch.Tensor:
    """src/util.py:188-220 with the defaults the test loop uses (11x11 window, 'mean' or 'sum')."""
    if window_size != 11 or full or reduction not in ("mean", "sum"):
        raise NotImplementedError("only window_size=11, reduction in {'mean','sum'}, full=False (the reference's call) is implemented")
    raw, hw = _raw(img1, img2, max_val)
    s = raw[:, 1].sum()
    return s / (raw.shape[0] * hw) if reduction == "mean" else s


def slice_metrics(pred: torch.Tensor, target: torch.Tensor, max_val: float = 1.0):
    """Per-slice (psnr, ssim, rmse), each of shape (n_images,): what the reference's batch-1 test loop appends per slice."""
    raw, hw = _raw(pred, target, max_val)
    mse = raw[:, 0] / hw
    return 10 * torch.log10(max_val * max_val / mse), raw[:, 1] / hw, torch.sqrt(mse)
