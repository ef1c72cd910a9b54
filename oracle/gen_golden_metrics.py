"""Generate tests/golden/metrics.npz: the reference's OWN `compute_ssim / compute_psnr / compute_rmse` (src/util.py:188-236)
run in this container.  src/util.py takes two primitives from kornia (`get_gaussian_kernel2d`, `filter2d`), which is not
installed; they are bound here to OpenCV's equivalents — an independent third-party implementation of the same published
semantics (cv2.getGaussianKernel = normalised exp(-x^2 / 2 sigma^2); cv2.filter2D = correlation, BORDER_REFLECT_101 = kornia's
default border_type='reflect', computed in float64) — NOT to oracle/metrics_oracle.py, so the fixture is independent of the
restatement it pins.  TEST INFRASTRUCTURE.   Usage:  python -m oracle.gen_golden_metrics
"""
import os
import sys

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle.gen_golden import npz, synth_slices  # noqa: E402


def get_gaussian_kernel2d(kernel_size, sigma):
    ky = cv2.getGaussianKernel(kernel_size[0], sigma[0], cv2.CV_64F)
    kx = cv2.getGaussianKernel(kernel_size[1], sigma[1], cv2.CV_64F)
    return torch.from_numpy(ky @ kx.T)


def filter2d(x, kernel):
    k = kernel[0].double().numpy()
    out = np.empty(x.shape, dtype=np.float64)
    xn = x.double().numpy()
    for b in range(x.shape[0]):
        for c in range(x.shape[1]):
            out[b, c] = cv2.filter2D(xn[b, c], cv2.CV_64F, k, borderType=cv2.BORDER_REFLECT_101)
    return torch.from_numpy(out).to(x.dtype)


def main():
    ref_shims.install()
    import kornia.filters as kf                       # the stub module fabricated by ref_shims
    kf.get_gaussian_kernel2d, kf.filter2d = get_gaussian_kernel2d, filter2d
    sys.modules.pop("src.util", None)
    import src.util as U
    U.get_gaussian_kernel2d, U.filter2d = get_gaussian_kernel2d, filter2d
    fx = {}
    for tag, (B, H, W, sigma) in dict(a=(2, 64, 96, 0.05), b=(1, 33, 50, 0.2), c=(1, 128, 128, 0.01)).items():
        ndct, ldct = synth_slices(B, H, W, seed=50 + H, sigma=sigma)
        fx[f"{tag}.pred"], fx[f"{tag}.target"] = ldct, ndct
        for i in range(B):                             # per slice, as Trainer.test calls them (src/DADiff.py:1883-1888)
            p, t = ldct[i:i + 1], ndct[i:i + 1]
            fx[f"{tag}.{i}.ssim"] = U.compute_ssim(p, t)
            fx[f"{tag}.{i}.psnr"] = U.compute_psnr(p, t)
            fx[f"{tag}.{i}.rmse"] = U.compute_rmse(p, t)
            print(tag, i, float(fx[f"{tag}.{i}.ssim"]), float(fx[f"{tag}.{i}.psnr"]), float(fx[f"{tag}.{i}.rmse"]))
    npz("metrics.npz", **fx)


if __name__ == "__main__":
    main()
