"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's evaluation metrics (src/util.py:188-236).  compute_ssim
there takes two primitives from kornia (`get_gaussian_kernel2d((11, 11), (1.5, 1.5))`, `filter2d`), which is NOT installed in
this container and not vendored by the reference: their published semantics are restated here (normalised 1-D Gaussian outer
product; `filter2d` = per-channel correlation with 'reflect' border, no kernel normalisation).  Pinned by
tests/golden/metrics.npz: the reference's OWN src/util.py functions run here with those two primitives bound to OpenCV's
independent implementations (oracle/gen_golden_metrics.py); what stays unpinned is kornia's code for the two primitives."""
import torch
import torch.nn.functional as F


def gaussian_kernel2d(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    x = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(x * x) / (2 * sigma * sigma))
    g = g / g.sum()
    return g[:, None] * g[None, :]


def filter2d(x: torch.Tensor, k: torch.Tensor) -> torch.Tensor:
    c = x.shape[1]
    pad = k.shape[-1] // 2
    xp = F.pad(x, (pad, pad, pad, pad), mode="reflect")
    return F.conv2d(xp, k[None, None].expand(c, 1, -1, -1), groups=c)


def compute_ssim(img1, img2, max_val: float = 1.0, reduction: str = "mean"):
    k = gaussian_kernel2d()
    C1, C2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    mu1, mu2 = filter2d(img1, k), filter2d(img2, k)
    s1 = filter2d(img1 * img1, k) - mu1 * mu1
    s2 = filter2d(img2 * img2, k) - mu2 * mu2
    s12 = filter2d(img1 * img2, k) - mu1 * mu2
    m = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2))
    m = m.clamp(0, 1)
    return m.mean() if reduction == "mean" else m.sum() if reduction == "sum" else m


def compute_psnr(a, b, max_val: float = 1.0):
    return 10 * torch.log10(torch.tensor(max_val * max_val) / F.mse_loss(a, b))


def compute_rmse(a, b):
    return torch.sqrt(F.mse_loss(a, b))
