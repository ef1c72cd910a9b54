"""fd_slice_metrics at the sizes where a partial tile's halo used to fold outside the image: 16x16, 40x33, 33x50, 9x34."""
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import metrics  # noqa: E402
from oracle import metrics_oracle as MO  # noqa: E402

for H, W in ((16, 16), (40, 33), (33, 50), (9, 34), (6, 6)):
    g = torch.Generator().manual_seed(H * W)
    a, b = torch.rand(2, 1, H, W, generator=g), torch.rand(2, 1, H, W, generator=g)
    psnr, ssim, rmse = metrics.slice_metrics(a.cuda(), b.cuda())
    for i in range(2):
        assert abs(float(psnr[i]) - float(MO.compute_psnr(a[i:i + 1], b[i:i + 1]))) < 1e-3, (H, W)
        assert abs(float(ssim[i]) - float(MO.compute_ssim(a[i:i + 1], b[i:i + 1]))) < 5e-5, (H, W)
    print("ok", H, W)
