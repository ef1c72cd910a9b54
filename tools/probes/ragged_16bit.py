"""Re-validation probe for 16-bit storage on an odd deepest-level scan length (48x80 -> L = 15), which UnetEngine.check_geometry
refuses by default.  Run on the GPU box, ideally under compute-sanitizer (tools/sanitize.sh):
    FD_ALLOW_ODD_16BIT=1 python tools/probes/ragged_16bit.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("FD_ALLOW_ODD_16BIT", "1")
from founddiff_b200 import weights  # noqa: E402
from founddiff_b200.diffusion import ResidualDiffusion, UnetRes  # noqa: E402

z = np.load(os.path.join(ROOT, "tests", "golden", "unet_48x80.npz"))
g = {k: torch.from_numpy(z[k]) for k in z.files}
m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, objective='pred_res', test_res_or_noise='res')
m.load_state_dict({"unet0." + k: v for k, v in weights.random_state_dict(10).items()})
d = ResidualDiffusion(m, image_size=48, sampling_timesteps=2, objective='pred_res', condition=True, sum_scale=0.01).cuda()
d.init()
for dt in (torch.float32, torch.float16, torch.bfloat16):
    m.compute_dtype = dt
    time = g["time"].cuda()
    out = m(g["x_in"].cuda(), [time, time])[0].cpu()
    r = float((out - g["out"]).norm() / g["out"].norm())
    outs = d.sample([g["ldct"].cuda()], batch_size=1, last=False, noise={"init": g["init_noise"]})
    torch.cuda.synchronize()
    r2 = float((outs[-1].cpu() - g["outs"][-1]).norm() / g["outs"][-1].norm())
    print(f"{dt}: Unet rel-L2 {r:.3e} (gate {1e-3 if dt == torch.float32 else 1e-2}), final image rel-L2 {r2:.3e}")
