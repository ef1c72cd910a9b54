"""Generates tests/golden/gaussian_*.npz from the UNMODIFIED reference (src/denoising_diffusion_pytorch.py) imported from
/root/reference with the import shims of oracle/ref_shims.py.  Run in the build container only (the GPU box has no
/root/reference); the fixtures are committed.

    python oracle/gen_golden_gaussian.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
sys.path.insert(0, "/root/reference")
from src.denoising_diffusion_pytorch import GaussianDiffusion, Unet  # noqa: E402

from founddiff_b200.gaussian import random_gaussian_state_dict  # noqa: E402
from oracle import gaussian_oracle as G  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.manual_seed(0)
torch.set_grad_enabled(False)

sd = random_gaussian_state_dict(11)
model = Unet(dim=64, dim_mults=(1, 2, 4, 8)).eval()
missing = model.load_state_dict(sd, strict=True)           # proves the schema's key names / shapes are the reference's
print("state dict loaded into the reference Unet:", missing)

# ---- Unet forward, non-square, two timesteps ---------------------------------------------------------------------
g = torch.Generator().manual_seed(5)
x = torch.randn(2, 3, 32, 48, generator=g)
fx = {"x": x.numpy()}
for t in (999, 250):
    time = torch.full((2,), t, dtype=torch.long)
    out = model(x, time)
    taps = {}
    mine = G.unet_forward(sd, x, time, taps=taps)
    print(f"t={t}: oracle vs reference rel-L2 {G.rel_l2(mine, out):.3e}")
    fx[f"t{t}.out"] = out.numpy()
    if t == 999:
        for k in ("downs.0", "downs.3", "mid", "ups.3"):
            fx["tap." + k] = taps[k][:, ::8].numpy()
np.savez_compressed(os.path.join(OUT, "gaussian_unet_32x48.npz"), **fx)

# ---- schedule -----------------------------------------------------------------------------------------------------
diff = GaussianDiffusion(model, image_size=32, timesteps=1000, sampling_timesteps=4, loss_type='l1')
sch = {k: v.numpy() for k, v in diff.state_dict().items() if not k.startswith("model.")}
np.savez_compressed(os.path.join(OUT, "gaussian_schedule.npz"), **sch)

# ---- DDIM, 4 steps, eta = 0: the reference draws x_T with torch.randn under the global seed ---------------------------
torch.manual_seed(123)
ref_img = diff.sample(batch_size=2)[0]
torch.manual_seed(123)
init = torch.randn(2, 3, 32, 32)
trace = []
mine = G.ddim_sample(sd, init, 4, trace=trace)
print(f"ddim-4: oracle vs reference rel-L2 {G.rel_l2(mine, ref_img):.3e}")
np.savez_compressed(os.path.join(OUT, "gaussian_ddim4_32.npz"), init=init.numpy(), out=ref_img.numpy(),
                    **{f"step{i}.pred_noise": tr["pred_noise"].numpy() for i, tr in enumerate(trace)},
                    **{f"step{i}.x_start": tr["x_start"].numpy() for i, tr in enumerate(trace)})

# ---- ancestral, 6-step schedule (timesteps=6 so that the full loop is cheap) -------------------------------------------
diff6 = GaussianDiffusion(model, image_size=32, timesteps=6, loss_type='l1')
torch.manual_seed(321)
ref6 = diff6.sample(batch_size=2)[0]
torch.manual_seed(321)
init6 = torch.randn(2, 3, 32, 32)
noises = {}
for t in reversed(range(6)):                    # p_sample draws randn_like(x) for t > 0, in this order
    if t > 0:
        noises[t] = torch.randn(2, 3, 32, 32)
mine6 = G.p_sample_loop(sd, init6, lambda t: noises[t], timesteps=6)
print(f"ancestral-6: oracle vs reference rel-L2 {G.rel_l2(mine6, ref6):.3e}")
np.savez_compressed(os.path.join(OUT, "gaussian_ancestral6_32.npz"), init=init6.numpy(), out=ref6.numpy(),
                    **{f"noise{t}": v.numpy() for t, v in noises.items()})
print("written to", OUT)
