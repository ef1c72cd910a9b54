"""Generate tests/golden/unet_48x80.npz from the UNMODIFIED reference: one Unet.forward + a 2-step DDIM sample() on a geometry whose
levels do not tile evenly (48x80 -> 24x40 -> 12x20 -> 6x10; scan lengths 960 / 240 / 60 / 15), so that the kernels' fallback and tail
paths are held to the same gates as the power-of-two sizes.  TEST INFRASTRUCTURE.   Usage:  python -m oracle.gen_golden_ragged"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from founddiff_b200 import weights  # noqa: E402
from oracle import ref_shims  # noqa: E402
from oracle.gen_golden import npz, synth_slices  # noqa: E402


@torch.no_grad()
def main():
    torch.manual_seed(0)
    ns, model, diffusion = ref_shims.build_reference(sampling_timesteps=2, image_size=48)
    res = model.unet0.load_state_dict(weights.random_state_dict(10), strict=False)
    assert not res.unexpected_keys
    diffusion.init()
    B, H, W = 1, 48, 80
    ndct, ldct = synth_slices(B, H, W, seed=4880)
    x_input = ldct * 2 - 1
    x_t = x_input + 0.1 * torch.randn(B, 1, H, W, generator=torch.Generator().manual_seed(5))
    x_in = torch.cat((x_t, x_input), dim=1)
    time = (diffusion.alphas_cumsum[499] * 1000).expand(B)
    fx = dict(ldct=ldct, ndct=ndct, x_in=x_in, time=time, out=model(x_in, [time, time])[0])
    seed = 99
    torch.manual_seed(seed)
    fx["init_noise"] = torch.randn(B, 1, H, W)
    torch.manual_seed(seed)
    fx["outs"] = torch.stack(diffusion.sample([ldct.clone()], batch_size=B, last=False))
    npz("unet_48x80.npz", **fx)


if __name__ == "__main__":
    main()
