"""Generate tests/golden/ddim10_32x48.npz from the UNMODIFIED reference: ResidualDiffusion.sample with sampling_timesteps = 10
(BASELINE config 2 quotes DDIM-2 and DDIM-10), last=False, one 32x48 slice.  TEST INFRASTRUCTURE.
Usage:  python -m oracle.gen_golden_ddim10"""
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
    ns, model, diffusion = ref_shims.build_reference(sampling_timesteps=10, image_size=32)
    res = model.unet0.load_state_dict(weights.random_state_dict(10), strict=False)
    assert not res.unexpected_keys
    diffusion.init()
    ndct, ldct = synth_slices(1, 32, 48, seed=1010)
    seed = 1234
    torch.manual_seed(seed)
    init_noise = torch.randn(1, 1, 32, 48)
    torch.manual_seed(seed)
    outs = diffusion.sample([ldct.clone()], batch_size=1, last=False)
    npz("ddim10_32x48.npz", ldct=ldct, ndct=ndct, init_noise=init_noise, outs=torch.stack(outs))


if __name__ == "__main__":
    main()
