"""Generate tests/golden/objectives.npz by running the UNMODIFIED reference (/root/reference) in this container:
the objectives other than the shipped 'pred_res' (src/DADiff.py:775-836, 1168-1207, 1399-1482).

Usage:  python -m oracle.gen_golden_objectives            (needs /root/reference; ~3 min on 8 cores)

TEST INFRASTRUCTURE.  Weights: unet0 = weights.random_state_dict(10), unet1 = weights.random_state_dict(11) — the
tests rebuild the same tensors.  For every configuration the fixture holds a 3-step DDIM chain and a 4-step ancestral
chain (num_timesteps overridden, as in gen_golden.py) with all intermediates, the per-step model_predictions taps of
the DDIM chain, and one p_losses evaluation with fixed (t, noise).

  tag      num_unet  objective        test_res_or_noise
  rn       2         pred_res_noise   res_noise          (train.py:75-77)
  rn_res   2         pred_res_noise   res
  rn_noise 2         pred_res_noise   noise
  x0n      2         pred_x0_noise    res_noise
  noise    1         pred_noise       -
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from founddiff_b200 import weights  # noqa: E402
from oracle import ref_shims  # noqa: E402
from oracle.gen_golden import npz, synth_slices  # noqa: E402

CONFIGS = dict(rn=(2, "pred_res_noise", "res_noise"), rn_res=(2, "pred_res_noise", "res"),
               rn_noise=(2, "pred_res_noise", "noise"), x0n=(2, "pred_x0_noise", "res_noise"),
               noise=(1, "pred_noise", "None"))
B, H, W = 2, 32, 48
S_DDIM, T_ANC = 3, 4


def load(unet, seed):
    res = unet.load_state_dict(weights.random_state_dict(seed), strict=False)
    assert not res.unexpected_keys, res.unexpected_keys


@torch.no_grad()
def main():
    ndct, ldct = synth_slices(B, H, W, seed=321)
    fx = dict(ldct=ldct, ndct=ndct)
    g = torch.Generator().manual_seed(2024)
    t_loss = torch.tensor([700, 130])
    noise_loss = torch.randn(B, 1, H, W, generator=g)
    fx["loss.t"], fx["loss.noise"] = t_loss, noise_loss
    for tag, (num_unet, objective, trn) in CONFIGS.items():
        torch.manual_seed(0)
        ns, model, diffusion = ref_shims.build_reference(sampling_timesteps=S_DDIM, image_size=H, num_unet=num_unet,
                                                         objective=objective, test_res_or_noise=trn, loss_type="l1")
        load(model.unet0, 10)
        if num_unet == 2:
            load(model.unet1, 11)
        diffusion.init()
        # ---- DDIM chain with per-step taps (model_predictions wrapped, not modified) --------------------------
        taps = []
        orig = diffusion.model_predictions

        def tapped(*a, **k):
            r = orig(*a, **k)
            taps.append(r)
            return r
        diffusion.model_predictions = tapped
        seed = 777
        torch.manual_seed(seed)
        init_noise = torch.randn(B, 1, H, W)
        torch.manual_seed(seed)
        outs = diffusion.sample([ldct.clone()], batch_size=B, last=False)
        fx[f"{tag}.ddim.init_noise"] = init_noise
        fx[f"{tag}.ddim.outs"] = torch.stack(outs)
        for i, r in enumerate(taps):
            fx[f"{tag}.ddim.step{i}.pred_res"], fx[f"{tag}.ddim.step{i}.pred_noise"] = r.pred_res, r.pred_noise
            fx[f"{tag}.ddim.step{i}.x_start"] = r.pred_x_start
        diffusion.model_predictions = orig
        # ---- ancestral chain, T_ANC steps ------------------------------------------------------------------------
        diffusion.sampling_timesteps, diffusion.is_ddim_sampling, diffusion.num_timesteps = 1000, False, T_ANC
        seed = 888
        torch.manual_seed(seed)
        init_noise = torch.randn(B, 1, H, W)
        step_noise = torch.stack([torch.randn(B, 1, H, W) for _ in range(T_ANC - 1)])
        torch.manual_seed(seed)
        outs = diffusion.sample([ldct.clone()], batch_size=B, last=False)
        diffusion.num_timesteps = 1000
        fx[f"{tag}.anc.init_noise"], fx[f"{tag}.anc.step_noise"] = init_noise, step_noise
        fx[f"{tag}.anc.outs"] = torch.stack(outs)
        # ---- p_losses (forward only) -------------------------------------------------------------------------------
        if trn in ("res_noise", "None"):
            x_start, x_input = ndct * 2 - 1, ldct * 2 - 1
            losses = diffusion.p_losses([x_start, x_input], t_loss, noise=noise_loss)
            fx[f"{tag}.loss"] = torch.stack([l.detach() for l in losses])
            print(tag, "losses", [float(l) for l in losses])
    npz("objectives.npz", **fx)


if __name__ == "__main__":
    main()
