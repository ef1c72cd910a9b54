"""CPU: the secondary-path oracle (oracle/gaussian_oracle.py) against fixtures generated from the unmodified reference
(src/denoising_diffusion_pytorch.py) by oracle/gen_golden_gaussian.py, and host-side logic of founddiff_b200.gaussian."""
import numpy as np
import torch

from conftest import load_golden
from oracle import gaussian_oracle as G


def _sd():
    from founddiff_b200.gaussian import random_gaussian_state_dict
    return random_gaussian_state_dict(11)


def test_schedule_matches_reference_buffers():
    g = load_golden("gaussian_schedule.npz")
    sch = G.make_schedule(1000, "cosine")
    for k, v in sch.items():
        assert torch.allclose(v, g[k], rtol=0, atol=0), k
    from founddiff_b200.gaussian import make_schedule
    mine = make_schedule(1000, "cosine")
    for k in g:
        assert torch.equal(mine[k], g[k]), k


def test_unet_forward_matches_reference():
    g = load_golden("gaussian_unet_32x48.npz")
    sd = _sd()
    taps = {}
    for t in (999, 250):
        out = G.unet_forward(sd, g["x"], torch.full((2,), t, dtype=torch.long), taps=taps if t == 999 else None)
        assert G.rel_l2(out, g[f"t{t}.out"]) < 2e-5
    for k in ("downs.0", "downs.3", "mid", "ups.3"):
        assert G.rel_l2(taps[k][:, ::8], g["tap." + k]) < 2e-5


def test_ddim_and_ancestral_match_reference():
    sd = _sd()
    g = load_golden("gaussian_ddim4_32.npz")
    trace = []
    out = G.ddim_sample(sd, g["init"], 4, trace=trace)
    assert G.rel_l2(out, g["out"]) < 2e-5
    for i, tr in enumerate(trace):
        assert G.rel_l2(tr["pred_noise"], g[f"step{i}.pred_noise"]) < 2e-5
    a = load_golden("gaussian_ancestral6_32.npz")
    out = G.p_sample_loop(sd, a["init"], lambda t: a[f"noise{t}"], timesteps=6)
    assert G.rel_l2(out, a["out"]) < 2e-5


def test_schema_and_step_plan():
    from founddiff_b200.gaussian import GaussianDiffusion, Unet, gaussian_unet_schema
    keys = [k for k, _, _ in gaussian_unet_schema()]
    assert len(keys) == len(set(keys))
    m = Unet(dim=64, dim_mults=(1, 2, 4, 8))
    assert set(m.state_dict().keys()) == set(keys)
    d = GaussianDiffusion(m, image_size=32, timesteps=1000, sampling_timesteps=4, loss_type='l1')
    plan = d._plan()
    assert [t for t, _ in plan] == [t for t, _ in G.ddim_times(1000, 4)]
    assert plan[-1][1][2:6] == [1., 0., 0., 0.]                 # last pair: img = x_start
    d6 = GaussianDiffusion(m, image_size=32, timesteps=6, loss_type='l1')
    p6 = d6._plan()
    assert [t for t, _ in p6] == [5, 4, 3, 2, 1, 0] and p6[-1][1][5] == 0.
    try:
        d.sample(batch_size=1)
        assert False, "CPU sample() must fail loudly"
    except RuntimeError:
        pass
