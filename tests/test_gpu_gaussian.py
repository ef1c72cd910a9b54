"""GPU: the secondary sampling path (founddiff_b200.gaussian) against fixtures generated from the unmodified reference
(src/denoising_diffusion_pytorch.py): Unet forward, DDIM and ancestral sampling with injected noise.
Gate: rel-L2 <= 1e-2 per step for 16-bit storage (north star), final image within 0.05 dB PSNR."""
import math

import pytest
import torch

from conftest import load_golden
from oracle import gaussian_oracle as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gmodel():
    from founddiff_b200.gaussian import Unet
    return Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda()          # seed 11 == the fixture weights


def rel(a, b):
    return G.rel_l2(a.detach().float().cpu(), b)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_unet_forward_vs_reference(gmodel, dt):
    g = load_golden("gaussian_unet_32x48.npz")
    gmodel.compute_dtype = dt
    for t in (999, 250):
        out = gmodel(g["x"].cuda(), torch.full((2,), t, device="cuda"))
        r = rel(out, g[f"t{t}.out"])
        assert r < (5e-3 if dt == torch.float16 else 2e-2), (dt, t, r)


def psnr(a, b):
    return 10 * math.log10(1.0 / float(((a - b) ** 2).mean()))


def _sd():
    from founddiff_b200.gaussian import random_gaussian_state_dict
    return random_gaussian_state_dict(11)


def test_ddim_sampling_vs_reference(gmodel):
    """Per-step gate with the reference's own x_t fed to the model (the oracle reproduces the reference bit-exactly, so its
    trace supplies x_t); the free-running chain is compared at the end.  At t = 999 the cosine schedule has
    sqrt(1/abar) ~ 2e4, so x_start = sr*x_t - srm1*eps amplifies ANY eps rounding before the clamp: free-running states
    decorrelate at the unsaturated pixels in the reference's own arithmetic too (fp32 vs fp64), hence the looser bound."""
    from founddiff_b200.gaussian import GaussianDiffusion
    g = load_golden("gaussian_ddim4_32.npz")
    gmodel.compute_dtype = torch.float16
    ref_trace = []
    ref = G.ddim_sample(_sd(), g["init"], 4, trace=ref_trace)
    assert G.rel_l2(ref, g["out"]) < 2e-5
    for i, tr in enumerate(ref_trace):
        assert G.rel_l2(tr["pred_noise"], g[f"step{i}.pred_noise"]) < 2e-5
        out = gmodel(tr["x_t"].cuda(), torch.full((2,), tr["t"], device="cuda"))
        assert rel(out, tr["pred_noise"]) < 1e-2, (i, rel(out, tr["pred_noise"]))
    d = GaussianDiffusion(gmodel, image_size=32, timesteps=1000, sampling_timesteps=4, loss_type='l1').cuda()
    trace = []
    out = d.sample(batch_size=2, noise={"init": g["init"]}, trace=trace)[0]
    assert rel(trace[0]["pred_noise"], g["step0.pred_noise"]) < 1e-2
    assert out.shape == (2, 3, 32, 32) and rel(out, g["out"]) < 0.15


def test_sampler_update_matches_reference_formulas(gmodel):
    """fd_ddpm_update driven by GaussianDiffusion._plan == the oracle's p_sample / DDIM algebra on the oracle's own eps
    (isolates the sampler from the Unet's 16-bit error)."""
    from founddiff_b200 import ops
    from founddiff_b200.gaussian import GaussianDiffusion
    for kw, oracle_fn in ((dict(timesteps=1000, sampling_timesteps=4), "ddim"), (dict(timesteps=6), "anc")):
        d = GaussianDiffusion(gmodel, image_size=32, loss_type='l1', **kw).cuda()
        sch = G.make_schedule(kw["timesteps"])
        gen = torch.Generator().manual_seed(7)
        plan = d._plan()
        for idx, (t, c) in enumerate(plan):
            x, eps, nz = (torch.randn(2, 3072, generator=gen) for _ in range(3))
            x0 = (sch["sqrt_recip_alphas_cumprod"][t] * x - sch["sqrt_recipm1_alphas_cumprod"][t] * eps).clamp(-1, 1)
            if oracle_fn == "anc":
                want = sch["posterior_mean_coef1"][t] * x0 + sch["posterior_mean_coef2"][t] * x
                if t > 0:
                    want = want + (0.5 * sch["posterior_log_variance_clipped"][t]).exp() * nz
            else:
                tn = plan[idx + 1][0] if idx + 1 < len(plan) else -1
                want = x0 if tn < 0 else x0 * sch["alphas_cumprod"][tn].sqrt() + (1 - sch["alphas_cumprod"][tn]).sqrt() * eps
            out = torch.empty(2, 3072, device="cuda")
            ops.ddpm_update(x.cuda(), eps.cuda(), nz.cuda() if c[5] != 0 else None, torch.tensor(c, dtype=torch.float32).cuda(), out)
            assert rel(out, want) < 1e-5, (oracle_fn, t)


def test_ancestral_sampling_vs_reference(gmodel):
    from founddiff_b200.gaussian import GaussianDiffusion
    a = load_golden("gaussian_ancestral6_32.npz")
    gmodel.compute_dtype = torch.float16
    ref_trace = []
    ref = G.p_sample_loop(_sd(), a["init"], lambda t: a[f"noise{t}"], timesteps=6, trace=ref_trace)
    assert G.rel_l2(ref, a["out"]) < 2e-5
    for tr in ref_trace:
        out = gmodel(tr["x_t"].cuda(), torch.full((2,), tr["t"], device="cuda"))
        assert rel(out, tr["pred_noise"]) < 1e-2, (tr["t"], rel(out, tr["pred_noise"]))
    d = GaussianDiffusion(gmodel, image_size=32, timesteps=6, loss_type='l1').cuda()
    out = d.sample(batch_size=2, noise={"init": a["init"], "steps": lambda t: a[f"noise{t}"]})[0]
    assert rel(out, a["out"]) < 0.15


def test_default_noise_shapes_and_graph_mode(gmodel):
    from founddiff_b200.gaussian import GaussianDiffusion
    gmodel.compute_dtype = torch.float16
    d = GaussianDiffusion(gmodel, image_size=64, timesteps=1000, sampling_timesteps=2, loss_type='l1').cuda()
    out = d.sample(batch_size=3)
    assert isinstance(out, list) and out[0].shape == (3, 3, 64, 64) and torch.isfinite(out[0]).all()
    init = torch.randn(3, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    eager = d.sample(batch_size=3, noise={"init": init})[0]
    d.use_cuda_graph = True
    graphed = d.sample(batch_size=3, noise={"init": init})[0]
    graphed2 = d.sample(batch_size=3, noise={"init": init})[0]
    assert rel(graphed, eager.float().cpu()) < 0.15 and rel(graphed2, graphed.float().cpu()) < 0.15


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_tensor_core_sizes_vs_oracle(gmodel, dt):
    """128 x 128: every 3x3 / 4x4s2 / up2 / 1x1 convolution of the Unet takes the tcgen05 path, linear attention runs its
    per-sample GEMM on it, the bottleneck attention sees 256 tokens; compared with the CPU oracle on a fresh input."""
    from founddiff_b200 import ops
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 3, 128, 128, generator=g)
    time = torch.full((1,), 400, dtype=torch.long)
    ref = G.unet_forward(_sd(), x, time)
    gmodel.compute_dtype = dt
    n0 = ops.LAUNCHES
    out = gmodel(x.cuda(), time.cuda())
    assert ops.LAUNCHES > n0
    eng = gmodel.engine(1, 128, 128, torch.device("cuda"))
    r = rel(out, ref)
    assert r < (5e-3 if dt == torch.float16 else 2e-2), (dt, r)
    assert eng.prefer_tc
