"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) in this container.

Usage:  python -m oracle.gen_golden            (needs /root/reference; ~2 min on 8 cores)

TEST INFRASTRUCTURE.  The reference is imported under oracle/ref_shims.py; weights come from
founddiff_b200.weights.random_state_dict(seed) — the same function the tests call to rebuild the identical
weights on the GPU box (the 67 M live parameters are far too large to commit).  Loading that state dict into
the reference modules (`strict=False`, asserting zero unexpected keys) is also the proof that our key schema
matches the reference's.

Every fixture records inputs, injected noise (replayed from the global torch RNG the reference draws from) and
the reference's outputs.  Fixtures:

  schedule.npz        the 12 schedule buffers after ResidualDiffusion.__init__ and after .init()
  unet_64x96.npz      Unet.forward (B=2, 64x96 — non-square on purpose) at t=999 and t=499 + stage taps
  blocks.npz          Mamba_block / SS2D / TransposedAttention / ResnetBlock in isolation (C=64, N=4, 16x24)
  ddim_64x96.npz      ResidualDiffusion.sample, DDIM S=2 and S=5, last=False (all intermediates)
  ancestral_32.npz    p_sample_loop with num_timesteps overridden to 12 (t=11..0), 32x32, last=False
  scan.npz            selective scan: C restatement vs an independent fp64 sequential evaluation of the
                      published recurrence (the published KERNEL's outputs: scan_vllm.npz, gen_golden_scan_vllm.py)
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from founddiff_b200 import weights  # noqa: E402
from oracle import ref_shims, scan_cpu  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
WEIGHT_SEED = 10
TAP_CH_STRIDE = 8


def synth_slices(B, H, W, seed=1234, sigma=0.05):
    """SURVEY §8d synthetic input: box-filtered uniform noise as 'NDCT', + Gaussian noise as 'LDCT', in [0,1]."""
    g = torch.Generator().manual_seed(seed)
    ndct = torch.rand(B, 1, H, W, generator=g)
    ndct = torch.nn.functional.avg_pool2d(torch.nn.functional.pad(ndct, (2, 2, 2, 2), mode="reflect"), 5, stride=1)
    ldct = (ndct + sigma * torch.randn(B, 1, H, W, generator=g)).clamp(0, 1)
    return ndct, ldct


def npz(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"wrote {path}  ({os.path.getsize(path) / 1024:.0f} KiB)")


def scan_fp64_sequential(u, delta, A, B, C, D, bias):
    """Independent evaluation of the published S6 recurrence in float64 (pure torch, sequential over L)."""
    u, delta, A, B, C, D, bias = [t.double() for t in (u, delta, A, B, C, D, bias)]
    b, KD, L = u.shape
    G, N = B.shape[1], B.shape[2]
    dt = torch.nn.functional.softplus(delta + bias[None, :, None], threshold=20.0)
    Bx = B.repeat_interleave(KD // G, dim=1)      # (b, KD, N, L)
    Cx = C.repeat_interleave(KD // G, dim=1)
    h = torch.zeros(b, KD, N, dtype=torch.float64)
    ys = []
    for l in range(L):
        h = torch.exp(dt[:, :, l, None] * A[None]) * h + dt[:, :, l, None] * Bx[..., l] * u[:, :, l, None]
        ys.append((h * Cx[..., l]).sum(-1) + D[None] * u[:, :, l])
    return torch.stack(ys, dim=-1)


@torch.no_grad()
def main():
    torch.manual_seed(0)
    sd = weights.random_state_dict(WEIGHT_SEED)
    ns, model, diffusion = ref_shims.build_reference(sampling_timesteps=2)
    unet = model.unet0
    res = unet.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    dead = [k for k in res.missing_keys if not (k.startswith("clip_model.") or k.startswith("dose_encoder.clip_model.")
                                                or k.startswith("dose_encoder.prompt_learner."))]
    assert not dead, dead
    DADiff = ns.DADiff

    # ---- schedule -------------------------------------------------------------------------------------
    names = ["alphas", "alphas_cumsum", "one_minus_alphas_cumsum", "betas2", "betas", "betas2_cumsum", "betas_cumsum",
             "posterior_mean_coef1", "posterior_mean_coef2", "posterior_mean_coef3", "posterior_variance",
             "posterior_log_variance_clipped"]
    sch = {f"ctor.{n}": getattr(diffusion, n).clone() for n in names}
    diffusion.init()                                   # what Trainer.test() does first (src/DADiff.py:1818)
    sch.update({f"init.{n}": getattr(diffusion, n).clone() for n in names})
    npz("schedule.npz", **sch)

    # ---- Unet.forward + taps --------------------------------------------------------------------------
    B, H, W = 2, 64, 96
    ndct, ldct = synth_slices(B, H, W)
    x_input = ldct * 2 - 1
    g = torch.Generator().manual_seed(77)
    x_t = x_input + 0.1 * torch.randn(B, 1, H, W, generator=g)
    x_in = torch.cat((x_t, x_input), dim=1)
    taps = {}
    hooks = []

    def hook(name):
        def f(_m, _i, o):
            taps[name] = (o[1] if isinstance(o, tuple) and name == "dose" else o)
        return f
    hooks.append(unet.init_conv.register_forward_hook(hook("init_conv")))
    for i in range(4):
        hooks.append(unet.downs[i][1].register_forward_hook(hook(f"downs.{i}.mamba")))
        hooks.append(unet.downs[i][0].register_forward_hook(hook(f"downs.{i}.res")))
        hooks.append(unet.downs[i][2].register_forward_hook(hook(f"downs.{i}.down")))
        hooks.append(unet.ups[i][0].register_forward_hook(hook(f"ups.{i}.res")))
        hooks.append(unet.ups[i][1].register_forward_hook(hook(f"ups.{i}.mamba")))
        hooks.append(unet.ups[i][2].register_forward_hook(hook(f"ups.{i}.up")))
    hooks.append(unet.mid_attn.register_forward_hook(hook("mid")))
    hooks.append(unet.final_res_block.register_forward_hook(hook("final_res")))
    emb = {}
    hooks.append(unet.dose_encoder.register_forward_hook(lambda m, i, o: emb.update(dose_emb=o[1], ctx_emb=o[2])))
    fx = dict(x_in=x_in, ndct=ndct, ldct=ldct)
    for t in (999, 499):
        time = (diffusion.alphas_cumsum[t] * 1000).expand(B)
        out = model(x_in, [time, time])[0]
        fx[f"t{t}.time"] = time
        fx[f"t{t}.out"] = out
        if t == 999:
            for k, v in taps.items():
                fx[f"t999.tap.{k}"] = v[:, ::TAP_CH_STRIDE].clone()   # every 8th channel keeps the file small
            fx["dose_emb"], fx["ctx_emb"] = emb["dose_emb"].clone(), emb["ctx_emb"].clone()
    for h in hooks:
        h.remove()
    npz("unet_64x96.npz", **fx)

    # ---- isolated blocks ------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(5)
    Bb, C, Hb, Wb = 2, 64, 16, 24
    xb = torch.randn(Bb, C, Hb, Wb, generator=g)
    cb = torch.nn.functional.normalize(torch.randn(Bb, 1, 256, generator=g), dim=-1)
    tb = torch.randn(Bb, 256, generator=g)
    blk = unet.downs[0][1]
    fb = dict(x=xb, c=cb, t=tb)
    fb["mamba_block.downs.0.1"] = blk(xb, cb, tb)
    fb["ss2d.downs.0.1"] = blk.mamba(xb.permute(0, 2, 3, 1).contiguous(), cb)
    fb["tattn.downs.0.1"] = blk.attn_blk(xb)
    fb["resnet.downs.0.0"] = unet.downs[0][0](xb)
    x128 = torch.randn(Bb, 128, Hb, Wb, generator=g)
    fb["x128"] = x128
    fb["resnet.final_res_block"] = unet.final_res_block(x128)                  # 128 -> 64 with res_conv
    fb["mamba_block.downs.2.1"] = unet.downs[2][1](x128, cb, tb)              # C=128, N=16, R=8
    xs = DADiff.SS2D  # noqa: F841  (documentational: class under test)
    es = ns.emamba2.EfficientScan.apply(xb, 2)
    fb["efficient_scan"] = es
    fb["efficient_merge"] = ns.emamba2.EfficientMerge.apply(es, Hb, Wb, 2)
    npz("blocks.npz", **fb)

    # ---- DDIM sample() --------------------------------------------------------------------------------
    fd = dict(ldct=ldct, ndct=ndct)
    for S in (2, 5):
        diffusion.sampling_timesteps = S
        diffusion.is_ddim_sampling = True
        seed = 4321 + S
        torch.manual_seed(seed)
        noise = torch.randn(B, 1, H, W)               # replay of the single randn(shape) at src/DADiff.py:1295
        torch.manual_seed(seed)
        outs = diffusion.sample([ldct.clone()], batch_size=B, last=False)
        fd[f"S{S}.init_noise"] = noise
        fd[f"S{S}.outs"] = torch.stack(outs)
        torch.manual_seed(seed)
        fd[f"S{S}.last"] = torch.stack(diffusion.sample([ldct.clone()], batch_size=B, last=True))
    npz("ddim_64x96.npz", **fd)

    # ---- ancestral p_sample_loop, 12 steps -------------------------------------------------------------
    Ba, Ha = 2, 32
    ndct_a, ldct_a = synth_slices(Ba, Ha, Ha, seed=99)
    T = 12
    diffusion.sampling_timesteps = 1000
    diffusion.is_ddim_sampling = False
    diffusion.num_timesteps = T                        # loop t = T-1 .. 0 over the first T schedule entries
    seed = 2468
    torch.manual_seed(seed)
    init_noise = torch.randn(Ba, 1, Ha, Ha)
    step_noise = torch.stack([torch.randn(Ba, 1, Ha, Ha) for _ in range(T - 1)])   # drawn at t = T-1 .. 1
    torch.manual_seed(seed)
    outs = diffusion.sample([ldct_a.clone()], batch_size=Ba, last=False)
    diffusion.num_timesteps = 1000
    npz("ancestral_32.npz", ldct=ldct_a, ndct=ndct_a, init_noise=init_noise, step_noise=step_noise,
        outs=torch.stack(outs), num_timesteps=T)

    # ---- selective scan --------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    fs = {}
    for tag, (b, K, Dk, N, L) in dict(a=(2, 4, 8, 4, 300), b=(1, 4, 16, 16, 65), c=(2, 1, 4, 32, 1)).items():
        u = torch.randn(b, K * Dk, L, generator=g)
        delta = torch.randn(b, K * Dk, L, generator=g) * 2
        delta[0, 0, : min(5, L)] = 25.0                # exercise the softplus threshold branch
        A = -torch.exp(torch.randn(K * Dk, N, generator=g) * 0.5)
        Bm = torch.randn(b, K, N, L, generator=g)
        Cm = torch.randn(b, K, N, L, generator=g)
        D = torch.randn(K * Dk, generator=g)
        bias = torch.randn(K * Dk, generator=g)
        y64 = scan_fp64_sequential(u, delta, A, Bm, Cm, D, bias)
        y32 = scan_cpu.selective_scan_fwd(u, delta, A, Bm, Cm, D, bias, True)
        err = (y32.double() - y64).abs().max().item() / y64.abs().max().item()
        print(f"scan[{tag}] C-fp32 vs fp64 sequential: max rel err {err:.2e}")
        assert err < 1e-5
        fs.update({f"{tag}.u": u, f"{tag}.delta": delta, f"{tag}.A": A, f"{tag}.B": Bm, f"{tag}.C": Cm, f"{tag}.D": D,
                   f"{tag}.bias": bias, f"{tag}.y64": y64})
    npz("scan.npz", **fs)


if __name__ == "__main__":
    main()
