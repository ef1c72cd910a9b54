"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch, functional, state_dict-driven) of the reference's SECONDARY
sampling path: the lucidrains `Unet` + epsilon-prediction `GaussianDiffusion` of src/denoising_diffusion_pytorch.py
(SURVEY.md section 8, rows a18 / a19).  Pinned against the unmodified reference by oracle/gen_golden_gaussian.py
(tests/golden/gaussian_*.npz).  Nothing under founddiff_b200/ may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------------------------- schedule (:425-521)
def make_schedule(timesteps: int = 1000, beta_schedule: str = "cosine"):
    if beta_schedule == "cosine":                                      # cosine_beta_schedule :425-435
        steps = timesteps + 1
        x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
        ac = torch.cos(((x / timesteps) + 0.008) / 1.008 * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    else:                                                              # linear_beta_schedule :419-423
        scale = 1000 / timesteps
        betas = torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
    alphas = 1. - betas
    abar = torch.cumprod(alphas, dim=0)
    abar_prev = F.pad(abar[:-1], (1, 0), value=1.)
    pv = betas * (1. - abar_prev) / (1. - abar)
    f32 = lambda t: t.to(torch.float32)  # noqa: E731   (register_buffer casts float64 -> float32, :489)
    return dict(betas=f32(betas), alphas_cumprod=f32(abar), alphas_cumprod_prev=f32(abar_prev),
                sqrt_recip_alphas_cumprod=f32(torch.sqrt(1. / abar)), sqrt_recipm1_alphas_cumprod=f32(torch.sqrt(1. / abar - 1)),
                posterior_variance=f32(pv), posterior_log_variance_clipped=f32(torch.log(pv.clamp(min=1e-20))),
                posterior_mean_coef1=f32(betas * torch.sqrt(abar_prev) / (1. - abar)),
                posterior_mean_coef2=f32((1. - abar_prev) * torch.sqrt(alphas) / (1. - abar)))


# ---------------------------------------------------------------------------------------------- blocks
def ws_weight(w: Tensor, eps: float = 1e-5) -> Tensor:                 # WeightStandardizedConv2d :112-125
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + eps).rsqrt()


def channel_layernorm(x: Tensor, g: Tensor, eps: float = 1e-5) -> Tensor:   # LayerNorm :127-136
    var = torch.var(x, dim=1, unbiased=False, keepdim=True)
    mean = torch.mean(x, dim=1, keepdim=True)
    return (x - mean) * (var + eps).rsqrt() * g


def block(sd, p: str, x: Tensor, scale_shift=None) -> Tensor:           # Block :183-199
    x = F.conv2d(x, ws_weight(sd[p + ".proj.weight"]), sd[p + ".proj.bias"], padding=1)
    x = F.group_norm(x, 8, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    if scale_shift is not None:
        scale, shift = scale_shift
        x = x * (scale + 1) + shift
    return F.silu(x)


def resnet_block(sd, p: str, x: Tensor, t: Tensor) -> Tensor:           # ResnetBlock :201-225
    te = F.linear(F.silu(t), sd[p + ".mlp.1.weight"], sd[p + ".mlp.1.bias"])[:, :, None, None]
    h = block(sd, p + ".block1", x, te.chunk(2, dim=1))
    h = block(sd, p + ".block2", h)
    res = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"]) if (p + ".res_conv.weight") in sd else x
    return h + res


def linear_attention(sd, p: str, x: Tensor, heads: int = 4) -> Tensor:  # LinearAttention :227-255
    b, c, h, w = x.shape
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, -1, h * w) for t in qkv)
    q = q.softmax(dim=-2) * (q.shape[2] ** -0.5)
    k = k.softmax(dim=-1)
    v = v / (h * w)
    context = torch.einsum('bhdn,bhen->bhde', k, v)
    out = torch.einsum('bhde,bhdn->bhen', context, q).reshape(b, -1, h, w)
    out = F.conv2d(out, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return channel_layernorm(out, sd[p + ".to_out.1.g"])


def attention(sd, p: str, x: Tensor, heads: int = 4) -> Tensor:         # Attention :257-279
    b, c, h, w = x.shape
    qkv = F.conv2d(x, sd[p + ".to_qkv.weight"]).chunk(3, dim=1)
    q, k, v = (t.reshape(b, heads, -1, h * w) for t in qkv)
    q = q * (q.shape[2] ** -0.5)
    sim = torch.einsum('bhdi,bhdj->bhij', q, k)
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bhij,bhdj->bhid', attn, v)
    out = out.permute(0, 1, 3, 2).reshape(b, -1, h, w)
    return F.conv2d(out, sd[p + ".to_out.weight"], sd[p + ".to_out.bias"])


def time_embedding(sd, time: Tensor, dim: int) -> Tensor:               # SinusoidalPosEmb + time_mlp :150-162, 326-331
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = time.to(torch.float32)[:, None] * emb[None, :]
    emb = torch.cat((emb.sin(), emb.cos()), dim=-1)
    h = F.gelu(F.linear(emb, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return F.linear(h, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def unet_forward(sd, x: Tensor, time: Tensor, taps: Optional[dict] = None) -> Tensor:
    """Unet.forward :371-410 (self_condition=False).  x: (B, channels, H, W); time: (B,) timestep indices."""
    dim = sd["init_conv.weight"].shape[0]
    n = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("downs."))
    tap = (lambda k, v: taps.__setitem__(k, v.clone())) if taps is not None else (lambda k, v: None)
    x = F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3)
    r = x
    t = time_embedding(sd, time, dim)
    tap("t", t)
    hs: List[Tensor] = []
    for i in range(n):
        x = resnet_block(sd, f"downs.{i}.0", x, t)
        hs.append(x)
        x = resnet_block(sd, f"downs.{i}.1", x, t)
        x = linear_attention(sd, f"downs.{i}.2.fn.fn", channel_layernorm(x, sd[f"downs.{i}.2.fn.norm.g"])) + x
        hs.append(x)
        tap(f"downs.{i}", x)
        w = sd[f"downs.{i}.3.weight"]
        x = F.conv2d(x, w, sd[f"downs.{i}.3.bias"], stride=2, padding=1) if w.shape[-1] == 4 else \
            F.conv2d(x, w, sd[f"downs.{i}.3.bias"], padding=1)
    x = resnet_block(sd, "mid_block1", x, t)
    x = attention(sd, "mid_attn.fn.fn", channel_layernorm(x, sd["mid_attn.fn.norm.g"])) + x
    x = resnet_block(sd, "mid_block2", x, t)
    tap("mid", x)
    for i in range(n):
        x = resnet_block(sd, f"ups.{i}.0", torch.cat((x, hs.pop()), dim=1), t)
        x = resnet_block(sd, f"ups.{i}.1", torch.cat((x, hs.pop()), dim=1), t)
        x = linear_attention(sd, f"ups.{i}.2.fn.fn", channel_layernorm(x, sd[f"ups.{i}.2.fn.norm.g"])) + x
        tap(f"ups.{i}", x)
        if f"ups.{i}.3.1.weight" in sd:                                   # nn.Sequential(nn.Upsample, nn.Conv2d) :103-107
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = F.conv2d(x, sd[f"ups.{i}.3.1.weight"], sd[f"ups.{i}.3.1.bias"], padding=1)
        else:
            x = F.conv2d(x, sd[f"ups.{i}.3.weight"], sd[f"ups.{i}.3.bias"], padding=1)
    x = resnet_block(sd, "final_res_block", torch.cat((x, r), dim=1), t)
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])


# ---------------------------------------------------------------------------------------------- samplers
def _x0(sch, x, t, eps, clip):
    x0 = sch["sqrt_recip_alphas_cumprod"][t] * x - sch["sqrt_recipm1_alphas_cumprod"][t] * eps      # :523-527
    return x0.clamp(-1., 1.) if clip else x0


def p_sample_loop(sd, init: Tensor, step_noise: Callable[[int], Tensor], timesteps: int = 1000, beta_schedule: str = "cosine",
                  trace: Optional[list] = None) -> Tensor:
    """GaussianDiffusion.p_sample_loop / p_sample / p_mean_variance / q_posterior (:547-610), objective pred_noise.
    init: the initial x_T; step_noise(t): the noise of step t (t > 0).  Returns the image in [0, 1]."""
    sch = make_schedule(timesteps, beta_schedule)
    img = init
    B = img.shape[0]
    for t in reversed(range(timesteps)):
        eps = unet_forward(sd, img, torch.full((B,), t, dtype=torch.long))
        x0 = _x0(sch, img, t, eps, True)
        mean = sch["posterior_mean_coef1"][t] * x0 + sch["posterior_mean_coef2"][t] * img
        if trace is not None:
            trace.append(dict(t=t, x_t=img.clone(), pred_noise=eps.clone(), x_start=x0.clone()))
        img = mean + ((0.5 * sch["posterior_log_variance_clipped"][t]).exp() * step_noise(t) if t > 0 else 0.)
    return (img + 1) * 0.5


def ddim_times(timesteps: int, sampling_timesteps: int):
    times = torch.linspace(-1, timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_sample(sd, init: Tensor, sampling_timesteps: int, step_noise: Optional[Callable[[int], Tensor]] = None, eta: float = 0.,
                timesteps: int = 1000, beta_schedule: str = "cosine", trace: Optional[list] = None) -> Tensor:
    """GaussianDiffusion.ddim_sample (:612-646)."""
    sch = make_schedule(timesteps, beta_schedule)
    img = init
    B = img.shape[0]
    for t, tn in ddim_times(timesteps, sampling_timesteps):
        eps = unet_forward(sd, img, torch.full((B,), t, dtype=torch.long))
        x0 = _x0(sch, img, t, eps, True)
        if trace is not None:
            trace.append(dict(t=t, x_t=img.clone(), pred_noise=eps.clone(), x_start=x0.clone()))
        if tn < 0:
            img = x0
            continue
        a, an = sch["alphas_cumprod"][t], sch["alphas_cumprod"][tn]
        sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        c = (1 - an - sigma ** 2).sqrt()
        noise = step_noise(t) if (step_noise is not None and eta > 0) else 0.
        img = x0 * an.sqrt() + c * eps + sigma * noise
    return (img + 1) * 0.5


def rel_l2(a: Tensor, b: Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
