"""CPU oracle: a functional fp32 restatement of FoundDiff's reverse-diffusion sampling hot path.

TEST INFRASTRUCTURE — only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` /
`--impl reference` legs may import this.  The product path (`founddiff_b200/`) never does.

It restates, as plain functions over a `state_dict` with the reference's key names, what the reference computes
with nn.Modules.  Every function cites the reference lines it follows.  It is PINNED against outputs of the
reference itself: `oracle/gen_golden.py` imports the unmodified reference (under `oracle/ref_shims.py`) in the
build container and stores small input/output fixtures in `tests/golden/`; `tests/test_oracle_golden.py` checks
this file against them.  The one piece with no reference source is the selective scan (third-party CUDA
extension, un-vendored, no version named; see oracle/selective_scan_ref.c).  It is pinned against a BUILD OF THE
PUBLISHED KERNEL: vLLM 0.22 (in the image) ships the state-spaces/mamba `selective_scan_fwd` CUDA kernel;
`oracle/gen_golden_scan_vllm.py` ran it on a B200 and `tests/golden/scan_vllm.npz` holds its outputs
(C restatement vs those: <= 2e-7 rel-L2), next to an independent fp64 evaluation of the recurrence.

The reference's floating-point arithmetic is fp32 throughout (train.py:141 amp=False), so is this.
`rt(tensor, kind)` ("round-trip") is an optional hook applied wherever the CUDA path stores an activation to
HBM (kind "trunk" = the residual/skip stream, "inner" = block-internal tensors); tests use it to model bf16
storage rounding.  Default: identity.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from . import scan_cpu

Tensor = torch.Tensor
_ID = lambda t, kind="inner": t  # noqa: E731


# ----------------------------------------------------------------------------------------------------------
# Schedule  (src/DADiff.py:946-1027 constructor; :1033-1118 `init()` — the variant Trainer.test() uses, :1818)
# ----------------------------------------------------------------------------------------------------------
def make_schedule(timesteps: int = 1000, variant: str = "init") -> Dict[str, Tensor]:
    betas = torch.linspace(0.0001, 0.02, timesteps, dtype=torch.float32)          # :952-953 / :1042-1043
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    alphas_cumsum = 1 - alphas_cumprod ** 0.5                                       # :969
    betas2_cumsum = 1 - alphas_cumprod                                              # :970
    alphas_cumsum_prev = F.pad(alphas_cumsum[:-1], (1, 0), value=1.)
    betas2_cumsum_prev = F.pad(betas2_cumsum[:-1], (1, 0), value=1.)
    alphas = alphas_cumsum - alphas_cumsum_prev
    betas2 = betas2_cumsum - betas2_cumsum_prev
    if variant == "init":                                                           # :1065, 1067
        alphas[0] = alphas[1]
        betas2[0] = betas2[1]
    elif variant == "ctor":                                                         # :975, 977
        alphas[0] = 0
        betas2[0] = 0
    else:
        raise ValueError(variant)
    betas_cumsum = torch.sqrt(betas2_cumsum)
    posterior_variance = betas2 * betas2_cumsum_prev / betas2_cumsum
    posterior_variance[0] = 0
    s = dict(
        alphas=alphas, alphas_cumsum=alphas_cumsum, one_minus_alphas_cumsum=1 - alphas_cumsum,
        betas2=betas2, betas=torch.sqrt(betas2), betas2_cumsum=betas2_cumsum, betas_cumsum=betas_cumsum,
        posterior_mean_coef1=betas2_cumsum_prev / betas2_cumsum,
        posterior_mean_coef2=(betas2 * alphas_cumsum_prev - betas2_cumsum_prev * alphas) / betas2_cumsum,
        posterior_mean_coef3=betas2 / betas2_cumsum,
        posterior_variance=posterior_variance,
        posterior_log_variance_clipped=torch.log(posterior_variance.clamp(min=1e-20)),
    )
    s["posterior_mean_coef1"][0] = 0                                                # :1024-1027 / :1115-1118
    s["posterior_mean_coef2"][0] = 0
    s["posterior_mean_coef3"][0] = 1
    s["one_minus_alphas_cumsum"][-1] = 1e-6
    return {k: v.to(torch.float32) for k, v in s.items()}


# ----------------------------------------------------------------------------------------------------------
# DA-CLIP conditioning  (src/DACLIP.py:1189-1221 CLIPIQA.forward; :329-349 ModifiedResNet; :226-259 attnpool)
# ----------------------------------------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def _bottleneck(sd, p, x, stride):                                                  # src/DACLIP.py:168-211
    out = F.relu(_bn(sd, p + "bn1", F.conv2d(x, sd[p + "conv1.weight"])))
    out = F.relu(_bn(sd, p + "bn2", F.conv2d(out, sd[p + "conv2.weight"], padding=1)))
    if stride > 1:
        out = F.avg_pool2d(out, stride)
    out = _bn(sd, p + "bn3", F.conv2d(out, sd[p + "conv3.weight"]))
    if (p + "downsample.0.weight") in sd:
        idt = F.avg_pool2d(x, stride) if stride > 1 else x
        idt = _bn(sd, p + "downsample.1", F.conv2d(idt, sd[p + "downsample.0.weight"]))
    else:
        idt = x
    return F.relu(out + idt)


def daclip_embed(sd: Dict[str, Tensor], x_input: Tensor, layers: Sequence[int] = (3, 4, 6, 3), heads: int = 32):
    """x_input: (B,1,H,W) in [-1,1] (channel 1 of the Unet input, src/DADiff.py:692). Returns
    (dose_embedding (B,1024) L2-normalised, context_embedding (B,256) L2-normalised)."""
    v = "dose_encoder.clip_model.visual."
    x = x_input.repeat(1, 3, 1, 1)                                                  # src/DADiff.py:692
    x = F.relu(_bn(sd, v + "bn1", F.conv2d(x, sd[v + "conv1.weight"], stride=2, padding=1)))   # stem :331-335
    x = F.relu(_bn(sd, v + "bn2", F.conv2d(x, sd[v + "conv2.weight"], padding=1)))
    x = F.relu(_bn(sd, v + "bn3", F.conv2d(x, sd[v + "conv3.weight"], padding=1)))
    x = F.avg_pool2d(x, 2)
    for li, blocks in enumerate(layers):
        for bi in range(blocks):
            x = _bottleneck(sd, f"{v}layer{li + 1}.{bi}.", x, 2 if (li > 0 and bi == 0) else 1)
    # AttentionPool2d without positional embedding (pos_embedding=False, src/DACLIP.py:1203, 226-259)
    a = v + "attnpool."
    B, C, H, W = x.shape
    tok = x.reshape(B, C, H * W).permute(2, 0, 1)                                   # (HW, B, C)
    tok = torch.cat([tok.mean(dim=0, keepdim=True), tok], dim=0)                    # (HW+1, B, C)
    q = F.linear(tok[:1], sd[a + "q_proj.weight"], sd[a + "q_proj.bias"])           # only the pooled query is used (x[0])
    k = F.linear(tok, sd[a + "k_proj.weight"], sd[a + "k_proj.bias"])
    vv = F.linear(tok, sd[a + "v_proj.weight"], sd[a + "v_proj.bias"])
    hd = C // heads
    q = q.reshape(1, B, heads, hd) * (hd ** -0.5)
    k = k.reshape(-1, B, heads, hd)
    vv = vv.reshape(-1, B, heads, hd)
    att = torch.einsum("qbhd,kbhd->bhqk", q, k).softmax(dim=-1)
    o = torch.einsum("bhqk,kbhd->qbhd", att, vv).reshape(1, B, C)
    feat = F.linear(o, sd[a + "c_proj.weight"], sd[a + "c_proj.bias"])[0]            # (B, 1024)
    h1 = F.linear(F.relu(F.linear(feat, sd["dose_encoder.head1.0.weight"], sd["dose_encoder.head1.0.bias"])),
                  sd["dose_encoder.head1.2.weight"], sd["dose_encoder.head1.2.bias"])
    h2 = F.linear(F.relu(F.linear(feat, sd["dose_encoder.head2.0.weight"], sd["dose_encoder.head2.0.bias"])),
                  sd["dose_encoder.head2.2.weight"], sd["dose_encoder.head2.2.bias"])
    dose = h1 / h1.norm(dim=-1, keepdim=True)                                       # src/DACLIP.py:1210
    ctx = F.normalize(h2, dim=1)                                                    # src/DACLIP.py:1207
    return dose, ctx


# ----------------------------------------------------------------------------------------------------------
# Unet blocks
# ----------------------------------------------------------------------------------------------------------
def ws_weight(w: Tensor, eps: float = 1e-5) -> Tensor:
    """Weight standardisation (src/DADiff.py:145-152), fp32 eps."""
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + eps).rsqrt()


def resnet_block(sd, p: str, x: Tensor, groups: int = 8, rt: Callable = _ID) -> Tensor:
    """SiLU(GroupNorm8(WSConv3x3(x))) + res_conv(x)   (src/DADiff.py:213-229, 397-430)."""
    h = F.conv2d(x, ws_weight(sd[p + ".block1.proj.weight"]), sd[p + ".block1.proj.bias"], padding=1)
    h = rt(h, "conv_out")
    h = F.silu(F.group_norm(h, groups, sd[p + ".block1.norm.weight"], sd[p + ".block1.norm.bias"], eps=1e-5))
    if (p + ".res_conv.weight") in sd:
        skip = F.conv2d(x, sd[p + ".res_conv.weight"], sd[p + ".res_conv.bias"])
    else:
        skip = x
    return rt(h + skip, "trunk")


def efficient_scan(x: Tensor) -> Tensor:
    """(B,D,H,W) -> (B,4,D,H/2*W/2): four stride-2 sub-grids; 0,2 row-major, 1,3 column-major
    (src/emamba2.py:186-213).  H, W even."""
    B, D, H, W = x.shape
    xs = x.new_empty(B, 4, D, (H // 2) * (W // 2))
    xs[:, 0] = x[:, :, 0::2, 0::2].reshape(B, D, -1)
    xs[:, 1] = x[:, :, 1::2, 0::2].transpose(2, 3).reshape(B, D, -1)
    xs[:, 2] = x[:, :, 0::2, 1::2].reshape(B, D, -1)
    xs[:, 3] = x[:, :, 1::2, 1::2].transpose(2, 3).reshape(B, D, -1)
    return xs


def efficient_merge(ys: Tensor, H: int, W: int) -> Tensor:
    """(B,4,D,L) -> (B,D,H,W), inverse of efficient_scan (src/emamba2.py:238-262)."""
    B, K, D, L = ys.shape
    h2, w2 = H // 2, W // 2
    y = ys.new_empty(B, D, H, W)
    y[:, :, 0::2, 0::2] = ys[:, 0].reshape(B, D, h2, w2)
    y[:, :, 1::2, 0::2] = ys[:, 1].reshape(B, D, w2, h2).transpose(2, 3)
    y[:, :, 0::2, 1::2] = ys[:, 2].reshape(B, D, h2, w2)
    y[:, :, 1::2, 1::2] = ys[:, 3].reshape(B, D, w2, h2).transpose(2, 3)
    return y


def ss2d(sd, p: str, x: Tensor, c: Tensor, rt: Callable = _ID, taps: Optional[dict] = None) -> Tensor:
    """SS2D.forward (src/emamba2.py:713-751) with forward_corev2 / cross_selective_scan (:698-711, 295-367).
    x: (B,H,W,C) channels-last; c: (B,1,256).  Returns (B,H,W,C)."""
    B, H, W, C = x.shape
    local = F.silu(F.linear(c, sd[p + ".attn.0.weight"]))                           # (B,1,2C)  :522-525, 715
    xz = F.linear(x, sd[p + ".in_proj.weight"])                                     # :717
    xx, z = xz.chunk(2, dim=-1)
    z = rt(F.silu(z), "z")                                                               # :720
    xx = rt(xx, "xz_x").permute(0, 3, 1, 2)
    xx = F.silu(F.conv2d(xx, sd[p + ".conv2d.weight"], sd[p + ".conv2d.bias"], padding=1, groups=xx.shape[1]))  # :722
    xx = rt(xx, "xs")
    D = xx.shape[1]
    Wx, Wdt = sd[p + ".x_proj_weight"], sd[p + ".dt_projs_weight"]
    K, _, R = Wdt.shape
    N = sd[p + ".A_logs"].shape[1]
    xs = efficient_scan(xx)                                                         # (B,4,D,L)
    x_dbl = torch.einsum("bkdl,kcd->bkcl", xs, Wx)                                  # :335
    dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)
    dts = rt(torch.einsum("bkrl,kdr->bkdl", dts, Wdt), "dts")                              # :340
    L = xs.shape[-1]
    As = -torch.exp(sd[p + ".A_logs"].float())                                      # :344
    ys = scan_cpu.selective_scan_fwd(xs.reshape(B, K * D, L), dts.reshape(B, K * D, L), As, Bs.contiguous(),
                                     Cs.contiguous(), sd[p + ".Ds"], sd[p + ".dt_projs_bias"].reshape(-1), True)
    if taps is not None:
        taps.update(xs=xs, dts=dts, Bs=Bs, Cs=Cs, ys=ys)
    ys = rt(ys, "ys").reshape(B, K, D, L)
    y = efficient_merge(ys, H, W).permute(0, 2, 3, 1)                               # (B,H,W,D)  :358-365
    y = F.layer_norm(y, (D,), sd[p + ".out_norm.weight"], sd[p + ".out_norm.bias"], eps=1e-5)
    y = rt(y * z + local.unsqueeze(1), "gated")                                              # :747-748
    return F.linear(y, sd[p + ".out_proj.weight"])


def transposed_attention(sd, p: str, x: Tensor, heads: int, rt: Callable = _ID) -> Tensor:
    """TransposedAttention.forward (src/DADiff.py:263-285). x: (B,C,H,W)."""
    B, C, H, W = x.shape
    qkv = rt(F.conv2d(x, sd[p + ".qkv.weight"]), "qkv")
    qkv = F.conv2d(qkv, sd[p + ".qkv_dwconv.weight"], padding=1, groups=3 * C)
    q, k, v = qkv.chunk(3, dim=1)
    v = rt(v, "v")
    q = F.normalize(q.reshape(B, heads, C // heads, H * W), dim=-1)                 # :273
    k = F.normalize(k.reshape(B, heads, C // heads, H * W), dim=-1)
    v = v.reshape(B, heads, C // heads, H * W)
    attn = ((q @ k.transpose(-2, -1)) * sd[p + ".temperature"]).softmax(dim=-1)     # :276-277
    out = (attn @ v).reshape(B, C, H, W)
    return F.conv2d(out, sd[p + ".project_out.weight"])


def mamba_block(sd, p: str, x: Tensor, c: Tensor, t: Tensor, rt: Callable = _ID, taps: Optional[dict] = None) -> Tensor:
    """Mamba_block.forward (src/DADiff.py:477-488). x: (B,C,H,W) -> (B,C,H,W); c: (B,1,256); t: (B,time_dim)."""
    B, C, H, W = x.shape
    x = x.permute(0, 2, 3, 1)
    mod = F.linear(F.silu(t), sd[p + ".adaLN_modulation.1.weight"], sd[p + ".adaLN_modulation.1.bias"])
    sh1, sc1, g1, sh2, sc2, g2 = [m[:, None, None, :] for m in mod.chunk(6, dim=1)]
    a = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], eps=1e-5)
    a = rt(a * (1 + sc1) + sh1, "ln1")
    x = rt(x + g1 * ss2d(sd, p + ".mamba", a, c, rt, taps), "trunk")                         # :486
    a = rt(F.layer_norm(x, (C,), None, None, eps=1e-6) * (1 + sc2) + sh2, "ln2")
    att = transposed_attention(sd, p + ".attn_blk", a.permute(0, 3, 1, 2), C // 32, rt)
    x = rt(x + g2 * att.permute(0, 2, 3, 1), "trunk")                                        # :487
    return x.permute(0, 3, 1, 2)


def time_embedding(sd, time: Tensor, dim: int) -> Tensor:
    """SinusoidalPosEmb(dim) -> Linear -> GELU(erf) -> Linear  (src/DADiff.py:173-185, 580-585)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    e = time[:, None].float() * freq[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = F.gelu(F.linear(e, sd["time_mlp.1.weight"], sd["time_mlp.1.bias"]))
    return F.linear(e, sd["time_mlp.3.weight"], sd["time_mlp.3.bias"])


def prompt_embedding(sd, dose_emb: Tensor) -> Tensor:
    """prompt_mlp(softmax(text_mlp(dose)) * prompt)  (src/DADiff.py:606-611, 706-707)."""
    h = F.linear(F.silu(F.linear(dose_emb, sd["text_mlp.0.weight"], sd["text_mlp.0.bias"])),
                 sd["text_mlp.2.weight"], sd["text_mlp.2.bias"])
    return F.linear(torch.softmax(h, dim=1) * sd["prompt"], sd["prompt_mlp.weight"], sd["prompt_mlp.bias"])


def unet_forward(sd, x: Tensor, time: Tensor, dose_emb: Optional[Tensor] = None, ctx_emb: Optional[Tensor] = None,
                 rt: Callable = _ID, taps: Optional[dict] = None) -> Tensor:
    """Unet.forward (src/DADiff.py:685-740).  x: (B,2,H,W) = cat(x_t, x_input); time: (B,) float
    (= alphas_cumsum[t]*1000, :1162).  dose/ctx embeddings may be passed in (they depend on x[:,1] only)."""
    dim = sd["init_conv.weight"].shape[0]
    if dose_emb is None:
        dose_emb, ctx_emb = daclip_embed(sd, x[:, 1:2])
    c = ctx_emb.unsqueeze(1)
    tap = (lambda k, v: taps.__setitem__(k, v.clone())) if taps is not None else (lambda k, v: None)
    tap("dose_emb", dose_emb), tap("ctx_emb", ctx_emb)
    x = rt(F.conv2d(x, sd["init_conv.weight"], sd["init_conv.bias"], padding=3), "trunk")     # :700
    r = x
    t = time_embedding(sd, time, dim) + prompt_embedding(sd, dose_emb)              # :703-709
    tap("t_emb", t), tap("init_conv", x)
    h: List[Tensor] = []
    for i in range(4):                                                              # :712-719
        x = mamba_block(sd, f"downs.{i}.1", x, c, t, rt)
        tap(f"downs.{i}.mamba", x)
        x = resnet_block(sd, f"downs.{i}.0", x, rt=rt)
        tap(f"downs.{i}.res", x)
        h.append(x)
        w, b = sd[f"downs.{i}.2.weight"], sd[f"downs.{i}.2.bias"]
        x = rt(F.conv2d(x, w, b, stride=2, padding=1) if w.shape[-1] == 4 else F.conv2d(x, w, b, padding=1), "trunk")
        tap(f"downs.{i}.down", x)
    x = resnet_block(sd, "mid_block", x, rt=rt)                                     # :721-722
    x = mamba_block(sd, "mid_attn", x, c, t, rt)
    tap("mid", x)
    for i in range(4):                                                              # :725-731
        x = torch.cat((x, h.pop()), dim=1)
        x = resnet_block(sd, f"ups.{i}.0", x, rt=rt)
        tap(f"ups.{i}.res", x)
        x = mamba_block(sd, f"ups.{i}.1", x, c, t, rt)
        tap(f"ups.{i}.mamba", x)
        if i < 3:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            x = rt(F.conv2d(x, sd[f"ups.{i}.2.1.weight"], sd[f"ups.{i}.2.1.bias"], padding=1), "trunk")
        else:
            x = rt(F.conv2d(x, sd[f"ups.{i}.2.weight"], sd[f"ups.{i}.2.bias"], padding=1), "trunk")
        tap(f"ups.{i}.up", x)
    x = torch.cat((x, r), dim=1)                                                    # :733
    x = resnet_block(sd, "final_res_block", x, rt=rt)
    tap("final_res", x)
    return F.conv2d(x, sd["final_conv.weight"], sd["final_conv.bias"])               # :740


# ----------------------------------------------------------------------------------------------------------
# Sampler  (src/DADiff.py:1153-1380), condition=True, eta forced to 0 (:940-942).  Default objective 'pred_res'
# (train.py:78-82); the other objectives take `objective=` (+ `sd1`, the second Unet's weights, for num_unet = 2).
# ----------------------------------------------------------------------------------------------------------
def eval_plan(objective: str, test_res_or_noise: str, num_unet: int):
    """(branch of model_predictions, [(unet idx, time idx)]) — UnetRes.forward :817-836 + :1168-1207."""
    if num_unet == 1:
        return {"pred_res": ("pred_res", [(0, 0)]), "pred_noise": ("pred_noise", [(0, 1)])}[objective]
    if objective == "pred_res_noise":
        return {"res_noise": ("pred_res_noise", [(0, 0), (1, 1)]), "res": ("pred_res", [(0, 0)]),
                "noise": ("pred_noise", [(1, 1)])}[test_res_or_noise]
    assert objective == "pred_x0_noise" and test_res_or_noise == "res_noise"
    return "pred_x0_noise", [(0, 0), (1, 1)]


def model_predictions(sd, sched, x_input: Tensor, x_t: Tensor, t: int, emb=None, rt: Callable = _ID,
                      num_timesteps: int = 1000, objective: str = "pred_res", test_res_or_noise: str = "res",
                      sd1=None, emb1=None):
    """:1153-1209, all objective branches. Returns (pred_res, pred_noise, x_start).  `emb` / `emb1`: cached DA-CLIP
    embeddings of unet0 / unet1 (each Unet owns a dose encoder)."""
    B = x_t.shape[0]
    acs, bcs, omacs = sched["alphas_cumsum"][t], sched["betas_cumsum"][t], sched["one_minus_alphas_cumsum"][t]
    times = [(acs * num_timesteps).expand(B), (bcs * num_timesteps).expand(B)]      # :1161-1163
    branch, evals = eval_plan(objective, test_res_or_noise, 1 if sd1 is None else 2)
    x_in = torch.cat((x_t, x_input), dim=1)
    outs = {}
    for idx, ti in evals:
        w = sd if idx == 0 else sd1
        e = emb if idx == 0 else emb1
        dose, ctx = e if e is not None else daclip_embed(w, x_input)
        outs[idx] = unet_forward(w, x_in, times[ti], dose, ctx, rt=rt)
    clip = lambda v: v.clamp(-1., 1.)                                               # noqa: E731
    if branch == "pred_res":                                                        # :1176-1182, 1202-1207
        pred_res = clip(outs[0])
        pred_noise = (x_t - x_input - (acs - 1) * pred_res) / bcs                   # :1120-1124
        x_start = clip(x_input - pred_res)
    elif branch == "pred_noise":                                                    # :1183-1189, 1194-1201
        pred_noise = outs[evals[0][0]]
        x_start = clip((x_t - acs * x_input - bcs * pred_noise) / omacs)            # :1126-1130
        pred_res = clip(x_input - x_start)
    elif branch == "pred_res_noise":                                                # :1169-1175
        pred_res, pred_noise = clip(outs[0]), outs[1]
        x_start = clip(x_t - acs * pred_res - bcs * pred_noise)                     # :1132-1136
    else:                                                                           # pred_x0_noise :1188-1192
        pred_res, pred_noise, x_start = clip(x_input - outs[0]), outs[1], clip(outs[0])
    return pred_res, pred_noise, x_start


def ddim_times(sampling_timesteps: int, total: int = 1000):
    times = torch.linspace(-1, total - 1, steps=sampling_timesteps + 1)             # :1287-1291
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


def ddim_sample(sd, sched, x_input: Tensor, init_noise: Tensor, sampling_timesteps: int, sum_scale: float = 0.01,
                last: bool = True, rt: Callable = _ID, trace: Optional[list] = None, **obj):
    """:1275-1365 with condition=True, eta=0, type 'use_pred_noise'.  x_input in [-1,1].  `obj`: objective,
    test_res_or_noise, sd1 (see model_predictions)."""
    img = x_input + math.sqrt(sum_scale) * init_noise                               # :1294-1295
    first = img
    emb = daclip_embed(sd, x_input)
    if obj.get("sd1") is not None:
        obj = dict(obj, emb1=daclip_embed(obj["sd1"], x_input))
    imgs = []
    for t, t_next in ddim_times(sampling_timesteps):
        pred_res, pred_noise, x_start = model_predictions(sd, sched, x_input, img, t, emb, rt, **obj)
        if trace is not None:
            trace.append(dict(t=t, pred_res=pred_res, pred_noise=pred_noise, x_start=x_start))
        if t_next < 0:
            img = x_start                                                           # :1317-1321
        else:
            alpha = sched["alphas_cumsum"][t] - sched["alphas_cumsum"][t_next]     # :1323-1325
            img = img - alpha * pred_res                                            # :1344 (sigma2 = 0)
        imgs.append(img)
    outs = [first] + (imgs if not last else [img])                                  # :1354-1359
    return [(o + 1) * 0.5 for o in outs]


def p_sample_loop(sd, sched, x_input: Tensor, init_noise: Tensor, step_noise: Callable[[int], Tensor],
                  num_timesteps: int = 1000, sum_scale: float = 0.01, last: bool = True, rt: Callable = _ID,
                  trace: Optional[list] = None, **obj):
    """:1232-1273 + p_sample :1221-1230 + q_posterior :1142-1151.  `step_noise(t)` supplies the N(0,1) tensor
    the reference draws with randn_like at step t (t > 0)."""
    img = x_input + math.sqrt(sum_scale) * init_noise
    first = img
    emb = daclip_embed(sd, x_input)
    if obj.get("sd1") is not None:
        obj = dict(obj, emb1=daclip_embed(obj["sd1"], x_input))
    imgs = []
    for t in reversed(range(num_timesteps)):
        # NB the reference scales the Unet time argument by self.num_timesteps (:1162), so a fixture that
        # overrides num_timesteps (tests/golden/ancestral_32.npz) sees alphas_cumsum[t]*num_timesteps.
        pred_res, pred_noise, x_start = model_predictions(sd, sched, x_input, img, t, emb, rt, num_timesteps, **obj)
        mean = (sched["posterior_mean_coef1"][t] * img + sched["posterior_mean_coef2"][t] * pred_res
                + sched["posterior_mean_coef3"][t] * x_start)
        if trace is not None:
            trace.append(dict(t=t, pred_res=pred_res, pred_noise=pred_noise, x_start=x_start))
        if t > 0:
            img = mean + (0.5 * sched["posterior_log_variance_clipped"][t]).exp() * step_noise(t)   # :1228-1229
        else:
            img = mean
        imgs.append(img)
    outs = [first] + (imgs if not last else [img])
    return [(o + 1) * 0.5 for o in outs]


def sample(sd, x_input01: Tensor, init_noise: Tensor, sampling_timesteps: int = 2, step_noise=None,
           schedule_variant: str = "init", last: bool = True, num_timesteps: int = 1000, rt: Callable = _ID,
           trace: Optional[list] = None, **obj):
    """ResidualDiffusion.sample (:1367-1380): x_input01 (B,1,H,W) in [0,1]; returns list in [0,1]."""
    sched = make_schedule(1000, schedule_variant)
    x_input = x_input01 * 2 - 1
    if sampling_timesteps < num_timesteps:
        return ddim_sample(sd, sched, x_input, init_noise, sampling_timesteps, last=last, rt=rt, trace=trace, **obj)
    return p_sample_loop(sd, sched, x_input, init_noise, step_noise, num_timesteps, last=last, rt=rt, trace=trace, **obj)


def p_losses(sd, sched, x_start: Tensor, x_input: Tensor, t: Tensor, noise: Tensor, loss_type: str = "l2",
             num_timesteps: int = 1000, objective: str = "pred_res", sd1=None):
    """:1399-1482 (forward only), condition=True: q_sample (:1382-1388), one model call, per-output loss
    `reduce(loss, 'b ... -> b (...)', 'mean').mean()`.  x_start / x_input in [-1,1]; t (B,) long."""
    ex = lambda n: sched[n][t].view(-1, 1, 1, 1)                                    # noqa: E731
    x_res = x_input - x_start
    x = x_start + ex("alphas_cumsum") * x_res + ex("betas_cumsum") * noise
    times = [sched["alphas_cumsum"][t] * num_timesteps, sched["betas_cumsum"][t] * num_timesteps]
    _, evals = eval_plan(objective, "res_noise", 1 if sd1 is None else 2)
    x_in = torch.cat((x, x_input), dim=1)
    outs = []
    for idx, ti in evals:
        w = sd if idx == 0 else sd1
        dose, ctx = daclip_embed(w, x_input)
        outs.append(unet_forward(w, x_in, times[ti], dose, ctx))
    target = {"pred_res_noise": [x_res, noise], "pred_x0_noise": [x_start, noise], "pred_noise": [noise],
              "pred_res": [x_res]}[objective]
    fn = F.l1_loss if loss_type == "l1" else F.mse_loss
    return [fn(o, tg, reduction="none").flatten(1).mean(1).mean() for o, tg in zip(outs, target)], outs


# ----------------------------------------------------------------------------------------------------------
# Metrics used by the parity gates  (src/util.py:223-232 compute_psnr, max_val = 1)
# ----------------------------------------------------------------------------------------------------------
def psnr(a: Tensor, b: Tensor, max_val: float = 1.0) -> float:
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(max_val ** 2 / mse)


def rel_l2(a: Tensor, ref: Tensor) -> float:
    return (torch.linalg.vector_norm(a.double() - ref.double()) / torch.linalg.vector_norm(ref.double()).clamp_min(1e-30)).item()
