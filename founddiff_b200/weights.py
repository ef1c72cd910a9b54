"""Weight schema of the live FoundDiff denoiser + DA-CLIP visual tower, and a seeded random initialiser.

Key names and shapes are exactly those of the reference `state_dict()` under the `unet0.` prefix
(reference constructors at src/DADiff.py:530-683, src/emamba2.py:404-532,
src/DACLIP.py:168-349, 1135-1188), so a checkpoint saved by the reference (`model-400.pt`, EMA weights, see
src/DADiff.py:1648-1669) can be ingested with `extract_live_weights` and a state dict produced here loads
into the reference modules unchanged (that is how tests/golden fixtures are generated).

Dead parameters of the reference (`clip_model.*` 102 M unused RN50, CLIP text tower, `prompt_learner`,
`perceploss.*`) are not part of the schema: their outputs are discarded by `Unet.forward`
(src/DADiff.py:692) and they never influence `sample()`.

The random initialiser deliberately de-zeroes `adaLN_modulation.1.{weight,bias}` (zeroed at
src/DADiff.py:473-474, which makes every Mamba_block an exact identity) and the CLIP `bn3.weight`
(zeroed at src/DACLIP.py:518-521) so that parity tests exercise the scan, the attention, the time embedding
and the conditioning path (SURVEY.md "Five facts" #1).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Iterable, List, Tuple

import torch

TIME_DIM_MULT = 4
CONTEXT_DIM = 1024          # DA-CLIP dose embedding width (src/DADiff.py:606)
ANATOMY_DIM = 256           # DA-CLIP content embedding width (src/DACLIP.py:1183-1187, src/emamba2.py:523)
BASE_D_STATE = 4            # src/DADiff.py:618
RN50_LAYERS = (3, 4, 6, 3)
RN50_WIDTH = 64
RN50_EMBED = 1024
RN50_HEADS = 32


class UnetConfig:
    """Static geometry of the denoiser (train.py:97-105: dim=64, dim_mults=(1,2,4,8), channels=1)."""

    def __init__(self, dim: int = 64, dim_mults: Tuple[int, ...] = (1, 2, 4, 8), channels: int = 1):
        assert len(dim_mults) == 4, "d_state schedule of the reference is written for 4 resolutions"
        self.dim = dim
        self.dim_mults = tuple(dim_mults)
        self.channels = channels
        self.in_channels = 2 * channels               # cat(x_t, x_input), src/DADiff.py:553-555, 1160
        self.time_dim = dim * TIME_DIM_MULT
        dims = [dim] + [dim * m for m in dim_mults]
        self.in_out = list(zip(dims[:-1], dims[1:]))  # src/DADiff.py:561-562
        self.mid_dim = dims[-1]
        # (hidden, d_state) of the nine Mamba blocks in execution order (src/DADiff.py:632-676)
        self.down_states = [BASE_D_STATE * 2 ** i for i in range(4)]
        self.up_states = [BASE_D_STATE * 2 ** (3 - i) for i in range(4)]

    def mamba_blocks(self) -> List[Tuple[str, int, int]]:
        out = []
        for i, (ci, _co) in enumerate(self.in_out):
            out.append((f"downs.{i}.1", ci, self.down_states[i]))
        out.append(("mid_attn", self.mid_dim, BASE_D_STATE * 8))
        for i, (ci, co) in enumerate(reversed(self.in_out)):
            out.append((f"ups.{i}.1", co, self.up_states[i]))
        return out


def _mamba_schema(prefix: str, C: int, N: int, time_dim: int) -> List[Tuple[str, Tuple[int, ...], str]]:
    R = math.ceil(C / 16)
    D = 2 * C
    p = prefix + "."
    return [
        (p + "norm1.weight", (C,), "norm_w"), (p + "norm1.bias", (C,), "norm_b"),
        (p + "adaLN_modulation.1.weight", (6 * C, time_dim), "adaln_w"),
        (p + "adaLN_modulation.1.bias", (6 * C,), "adaln_b"),
        (p + "mamba.in_proj.weight", (2 * D, C), "linear"),
        (p + "mamba.conv2d.weight", (D, 1, 3, 3), "conv"), (p + "mamba.conv2d.bias", (D,), "bias:9"),
        (p + "mamba.x_proj_weight", (4, R + 2 * N, D), "linear_k"),
        (p + "mamba.dt_projs_weight", (4, D, R), "dt_w"), (p + "mamba.dt_projs_bias", (4, D), "dt_b"),
        (p + "mamba.A_logs", (4 * D, N), "a_log"), (p + "mamba.Ds", (4 * D,), "d_skip"),
        (p + "mamba.out_norm.weight", (D,), "norm_w"), (p + "mamba.out_norm.bias", (D,), "norm_b"),
        (p + "mamba.out_proj.weight", (C, D), "linear"),
        (p + "mamba.attn.0.weight", (D, ANATOMY_DIM), "linear"),
        (p + "attn_blk.temperature", (C // 32, 1, 1), "temperature"),
        (p + "attn_blk.qkv.weight", (3 * C, C, 1, 1), "conv"),
        (p + "attn_blk.qkv_dwconv.weight", (3 * C, 1, 3, 3), "conv"),
        (p + "attn_blk.project_out.weight", (C, C, 1, 1), "conv"),
    ]


def _resblock_schema(prefix: str, ci: int, co: int):
    p = prefix + "."
    s = [
        (p + "block1.proj.weight", (co, ci, 3, 3), "conv"), (p + "block1.proj.bias", (co,), f"bias:{ci * 9}"),
        (p + "block1.norm.weight", (co,), "norm_w"), (p + "block1.norm.bias", (co,), "norm_b"),
    ]
    if ci != co:
        s += [(p + "res_conv.weight", (co, ci, 1, 1), "conv"), (p + "res_conv.bias", (co,), f"bias:{ci}")]
    return s


def unet_schema(cfg: UnetConfig):
    """(key, shape, init-kind) of every live denoiser parameter, reference key order not required."""
    d, td = cfg.dim, cfg.time_dim
    s = [
        ("prompt", (1, td), "uniform01"),
        ("init_conv.weight", (d, cfg.in_channels, 7, 7), "conv"), ("init_conv.bias", (d,), f"bias:{cfg.in_channels * 49}"),
        ("time_mlp.1.weight", (td, d), "linear"), ("time_mlp.1.bias", (td,), f"bias:{d}"),
        ("time_mlp.3.weight", (td, td), "linear"), ("time_mlp.3.bias", (td,), f"bias:{td}"),
        ("text_mlp.0.weight", (td, CONTEXT_DIM), "linear"), ("text_mlp.0.bias", (td,), f"bias:{CONTEXT_DIM}"),
        ("text_mlp.2.weight", (td, td), "linear"), ("text_mlp.2.bias", (td,), f"bias:{td}"),
        ("prompt_mlp.weight", (td, td), "linear"), ("prompt_mlp.bias", (td,), f"bias:{td}"),
    ]
    n = len(cfg.in_out)
    for i, (ci, co) in enumerate(cfg.in_out):
        s += _resblock_schema(f"downs.{i}.0", ci, ci)
        s += _mamba_schema(f"downs.{i}.1", ci, cfg.down_states[i], td)
        k = 4 if i < n - 1 else 3                        # Downsample 4x4 s2 / last: conv3x3 (src/DADiff.py:642-643)
        s += [(f"downs.{i}.2.weight", (co, ci, k, k), "conv"), (f"downs.{i}.2.bias", (co,), f"bias:{ci * k * k}")]
    s += _resblock_schema("mid_block", cfg.mid_dim, cfg.mid_dim)
    s += _mamba_schema("mid_attn", cfg.mid_dim, BASE_D_STATE * 8, td)
    for i, (ci, co) in enumerate(reversed(cfg.in_out)):
        s += _resblock_schema(f"ups.{i}.0", co + ci, co)
        s += _mamba_schema(f"ups.{i}.1", co, cfg.up_states[i], td)
        key = f"ups.{i}.2.1" if i < n - 1 else f"ups.{i}.2"   # nn.Sequential(Upsample, Conv) / plain conv (:674-675)
        s += [(key + ".weight", (ci, co, 3, 3), "conv"), (key + ".bias", (ci,), f"bias:{co * 9}")]
    s += _resblock_schema("final_res_block", 2 * d, d)
    s += [("final_conv.weight", (cfg.channels, d, 1, 1), "conv"), ("final_conv.bias", (cfg.channels,), f"bias:{d}")]
    return s


def _bn_schema(p: str, c: int):
    return [(p + ".weight", (c,), "bn_w"), (p + ".bias", (c,), "bn_b"),
            (p + ".running_mean", (c,), "bn_mean"), (p + ".running_var", (c,), "bn_var"),
            (p + ".num_batches_tracked", (), "zero_long")]


def daclip_schema():
    """Live part of `dose_encoder` (CLIPIQA): RN50 ModifiedResNet visual tower + attnpool + head1/head2."""
    v = "dose_encoder.clip_model.visual."
    w = RN50_WIDTH
    s = [(v + "conv1.weight", (w // 2, 3, 3, 3), "he")] + _bn_schema(v + "bn1", w // 2)
    s += [(v + "conv2.weight", (w // 2, w // 2, 3, 3), "he")] + _bn_schema(v + "bn2", w // 2)
    s += [(v + "conv3.weight", (w, w // 2, 3, 3), "he")] + _bn_schema(v + "bn3", w)
    inplanes = w
    for li, blocks in enumerate(RN50_LAYERS):
        planes = w * 2 ** li
        for bi in range(blocks):
            p = f"{v}layer{li + 1}.{bi}."
            stride = 2 if (li > 0 and bi == 0) else 1
            s += [(p + "conv1.weight", (planes, inplanes, 1, 1), "he")] + _bn_schema(p + "bn1", planes)
            s += [(p + "conv2.weight", (planes, planes, 3, 3), "he")] + _bn_schema(p + "bn2", planes)
            s += [(p + "conv3.weight", (planes * 4, planes, 1, 1), "he")] + _bn_schema(p + "bn3", planes * 4)
            if stride > 1 or inplanes != planes * 4:
                s += [(p + "downsample.0.weight", (planes * 4, inplanes, 1, 1), "he")] + _bn_schema(p + "downsample.1", planes * 4)
            inplanes = planes * 4
    e = w * 32
    a = v + "attnpool."
    s += [(a + "positional_embedding", ((224 // 32) ** 2 + 1, e), "small")]
    for nm, o in (("k_proj", e), ("q_proj", e), ("v_proj", e), ("c_proj", RN50_EMBED)):
        s += [(a + nm + ".weight", (o, e), "attnpool"), (a + nm + ".bias", (o,), f"bias:{e}")]
    for h, o in (("head1", CONTEXT_DIM), ("head2", ANATOMY_DIM)):
        s += [(f"dose_encoder.{h}.0.weight", (1024, 1024), "linear"), (f"dose_encoder.{h}.0.bias", (1024,), "bias:1024"),
              (f"dose_encoder.{h}.2.weight", (o, 1024), "linear"), (f"dose_encoder.{h}.2.bias", (o,), "bias:1024")]
    return s


def full_schema(cfg: UnetConfig):
    return unet_schema(cfg) + daclip_schema()


def _fan_in(shape):
    f = 1
    for d in shape[1:]:
        f *= d
    return max(f, 1)


def _init(kind: str, shape, g: torch.Generator) -> torch.Tensor:
    def U(lo, hi):
        return torch.rand(shape, generator=g) * (hi - lo) + lo

    def Nrm(std, mean=0.0):
        return torch.randn(shape, generator=g) * std + mean

    if kind in ("linear", "conv"):                    # nn.Linear / nn.Conv2d default: U(+-1/sqrt(fan_in))
        b = 1.0 / math.sqrt(_fan_in(shape))
        return U(-b, b)
    if kind == "linear_k":                            # (K, out, in) stack of nn.Linear weights
        b = 1.0 / math.sqrt(shape[-1])
        return U(-b, b)
    if kind.startswith("bias:"):
        b = 1.0 / math.sqrt(int(kind.split(":")[1]))
        return U(-b, b)
    if kind == "he":
        return Nrm(math.sqrt(2.0 / _fan_in(shape)))
    if kind == "norm_w":
        return Nrm(0.1, 1.0)
    if kind == "norm_b":
        return Nrm(0.1)
    if kind == "adaln_w":
        return Nrm(0.02)
    if kind == "adaln_b":
        return Nrm(0.5)
    if kind == "uniform01":
        return U(0.0, 1.0)
    if kind == "dt_w":                                # src/emamba2.py:539-543
        b = shape[-1] ** -0.5
        return U(-b, b)
    if kind == "dt_b":                                # inverse softplus of dt in [1e-3, 1e-1] (src/emamba2.py:548-555)
        dt = torch.exp(torch.rand(shape, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001)).clamp(min=1e-4)
        return dt + torch.log(-torch.expm1(-dt))
    if kind == "a_log":                               # S4D-real: log(1..N) (src/emamba2.py:560-574), jittered
        n = shape[1]
        base = torch.log(torch.arange(1, n + 1, dtype=torch.float32)).expand(shape)
        return base + Nrm(0.05)
    if kind == "d_skip":
        return Nrm(0.1, 1.0)
    if kind == "temperature":
        return U(0.5, 2.0)
    if kind == "bn_w":
        return U(0.5, 1.0)
    if kind == "bn_b":
        return Nrm(0.05)
    if kind == "bn_mean":
        return Nrm(0.05)
    if kind == "bn_var":
        return U(0.8, 1.2)
    if kind == "zero_long":
        return torch.zeros(shape, dtype=torch.long)
    if kind == "small":
        return Nrm(shape[-1] ** -0.5)
    if kind == "attnpool":
        return Nrm(shape[-1] ** -0.5)
    raise KeyError(kind)


def random_state_dict(seed: int = 10, cfg: UnetConfig | None = None, with_daclip: bool = True) -> "OrderedDict[str, torch.Tensor]":
    """Deterministic (torch CPU generator) weights for every live parameter, reference key names."""
    cfg = cfg or UnetConfig()
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    sd = OrderedDict()
    schema = unet_schema(cfg) + (daclip_schema() if with_daclip else [])
    for key, shape, kind in schema:
        sd[key] = _init(kind, tuple(shape), g).contiguous()
    return sd


_PREFIXES = ("ema_model.model.unet0.", "model.unet0.", "unet0.", "")


def extract_live_weights(state_dict: Dict[str, torch.Tensor], cfg: UnetConfig | None = None, unet: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Pick the live keys of `unet{unet}` out of a reference checkpoint state dict (any of the EMA / diffusion / UnetRes / Unet
    prefixes, src/DADiff.py:1630-1636 + SURVEY Appendix B) and validate shapes.  Dead keys are dropped."""
    cfg = cfg or UnetConfig()
    out = OrderedDict()
    schema = full_schema(cfg)
    prefixes = tuple(p.replace("unet0.", f"unet{unet}.") for p in _PREFIXES if unet == 0 or "unet0." in p)
    for prefix in prefixes:
        if prefix + schema[0][0] in state_dict:
            break
    else:
        raise KeyError(f"no FoundDiff unet{unet} weights found (looked for 'prompt' under " + ", ".join(map(repr, prefixes)) + ")")
    for key, shape, _ in schema:
        t = state_dict[prefix + key]
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{key}: expected {tuple(shape)}, got {tuple(t.shape)}")
        out[key] = t.detach()
    return out
