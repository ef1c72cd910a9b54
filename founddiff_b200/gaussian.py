"""Secondary sampling path of the reference: the lucidrains `Unet` and the
epsilon-prediction `GaussianDiffusion` of src/denoising_diffusion_pytorch.py (`train.py:84-95`,
`original_ddim_ddpm=True`), behind the same class names / constructor arguments / `sample(batch_size=...)` call.

Same design as the primary path (engine.py / diffusion.py): channels-last activations in a 16-bit or fp32 storage
type, every layer a hand-written kernel of libfounddiff_b200 —
  ResnetBlock  (:201-225)  WS-conv3x3 on fd_conv2d_tc (GroupNorm partial sums in the epilogue) -> fd_gn_scale_shift_silu
                           (time scale/shift) -> conv3x3 -> fd_gn_silu_add (+ identity / 1x1 res_conv)
  LinearAttention (:227-255)  channel LayerNorm (fd_ln_modulate) -> 1x1 qkv -> fd_linattn_context / fd_linattn_weff /
                           fd_softmax_d32 -> per-sample 1x1 GEMM -> channel LayerNorm, residual
  Attention    (:257-279)  channel LayerNorm -> 1x1 qkv -> fd_flash_attn_d32 -> 1x1 to_out with the residual as addend
  Down/Upsample (:103-110) 4x4 stride-2 conv / nearest x2 folded into four phase 2x2 convolutions
  sampler      (:547-646)  fd_ddpm_update (x0 from eps, clip, posterior mean or DDIM step, noise injection) per step
There is no CPU path.  Parity: tests/golden/gaussian_*.npz, generated from the unmodified reference by
oracle/gen_golden_gaussian.py.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import ops
from .engine import GN_GROUPS, pack_conv, ws_fold
from .weights import _init

HEADS, DIM_HEAD = 4, 32


# --------------------------------------------------------------------------------------------------------------
def gaussian_unet_schema(dim: int = 64, dim_mults=(1, 2, 4, 8), channels: int = 3):
    """(key, shape, init-kind) of every parameter of the reference `Unet` (src/denoising_diffusion_pytorch.py:283-369)."""
    td = dim * 4
    dims = [dim] + [dim * m for m in dim_mults]
    in_out = list(zip(dims[:-1], dims[1:]))
    hid = HEADS * DIM_HEAD

    def res(p, ci, co):
        s = [(f"{p}.mlp.1.weight", (2 * co, td), "linear"), (f"{p}.mlp.1.bias", (2 * co,), f"bias:{td}")]
        for b, c_in in (("block1", ci), ("block2", co)):
            s += [(f"{p}.{b}.proj.weight", (co, c_in, 3, 3), "conv"), (f"{p}.{b}.proj.bias", (co,), f"bias:{c_in * 9}"),
                  (f"{p}.{b}.norm.weight", (co,), "norm_w"), (f"{p}.{b}.norm.bias", (co,), "norm_b")]
        if ci != co:
            s += [(f"{p}.res_conv.weight", (co, ci, 1, 1), "conv"), (f"{p}.res_conv.bias", (co,), f"bias:{ci}")]
        return s

    def linattn(p, c):
        return [(f"{p}.fn.norm.g", (1, c, 1, 1), "norm_w"), (f"{p}.fn.fn.to_qkv.weight", (3 * hid, c, 1, 1), "conv"),
                (f"{p}.fn.fn.to_out.0.weight", (c, hid, 1, 1), "conv"), (f"{p}.fn.fn.to_out.0.bias", (c,), f"bias:{hid}"),
                (f"{p}.fn.fn.to_out.1.g", (1, c, 1, 1), "norm_w")]

    s = [("init_conv.weight", (dim, channels, 7, 7), "conv"), ("init_conv.bias", (dim,), f"bias:{channels * 49}"),
         ("time_mlp.1.weight", (td, dim), "linear"), ("time_mlp.1.bias", (td,), f"bias:{dim}"),
         ("time_mlp.3.weight", (td, td), "linear"), ("time_mlp.3.bias", (td,), f"bias:{td}")]
    n = len(in_out)
    for i, (ci, co) in enumerate(in_out):
        s += res(f"downs.{i}.0", ci, ci) + res(f"downs.{i}.1", ci, ci) + linattn(f"downs.{i}.2", ci)
        k = 4 if i < n - 1 else 3
        s += [(f"downs.{i}.3.weight", (co, ci, k, k), "conv"), (f"downs.{i}.3.bias", (co,), f"bias:{ci * k * k}")]
    mid = dims[-1]
    s += res("mid_block1", mid, mid)
    s += [("mid_attn.fn.norm.g", (1, mid, 1, 1), "norm_w"), ("mid_attn.fn.fn.to_qkv.weight", (3 * hid, mid, 1, 1), "conv"),
          ("mid_attn.fn.fn.to_out.weight", (mid, hid, 1, 1), "conv"), ("mid_attn.fn.fn.to_out.bias", (mid,), f"bias:{hid}")]
    s += res("mid_block2", mid, mid)
    for i, (ci, co) in enumerate(reversed(in_out)):
        s += res(f"ups.{i}.0", co + ci, co) + res(f"ups.{i}.1", co + ci, co) + linattn(f"ups.{i}.2", co)
        key = f"ups.{i}.3.1" if i < n - 1 else f"ups.{i}.3"
        s += [(key + ".weight", (ci, co, 3, 3), "conv"), (key + ".bias", (ci,), f"bias:{co * 9}")]
    s += res("final_res_block", 2 * dim, dim)
    s += [("final_conv.weight", (channels, dim, 1, 1), "conv"), ("final_conv.bias", (channels,), f"bias:{dim}")]
    return s


def random_gaussian_state_dict(seed: int = 11, dim: int = 64, dim_mults=(1, 2, 4, 8), channels: int = 3):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return OrderedDict((k, _init(kind, tuple(shape), g).contiguous()) for k, shape, kind in gaussian_unet_schema(dim, dim_mults, channels))


def _register(root: nn.Module, key: str, tensor: torch.Tensor):
    parts = key.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, nn.Module())
        m = m._modules[p]
    m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class Unet(nn.Module):
    """Parameter container with the reference `Unet` state_dict layout; `forward` runs the CUDA engine."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=3, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False, random_fourier_features=False,
                 learned_sinusoidal_dim=16, seed: int = 11):
        super().__init__()
        if self_condition or learned_variance or learned_sinusoidal_cond or random_fourier_features:
            raise NotImplementedError("only the configuration train.py builds (Unet(dim=64, dim_mults=(1,2,4,8))) is implemented")
        if init_dim not in (None, dim) or out_dim not in (None, channels) or resnet_block_groups != 8 or len(dim_mults) != 4:
            raise NotImplementedError("init_dim / out_dim / resnet_block_groups / depth overrides are not implemented")
        self.dim, self.dim_mults, self.channels, self.out_dim = dim, tuple(dim_mults), channels, channels
        self.self_condition = False
        self.random_or_learned_sinusoidal_cond = False
        self.compute_dtype = torch.float16
        for k, v in random_gaussian_state_dict(seed, dim, dim_mults, channels).items():
            _register(self, k, v)
        self._engines: Dict = {}
        # a post-hook (not a load_state_dict override) also runs when a parent module (GaussianDiffusion, an EMA wrapper) loads
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._engines.clear())

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)
        self._engines.clear()
        return out

    def engine(self, B, H, W, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())       # 'cuda' and 'cuda:0' are one engine
        key = (B, H, W, self.compute_dtype, str(device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines.clear()
            sd = OrderedDict((k, v.detach()) for k, v in self.state_dict().items())
            eng = GaussianUnetEngine(sd, self.dim, self.dim_mults, self.channels, B, H, W, dtype=self.compute_dtype, device=device)
            self._engines[key] = eng
        return eng

    @torch.no_grad()
    def forward(self, x, time, x_self_cond=None):
        """x (B, channels, H, W) fp32 on a CUDA device, time (B,) -> model output (B, channels, H, W) fp32."""
        if not x.is_cuda:
            raise RuntimeError("founddiff_b200 has no CPU path")
        B, C, H, W = x.shape
        eng = self.engine(B, H, W, x.device)
        eng.x_t.copy_(x.permute(0, 2, 3, 1).reshape(B, H * W, C))
        eng.time.copy_(time.to(torch.float32).reshape(-1).expand(B))
        eng.forward()
        return eng.eps.view(B, H, W, C).permute(0, 3, 1, 2).contiguous()


# --------------------------------------------------------------------------------------------------------------
class GaussianUnetEngine:
    """Fixed kernel sequence of one `Unet.forward` (:371-410) over pre-allocated channels-last buffers."""

    def __init__(self, sd, dim, dim_mults, channels, B, H, W, dtype=torch.float16, device="cuda", prefer_tc=True):
        if H % 8 or W % 8:
            raise ValueError("H and W must be multiples of 8 (three 2x downsamplings)")
        self.sd = {k: v.detach().to(device) for k, v in sd.items()}
        self.B, self.H, self.W, self.dtype, self.device = B, H, W, dtype, torch.device(device)
        self.dim, self.channels, self.prefer_tc = dim, channels, prefer_tc and dtype != torch.float32
        dev, dt = self.device, dtype
        f32 = lambda k: self.sd[k].to(torch.float32).contiguous()  # noqa: E731
        self.f32 = f32
        td = dim * 4
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)  # noqa: E731
        self.time, self.t_sin, self.t_hid, self.t_emb = z(B), z(B, dim), z(B, td), z(B, td)
        self.x_t = z(B, H * W, channels)                 # sampler state, fp32 channels-last
        self.x_in = torch.zeros(B, H * W, channels, device=dev, dtype=dt)
        self.eps = z(B, H * W, channels)
        self.zero_mod = z(B, 8 * dim)                    # shift = scale = 0 rows for the plain channel LayerNorm
        self.steps: List = []
        self._acc_n = 0
        self._acc_bind: List = []
        # every ResnetBlock's time projection in one GEMV: mlp weights concatenated (:203-206, 215-218)
        self._mlp_w, self._mlp_b, self._mlp_off = [], [], {}
        self._build()
        self.acc_buf = z(max(self._acc_n, 1))
        for fn in self._acc_bind:
            fn()
        self.mlp_w = torch.cat(self._mlp_w, dim=0).contiguous()
        self.mlp_b = torch.cat(self._mlp_b, dim=0).contiguous()
        self.ss = z(B, self.mlp_w.shape[0])

    # ------------------------------------------------------------------------------------------------------
    def _new(self, P, C):
        return torch.empty(self.B, P, C, device=self.device, dtype=self.dtype)

    def _acc(self, n):
        off = self._acc_n
        self._acc_n += n
        return off

    def _conv(self, src, w_key, b_key, out, h, w, k, stride=1, upsample=False, src1=None, **kw):
        wt = pack_conv(self.sd[w_key], self.dtype, self.device)
        c = ops.Conv(src, wt, out, B=self.B, Hin=h, Win=w, KH=k, KW=k, stride=stride, pad=(k - 1) // 2 if k != 4 else 1,
                     upsample=upsample, src1=src1, bias=self.f32(b_key) if b_key else None, prefer_tc=self.prefer_tc, **kw)
        return c

    def _resblock(self, p, srcs, cout, h, w):
        """ResnetBlock :201-225; returns the output buffer."""
        B, dev, dt = self.B, self.device, self.dtype
        P = h * w
        cin = sum(s.shape[-1] for s in srcs)
        src1 = srcs[1] if len(srcs) > 1 else None
        off = sum(t.shape[0] for t in self._mlp_w)
        self._mlp_w.append(self.f32(p + ".mlp.1.weight"))
        self._mlp_b.append(self.f32(p + ".mlp.1.bias"))
        y1, h1, y2, out = self._new(P, cout), self._new(P, cout), self._new(P, cout), self._new(P, cout)
        w1 = pack_conv(ws_fold(self.sd[p + ".block1.proj.weight"]), dt, dev)
        w2 = pack_conv(ws_fold(self.sd[p + ".block2.proj.weight"]), dt, dev)
        b1, b2 = self.f32(p + ".block1.proj.bias"), self.f32(p + ".block2.proj.bias")
        g1, be1 = self.f32(p + ".block1.norm.weight"), self.f32(p + ".block1.norm.bias")
        g2, be2 = self.f32(p + ".block2.norm.weight"), self.f32(p + ".block2.norm.bias")
        a1, a2 = self._acc(B * GN_GROUPS * 2), self._acc(B * GN_GROUPS * 2)
        has_res = (p + ".res_conv.weight") in self.sd
        if has_res:
            sk = self._new(P, cout)
            rc = self._conv(srcs[0], p + ".res_conv.weight", p + ".res_conv.bias", sk, h, w, 1, src1=src1)
        else:
            assert len(srcs) == 1 and cin == cout
            sk = srcs[0]
        holder = {}

        def bind():
            s1 = self.acc_buf[a1:a1 + B * GN_GROUPS * 2].view(B, GN_GROUPS, 2)
            s2 = self.acc_buf[a2:a2 + B * GN_GROUPS * 2].view(B, GN_GROUPS, 2)
            holder["s1"], holder["s2"] = s1, s2
            holder["c1"] = ops.Conv(srcs[0], w1, y1, B=B, Hin=h, Win=w, KH=3, KW=3, pad=1, src1=src1, bias=b1, gn_sums=s1,
                                    gn_groups=GN_GROUPS, prefer_tc=self.prefer_tc)
            holder["c2"] = ops.Conv(h1, w2, y2, B=B, Hin=h, Win=w, KH=3, KW=3, pad=1, bias=b2, gn_sums=s2, gn_groups=GN_GROUPS,
                                    prefer_tc=self.prefer_tc)
        self._acc_bind.append(bind)

        def run():
            total = self.ss.shape[1]
            scale = ops_view(self.ss, off, cout)
            shift = ops_view(self.ss, off + cout, cout)
            holder["c1"].run()
            ops.gn_scale_shift_silu(y1, holder["s1"], g1, be1, scale, shift, total, None, h1, B, P, cout, GN_GROUPS)
            holder["c2"].run()
            if has_res:
                rc.run()
            ops.gn_silu_add(y2, holder["s2"], g2, be2, sk, out, B, P, cout, GN_GROUPS)
        self.steps.append(run)
        return out

    def _channel_ln(self, x, g_key, out, P, C):
        g = self.f32(g_key).reshape(-1).contiguous()
        zb = torch.zeros_like(g)
        zm = self.zero_mod
        self.steps.append(lambda: ops.ln_modulate(x, out, g, zb, zm, zm, zm.shape[1], self.B, P, C, 1e-5))

    def _linattn(self, p, x, C, h, w):
        """Residual(PreNorm(LinearAttention)) :227-255, in place on x."""
        B, P, hid = self.B, h * w, HEADS * DIM_HEAD
        xn, qkv, tmp = self._new(P, C), self._new(P, 3 * hid), self._new(P, C)
        self._channel_ln(x, p + ".fn.norm.g", xn, P, C)
        cq = self._conv(xn, p + ".fn.fn.to_qkv.weight", None, qkv, h, w, 1)
        wout = self.f32(p + ".fn.fn.to_out.0.weight").reshape(C, hid).contiguous()
        bout = self.f32(p + ".fn.fn.to_out.0.bias")
        g = self.f32(p + ".fn.fn.to_out.1.g").reshape(-1).contiguous()

        la = ops.LinearAttention(qkv, wout, bout, g, tmp, B, h, w, HEADS, C, prefer_tc=self.prefer_tc)

        def run():
            cq.run()
            la.run()
            x.add_(tmp)                                   # the Residual wrapper (:95-101)
        self.steps.append(run)

    def _attention(self, p, x, C, h, w):
        """Residual(PreNorm(Attention)) :257-279, in place on x."""
        B, P, hid = self.B, h * w, HEADS * DIM_HEAD
        xn, qkv, o = self._new(P, C), self._new(P, 3 * hid), self._new(P, hid)
        self._channel_ln(x, p + ".fn.norm.g", xn, P, C)
        cq = self._conv(xn, p + ".fn.fn.to_qkv.weight", None, qkv, h, w, 1)
        co = self._conv(o, p + ".fn.fn.to_out.weight", p + ".fn.fn.to_out.bias", x, h, w, 1, addend=x)

        def run():
            cq.run()
            ops.flash_attn_d32(qkv, o, B, P, HEADS, DIM_HEAD ** -0.5)
            co.run()
        self.steps.append(run)

    def _build(self):
        B, H, W, dim = self.B, self.H, self.W, self.dim
        dims = [dim] + [dim * m for m in (1, 2, 4, 8)]
        in_out = list(zip(dims[:-1], dims[1:]))
        sizes = [(H >> i, W >> i) for i in range(4)]
        # init_conv on the 16-bit copy of the fp32 sampler state
        r = self._new(H * W, dim)
        ci = self._conv(self.x_in, "init_conv.weight", "init_conv.bias", r, H, W, 7)
        self.steps.append(lambda: (self.x_in.copy_(self.x_t), ci.run()))
        x = r
        hs = []
        for i, (cin, cout) in enumerate(in_out):
            h, w = sizes[i]
            x = self._resblock(f"downs.{i}.0", [x], cin, h, w)
            hs.append(x)
            x2 = self._resblock(f"downs.{i}.1", [x], cin, h, w)
            self._linattn(f"downs.{i}.2", x2, cin, h, w)
            hs.append(x2)
            if i < 3:
                nh, nw = sizes[i + 1]
                nx = self._new(nh * nw, cout)
                c = self._conv(x2, f"downs.{i}.3.weight", f"downs.{i}.3.bias", nx, h, w, 4, stride=2)
            else:
                nx = self._new(h * w, cout)
                c = self._conv(x2, f"downs.{i}.3.weight", f"downs.{i}.3.bias", nx, h, w, 3)
            self.steps.append(c.run)
            x = nx
        h, w = sizes[3]
        mid = dims[-1]
        x = self._resblock("mid_block1", [x], mid, h, w)
        self._attention("mid_attn", x, mid, h, w)
        x = self._resblock("mid_block2", [x], mid, h, w)
        for i, (cin, cout) in enumerate(reversed(in_out)):
            l = 3 - i
            h, w = sizes[l]
            x = self._resblock(f"ups.{i}.0", [x, hs.pop()], cout, h, w)
            x = self._resblock(f"ups.{i}.1", [x, hs.pop()], cout, h, w)
            self._linattn(f"ups.{i}.2", x, cout, h, w)
            if i < 3:
                nh, nw = sizes[l - 1]
                nx = self._new(nh * nw, cin)
                c = self._conv(x, f"ups.{i}.3.1.weight", f"ups.{i}.3.1.bias", nx, h, w, 3, upsample=True)
            else:
                nx = self._new(h * w, cin)
                c = self._conv(x, f"ups.{i}.3.weight", f"ups.{i}.3.bias", nx, h, w, 3)
            self.steps.append(c.run)
            x = nx
        x = self._resblock("final_res_block", [x, r], dim, H, W)
        out16 = self._new(H * W, self.channels)
        cf = self._conv(x, "final_conv.weight", "final_conv.bias", out16, H, W, 1)
        self.steps.append(lambda: (cf.run(), self.eps.copy_(out16)))
        self.time_w1, self.time_b1 = self.f32("time_mlp.1.weight"), self.f32("time_mlp.1.bias")
        self.time_w2, self.time_b2 = self.f32("time_mlp.3.weight"), self.f32("time_mlp.3.bias")

    def forward(self):
        """One Unet evaluation on (self.x_t, self.time) -> self.eps (B, H*W, channels) fp32."""
        self.acc_buf.zero_()
        ops.time_sinusoid(self.time, self.t_sin)
        ops.linear_small(self.t_sin, self.time_w1, self.time_b1, self.t_hid, act_out=2)        # GELU
        ops.linear_small(self.t_hid, self.time_w2, self.time_b2, self.t_emb)
        ops.linear_small(self.t_emb, self.mlp_w, self.mlp_b, self.ss, act_in=1)                # SiLU -> Linear, all blocks
        for fn in self.steps:
            fn()
        return self.eps


def ops_view(t: torch.Tensor, off: int, n: int):
    """Column slice [off, off+n) of a (B, total) fp32 buffer as a raw-pointer view (row pitch = total)."""
    from .engine import _view_ptr
    return _view_ptr(t[:, off:off + n])


# --------------------------------------------------------------------------------------------------------------
def make_schedule(timesteps: int = 1000, beta_schedule: str = "cosine"):
    """float64 schedule tables of GaussianDiffusion.__init__ (:464-521), cast to fp32 like `register_buffer` does."""
    if beta_schedule == "cosine":
        x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
        ac = torch.cos(((x / timesteps) + 0.008) / 1.008 * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    elif beta_schedule == "linear":
        scale = 1000 / timesteps
        betas = torch.linspace(scale * 0.0001, scale * 0.02, timesteps, dtype=torch.float64)
    else:
        raise ValueError(f"unknown beta schedule {beta_schedule}")
    alphas = 1. - betas
    abar = torch.cumprod(alphas, dim=0)
    abar_prev = torch.nn.functional.pad(abar[:-1], (1, 0), value=1.)
    pv = betas * (1. - abar_prev) / (1. - abar)
    f = lambda t: t.to(torch.float32)  # noqa: E731
    return OrderedDict(betas=f(betas), alphas_cumprod=f(abar), alphas_cumprod_prev=f(abar_prev),
                       sqrt_alphas_cumprod=f(torch.sqrt(abar)), sqrt_one_minus_alphas_cumprod=f(torch.sqrt(1. - abar)),
                       log_one_minus_alphas_cumprod=f(torch.log(1. - abar)), sqrt_recip_alphas_cumprod=f(torch.sqrt(1. / abar)),
                       sqrt_recipm1_alphas_cumprod=f(torch.sqrt(1. / abar - 1)), posterior_variance=f(pv),
                       posterior_log_variance_clipped=f(torch.log(pv.clamp(min=1e-20))),
                       posterior_mean_coef1=f(betas * torch.sqrt(abar_prev) / (1. - abar)),
                       posterior_mean_coef2=f((1. - abar_prev) * torch.sqrt(alphas) / (1. - abar)),
                       p2_loss_weight=f((1 + abar / (1 - abar)) ** -0.))


class GaussianDiffusion(nn.Module):
    """src/denoising_diffusion_pytorch.py:437-652 (sampling side; objective 'pred_noise')."""

    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, loss_type='l1', objective='pred_noise',
                 beta_schedule='cosine', p2_loss_weight_gamma=0., p2_loss_weight_k=1, ddim_sampling_eta=0.):
        super().__init__()
        if objective != 'pred_noise':
            raise NotImplementedError("only objective='pred_noise' (the constructor default train.py uses) is implemented")
        self.model = model
        self.channels = model.channels
        self.self_condition = False
        self.image_size = image_size
        self.objective = objective
        self.num_timesteps = int(timesteps)
        self.loss_type = loss_type
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        self.ddim_sampling_eta = ddim_sampling_eta
        for k, v in make_schedule(timesteps, beta_schedule).items():
            self.register_buffer(k, v)
        self._host = {k: v.double().tolist() for k, v in make_schedule(timesteps, beta_schedule).items()}
        self.use_cuda_graph = True                        # one CUDA graph per timestep, as the primary path
        self._graphs: Dict = {}

    def _plan(self):
        """[(t, coef[8])] for fd_ddpm_update: {sr, srm1, a0 (x0), a1 (x_t), a2 (eps), a3 (noise), clip, 0}."""
        hs = self._host
        plan = []
        if not self.is_ddim_sampling:
            for t in reversed(range(self.num_timesteps)):
                sig = math.exp(0.5 * hs["posterior_log_variance_clipped"][t]) if t > 0 else 0.
                plan.append((t, [hs["sqrt_recip_alphas_cumprod"][t], hs["sqrt_recipm1_alphas_cumprod"][t],
                                 hs["posterior_mean_coef1"][t], hs["posterior_mean_coef2"][t], 0., sig, 1., 0.]))
            return plan
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
        times = list(reversed(times.int().tolist()))
        for t, tn in zip(times[:-1], times[1:]):
            sr, srm1 = hs["sqrt_recip_alphas_cumprod"][t], hs["sqrt_recipm1_alphas_cumprod"][t]
            if tn < 0:
                plan.append((t, [sr, srm1, 1., 0., 0., 0., 1., 0.]))
                continue
            a, an = float(self.alphas_cumprod[t]), float(self.alphas_cumprod[tn])        # fp32 buffers, as the reference reads them
            sigma = self.ddim_sampling_eta * math.sqrt((1 - a / an) * (1 - an) / (1 - a))
            c = math.sqrt(1 - an - sigma ** 2)
            plan.append((t, [sr, srm1, math.sqrt(an), 0., c, sigma, 1., 0.]))
        return plan

    @torch.no_grad()
    def sample(self, x_input=0, batch_size=16, *, noise=None, trace: Optional[list] = None):
        """:648-652.  Returns [img] with img (B, channels, S, S) in [0, 1].  `noise` (additive keyword): dict with 'init'
        (B, C, S, S) and 'steps': callable t -> (B, C, S, S); default: torch.randn on the device in the reference's order."""
        dev = self.betas.device
        if dev.type != "cuda":
            raise RuntimeError("founddiff_b200 has no CPU path: move the module to a CUDA device")
        B, C, S = batch_size, self.channels, self.image_size
        eng = self.model.engine(B, S, S, dev)
        nhwc = lambda t: t.to(dev, torch.float32).permute(0, 2, 3, 1).reshape(B, S * S, C)  # noqa: E731
        init = noise["init"] if noise is not None else torch.randn(B, C, S, S, device=dev)
        eng.x_t.copy_(nhwc(init))
        bufs = eng.__dict__.setdefault("_sampler_bufs", {})
        if "coef" not in bufs:
            bufs.update(coef=torch.zeros(8, device=dev), nz=torch.zeros(B, S * S, C, device=dev), x0=torch.zeros(B, S * S, C, device=dev))
        coef, nz, x0 = bufs["coef"], bufs["nz"], bufs["x0"]
        plan = self._plan()
        coef_host = torch.tensor([c for _, c in plan], dtype=torch.float32).pin_memory()
        time_host = torch.tensor([[float(t)] * B for t, _ in plan], dtype=torch.float32).pin_memory()

        def one_step():                                   # Unet + fused x0 / clip / posterior-or-DDIM update (+ noise)
            eng.forward()
            ops.ddpm_update(eng.x_t, eng.eps, nz, coef, eng.x_t, x0)
        step_fn = self._graphed(eng, one_step) if self.use_cuda_graph else one_step
        for i, (t, c) in enumerate(plan):
            coef.copy_(coef_host[i], non_blocking=True)
            eng.time.copy_(time_host[i], non_blocking=True)
            if c[5] != 0.:
                nz.copy_(nhwc(noise["steps"](t)) if noise is not None else nhwc(torch.randn(B, C, S, S, device=dev)))
            step_fn()
            if trace is not None:
                to_nchw = lambda v: v.view(B, S, S, C).permute(0, 3, 1, 2).clone()  # noqa: E731
                trace.append(dict(t=t, pred_noise=to_nchw(eng.eps), x_start=to_nchw(x0)))
        img = (eng.x_t.view(B, S, S, C).permute(0, 3, 1, 2) + 1) * 0.5
        return [img.contiguous()]

    def _graphed(self, eng, fn):
        """One timestep as a CUDA graph (captured once per engine; time / coefficients / noise live in device buffers)."""
        # the cache entry holds the engine itself: a dropped engine's id() can be handed to its successor by CPython, and a
        # graph replayed against freed buffers / old weights fails silently
        ent = self._graphs.get("step")
        g = ent[0] if ent is not None and ent[1] is eng else None
        if g is None:
            self._graphs.clear()
            saved = eng.x_t.clone()
            side = torch.cuda.Stream(device=eng.device)
            side.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(side):
                fn()                                      # warm-up outside capture
            torch.cuda.current_stream(eng.device).wait_stream(side)
            torch.cuda.synchronize(eng.device)
            eng.x_t.copy_(saved)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            eng.x_t.copy_(saved)
            self._graphs["step"] = (g, eng)
        return g.replay

    def p_sample_loop(self, shape, **kw):
        assert not self.is_ddim_sampling
        return self.sample(batch_size=shape[0], **kw)

    def ddim_sample(self, shape, clip_denoised=True, **kw):
        assert self.is_ddim_sampling and clip_denoised
        return self.sample(batch_size=shape[0], **kw)
