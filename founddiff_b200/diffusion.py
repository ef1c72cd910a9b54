"""Host-side mirror of the reference's model-construction and sampling API (train.py:97-119):

    model = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, input_condition=False,
                    objective='pred_res', test_res_or_noise='res')
    diffusion = ResidualDiffusion(model, image_size=512, timesteps=1000, sampling_timesteps=2, objective='pred_res',
                                  loss_type='l2', condition=True, sum_scale=0.01, ...)
    diffusion.init(); out = diffusion.sample([ldct], batch_size=B, last=True)[-1]        # src/DADiff.py:1818, 1868-1870

Same class names, constructor arguments, attribute names, `state_dict()` keys (for the live parameters) and return
structure as src/DADiff.py:743-836 (UnetRes) and :908-1380 (ResidualDiffusion), so `Trainer.load/test/sample` and an
EMA wrapper work unchanged.  Underneath, `sample()` drives `UnetEngine` (hand-written sm_100a kernels through the C
ABI): DA-CLIP conditioning once per call, one CUDA graph replay per timestep, fused final_conv + update kernel.
There is no CPU path: calling `sample()` on CPU tensors raises.

Additive extension: `sample(..., noise=...)` injects host-supplied noise (dict with "init": (B,1,H,W) and, for
ancestral sampling, "steps": (T-1,B,1,H,W) ordered t = T-1 .. 1, or a callable t -> tensor) so that runs are
reproducible against the oracle.  Without it the same `torch.randn` calls as the reference are made on the device.

Only the configuration the reference ships (train.py:78-82: num_unet=1, objective='pred_res', condition=True,
input_condition=False) is implemented; anything else raises NotImplementedError.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, List, Optional, Union

import torch
import torch.nn.functional as F
from torch import nn

from . import ops, weights
from .daclip import DAClipEncoder
from .engine import UnetEngine
from .weights import UnetConfig


def _register(root: nn.Module, key: str, tensor: torch.Tensor, buffer: bool):
    """Create nested container modules so that `root.state_dict()` has exactly the reference key."""
    parts = key.split(".")
    m = root
    for p in parts[:-1]:
        if p not in m._modules:
            m.add_module(p, nn.Module())
        m = m._modules[p]
    if buffer:
        m.register_buffer(parts[-1], tensor)
    else:
        m.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=False))


class Unet(nn.Module):
    """Parameter container with the reference's `Unet` key layout (src/DADiff.py:530-683).
    Dead reference members (`clip_model.*`, the CLIP text tower, `prompt_learner`) are not instantiated."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=1, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, condition=False, input_condition=False,
                 seed: int = 10):
        super().__init__()
        if self_condition or learned_variance or learned_sinusoidal_cond or random_fourier_features or input_condition:
            raise NotImplementedError("only the shipped FoundDiff configuration is implemented (train.py:97-105)")
        if (init_dim not in (None, dim)) or (out_dim not in (None, channels)) or resnet_block_groups != 8:
            raise NotImplementedError("init_dim/out_dim/resnet_block_groups overrides are not implemented")
        self.cfg = UnetConfig(dim, tuple(dim_mults), channels)
        self.channels = channels
        self.out_dim = channels
        self.self_condition = False
        self.random_or_learned_sinusoidal_cond = False
        sd = weights.random_state_dict(seed, self.cfg)
        for k, v in sd.items():
            _register(self, k, v, buffer=("running_" in k or "num_batches_tracked" in k))


class UnetRes(nn.Module):
    """src/DADiff.py:743-836."""

    def __init__(self, dim, init_dim=None, out_dim=None, dim_mults=(1, 2, 4, 8), channels=1, self_condition=False,
                 resnet_block_groups=8, learned_variance=False, learned_sinusoidal_cond=False,
                 random_fourier_features=False, learned_sinusoidal_dim=16, num_unet=1, condition=False,
                 input_condition=False, objective='pred_res_noise', test_res_or_noise="res_noise", seed: int = 10):
        super().__init__()
        if num_unet != 1 or objective != 'pred_res':
            raise NotImplementedError("only num_unet=1 / objective='pred_res' is implemented (train.py:78-82)")
        self.condition = condition
        self.input_condition = input_condition
        self.channels = channels
        self.out_dim = channels
        self.random_or_learned_sinusoidal_cond = False
        self.self_condition = self_condition
        self.num_unet = num_unet
        self.objective = objective
        self.test_res_or_noise = test_res_or_noise
        self.unet0 = Unet(dim, init_dim=init_dim, out_dim=out_dim, dim_mults=dim_mults, channels=channels,
                          self_condition=self_condition, resnet_block_groups=resnet_block_groups,
                          learned_variance=learned_variance, learned_sinusoidal_cond=learned_sinusoidal_cond,
                          random_fourier_features=random_fourier_features, learned_sinusoidal_dim=learned_sinusoidal_dim,
                          condition=condition, input_condition=input_condition, seed=seed)
        self.compute_dtype = torch.bfloat16
        # storage of the residual stream / pre-GroupNorm conv outputs; None = automatic: fp16 under bf16 sampling (the
        # tensors whose rounding error accumulates get the 11-bit mantissa, the unnormalised ones keep bf16's range),
        # otherwise the compute dtype.  Set to torch.bfloat16 for pure-bf16 storage (misses the 1e-2 parity gate).
        self.trunk_dtype: Optional[torch.dtype] = None
        self._engines: Dict = {}
        self._daclip = None
        self._version = 0

    # -- engine management ----------------------------------------------------------------------------------
    def invalidate(self):
        """Call after changing weights in place; load_state_dict does it automatically."""
        self._engines.clear()
        self._daclip = None

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        """Accepts a reference checkpoint: live keys are loaded, the reference's dead members (`clip_model.*`,
        CLIP text tower, `prompt_learner`) are ignored.  Missing live keys still raise."""
        own = super().state_dict()
        live = {k: v for k, v in state_dict.items() if k in own}
        dead_ok = ("unet0.clip_model.", "unet0.dose_encoder.clip_model.", "unet0.dose_encoder.prompt_learner.")
        unexpected = [k for k in state_dict if k not in own and not k.startswith(dead_ok)]
        if strict and unexpected:
            raise RuntimeError(f"unexpected keys: {unexpected[:5]}...")
        res = super().load_state_dict(live, strict=strict, assign=assign)
        self.invalidate()
        return res

    def _live_sd(self):
        return OrderedDict((k, v.detach()) for k, v in self.unet0.state_dict().items())

    def engine(self, B, H, W, device) -> UnetEngine:
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())       # 'cuda' and 'cuda:0' are one engine
        trunk = self.trunk_dtype
        if trunk is None:
            trunk = torch.float16 if self.compute_dtype == torch.bfloat16 else self.compute_dtype
        key = (B, H, W, self.compute_dtype, trunk, str(device))
        eng = self._engines.get(key)
        if eng is None:
            self._engines.clear()          # one resident engine: activations for B=16 at 512^2 are ~25 GB
            eng = UnetEngine(self._live_sd(), self.unet0.cfg, B, H, W, dtype=self.compute_dtype, device=device,
                             trunk_dtype=trunk)
            self._engines[key] = eng
        return eng

    def daclip(self, device) -> DAClipEncoder:
        cdt = torch.float32 if self.compute_dtype == torch.float32 else torch.bfloat16
        if self._daclip is None or str(self._daclip_dev) != str(device) or self._daclip.conv_dtype != cdt:
            self._daclip = DAClipEncoder(self._live_sd(), device, conv_dtype=cdt)
            self._daclip_dev = device
        return self._daclip

    @torch.no_grad()
    def forward(self, x, time, x_self_cond=None):
        """Model-call boundary (src/DADiff.py:817-836, 1161-1164): x = cat(x_t, x_input) (B,2,H,W),
        time = [t_res, t_noise]; returns [pred (B,1,H,W)] (raw, un-clamped)."""
        if not x.is_cuda:
            raise RuntimeError("founddiff_b200 has no CPU path")
        t = time[0] if isinstance(time, (list, tuple)) else time
        B, _, H, W = x.shape
        eng = self.engine(B, H, W, x.device)
        eng.x_t.copy_(x[:, 0].reshape(B, -1))
        eng.x_input.copy_(x[:, 1].reshape(B, -1))
        dose, ctx = self.daclip(x.device).embed(x[:, 1:2])
        eng.set_condition(dose, ctx)
        eng.time.copy_(t.to(torch.float32).reshape(-1).expand(B))
        feat = eng.forward()
        out = torch.empty(B, H * W, 1, device=x.device, dtype=feat.dtype)
        w = eng.final_w.to(feat.dtype).reshape(1, 1, 1, -1).contiguous()
        ops.Conv(feat, w, out, B=B, Hin=H, Win=W, bias=eng.final_b, prefer_tc=False).run()
        return [out.float().reshape(B, 1, H, W)]


def make_schedule(timesteps: int = 1000, variant: str = "init") -> Dict[str, torch.Tensor]:
    """The 12 schedule buffers of ResidualDiffusion: `ctor` = __init__ (src/DADiff.py:946-1027), `init` = .init()
    (:1033-1118, what Trainer.test() uses, :1818).  They differ only at index 0 of alphas / betas2 / betas."""
    betas = torch.linspace(0.0001, 0.02, timesteps, dtype=torch.float32)
    alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
    alphas_cumsum = 1 - alphas_cumprod ** 0.5
    betas2_cumsum = 1 - alphas_cumprod
    alphas_cumsum_prev = F.pad(alphas_cumsum[:-1], (1, 0), value=1.)
    betas2_cumsum_prev = F.pad(betas2_cumsum[:-1], (1, 0), value=1.)
    alphas = alphas_cumsum - alphas_cumsum_prev
    betas2 = betas2_cumsum - betas2_cumsum_prev
    if variant == "init":
        alphas[0], betas2[0] = alphas[1].clone(), betas2[1].clone()
    else:
        alphas[0], betas2[0] = 0, 0
    betas_cumsum = torch.sqrt(betas2_cumsum)
    posterior_variance = betas2 * betas2_cumsum_prev / betas2_cumsum
    posterior_variance[0] = 0
    s = OrderedDict(
        alphas=alphas, alphas_cumsum=alphas_cumsum, one_minus_alphas_cumsum=1 - alphas_cumsum, betas2=betas2,
        betas=torch.sqrt(betas2), betas2_cumsum=betas2_cumsum, betas_cumsum=betas_cumsum,
        posterior_mean_coef1=betas2_cumsum_prev / betas2_cumsum,
        posterior_mean_coef2=(betas2 * alphas_cumsum_prev - betas2_cumsum_prev * alphas) / betas2_cumsum,
        posterior_mean_coef3=betas2 / betas2_cumsum, posterior_variance=posterior_variance,
        posterior_log_variance_clipped=torch.log(posterior_variance.clamp(min=1e-20)))
    s["posterior_mean_coef1"][0] = 0
    s["posterior_mean_coef2"][0] = 0
    s["posterior_mean_coef3"][0] = 1
    s["one_minus_alphas_cumsum"][-1] = 1e-6
    return OrderedDict((k, v.to(torch.float32)) for k, v in s.items())


class ResidualDiffusion(nn.Module):
    """src/DADiff.py:908-1380 (sampler half).  Training members (q_sample, p_losses, forward) are out of scope."""

    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, loss_type='l1',
                 objective='pred_res_noise', ddim_sampling_eta=0., condition=False, sum_scale=None,
                 input_condition=False, input_condition_mask=False, test_res_or_noise="None"):
        super().__init__()
        if objective != 'pred_res' or not condition or input_condition:
            raise NotImplementedError("only objective='pred_res', condition=True, input_condition=False is implemented")
        if timesteps != 1000:
            raise NotImplementedError("the reference's init() hard-codes 1000 timesteps (src/DADiff.py:1034)")
        self.model = model
        self.channels = model.channels
        self.self_condition = False
        self.image_size = image_size
        self.objective = objective
        self.condition = condition
        self.input_condition = input_condition
        self.input_condition_mask = input_condition_mask
        self.test_res_or_noise = test_res_or_noise
        self.sum_scale = sum_scale if sum_scale else 0.01          # :940-941
        self.ddim_sampling_eta = 0.                                # :942 (forced when condition=True)
        self.loss_type = loss_type
        self.num_timesteps = int(timesteps)
        self.sampling_timesteps = sampling_timesteps if sampling_timesteps is not None else timesteps
        assert self.sampling_timesteps <= timesteps
        self.is_ddim_sampling = self.sampling_timesteps < timesteps
        for k, v in make_schedule(timesteps, "ctor").items():
            self.register_buffer(k, v)
        self.use_cuda_graph = True
        self._graphs: Dict = {}

    def init(self):
        """src/DADiff.py:1033-1118: rebuilds the schedule (as plain CPU tensors in the reference)."""
        for k, v in make_schedule(1000, "init").items():
            setattr(self, k, v.to(self.betas.device))
        self.num_timesteps = 1000

    # -- helpers -------------------------------------------------------------------------------------------
    def _sched(self, name, t) -> float:
        """Schedule scalars are read from a host copy (no device sync inside the sampling loop)."""
        cache = self.__dict__.setdefault("_sched_host", {})
        buf = getattr(self, name)
        key = (name, buf.data_ptr(), buf._version)
        if cache.get("key_" + name) != key:
            cache[name] = buf.detach().float().cpu()
            cache["key_" + name] = key
        return float(cache[name][t])

    def _ddim_pairs(self):
        times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)   # :1287-1291
        times = list(reversed(times.int().tolist()))
        return list(zip(times[:-1], times[1:]))

    def _step_plan(self):
        """[(t, coef[6])] with coef = {c_xt, c_res, c_x0, c_noise, alphas_cumsum[t], betas_cumsum[t]}."""
        plan = []
        if self.is_ddim_sampling:
            for t, t_next in self._ddim_pairs():
                acs, bcs = self._sched("alphas_cumsum", t), self._sched("betas_cumsum", t)
                if t_next < 0:
                    plan.append((t, [0., 0., 1., 0., acs, bcs]))                                   # :1317-1321
                else:
                    import numpy as np
                    alpha = float(np.float32(acs) - np.float32(self._sched("alphas_cumsum", t_next)))  # :1323-1325 (fp32)
                    plan.append((t, [1., -alpha, 0., 0., acs, bcs]))                                # :1344, sigma2 = 0
        else:
            for t in reversed(range(self.num_timesteps)):                                           # :1254
                acs, bcs = self._sched("alphas_cumsum", t), self._sched("betas_cumsum", t)
                cn = float(torch.tensor(0.5 * self._sched("posterior_log_variance_clipped", t), dtype=torch.float32).exp()) if t > 0 else 0.   # :1228-1229
                plan.append((t, [self._sched("posterior_mean_coef1", t), self._sched("posterior_mean_coef2", t),
                                 self._sched("posterior_mean_coef3", t), cn, acs, bcs]))
        return plan

    # -- sampling --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, x_input=0, batch_size=16, last=True, *, noise=None, trace: Optional[list] = None):
        """src/DADiff.py:1367-1380.  x_input: list [ldct (B,1,H,W) in [0,1]] on a CUDA device.  Returns the list
        [x_input_plus_noise, denoised] (or every intermediate when last=False), each (B,1,H,W) in [0,1]."""
        if not isinstance(x_input, (list, tuple)):
            raise TypeError("condition=True: x_input must be a list [ldct] (src/DADiff.py:1371-1376)")
        ldct = x_input[0]
        if not ldct.is_cuda:
            raise RuntimeError("founddiff_b200 has no CPU path: pass CUDA tensors")
        B, C, H, W = ldct.shape
        assert C == 1
        dev = ldct.device
        model = self.model
        eng = model.engine(B, H, W, dev)
        P = H * W
        plan = self._step_plan()
        n_noise = 0 if self.is_ddim_sampling else len(plan) - 1

        def get_noise(kind, idx=None, t=None):
            if noise is None:
                return torch.randn(B, 1, H, W, device=dev)           # same call order as the reference (:1295, :1228)
            if kind == "init":
                return noise["init"].to(dev, torch.float32)
            st = noise["steps"]
            return (st(t) if callable(st) else st[idx]).to(dev, torch.float32)

        ldct32 = ldct.to(torch.float32).contiguous().view(B, P)
        first = torch.empty(B, P, device=dev, dtype=torch.float32)
        ops.sampler_init(ldct32, get_noise("init").contiguous().view(B, P), math.sqrt(self.sum_scale), eng.x_input, eng.x_t, first)
        dose, ctx = model.daclip(dev).embed(eng.x_input.view(B, 1, H, W))       # once per slice (cached across steps)
        eng.set_condition(dose, ctx)

        coef = self._buffer(eng, "coef", 8)
        noise_buf = self._buffer(eng, "noise", B * P).view(B, P)
        taps = None
        if trace is not None:
            taps = [self._buffer(eng, n, B * P).view(B, P) for n in ("pred_res", "pred_noise", "x_start")]
        coef_host = torch.tensor([[*c, 0., 0.] for _, c in plan], dtype=torch.float32).pin_memory()
        time_host = torch.tensor([self._sched("alphas_cumsum", t) for t, _ in plan], dtype=torch.float32) * self.num_timesteps  # :1162
        time_rows = time_host[:, None].expand(-1, B).contiguous().pin_memory()

        def one_step():
            eng.forward()
            ops.final_conv_update(eng.feat, eng.final_w, eng.final_b, eng.x_input, eng.x_t, noise_buf, coef, eng.x_t,
                                  *(taps if taps is not None else (None, None, None)))

        step_fn = one_step
        if self.use_cuda_graph:
            step_fn = self._graphed(eng, one_step, taps is not None)

        imgs = []
        for i, (t, c) in enumerate(plan):
            coef.copy_(coef_host[i], non_blocking=True)
            eng.time.copy_(time_rows[i], non_blocking=True)
            if c[3] != 0.:
                noise_buf.copy_(get_noise("step", i, t).contiguous().view(B, P))
            step_fn()
            if trace is not None:
                trace.append(dict(t=t, pred_res=taps[0].clone().view(B, 1, H, W), pred_noise=taps[1].clone().view(B, 1, H, W),
                                  x_start=taps[2].clone().view(B, 1, H, W)))
            if not last:
                o = torch.empty(B, P, device=dev, dtype=torch.float32)
                ops.unnormalize(eng.x_t, o)
                imgs.append(o.view(B, 1, H, W))
        if last:
            o = torch.empty(B, P, device=dev, dtype=torch.float32)
            ops.unnormalize(eng.x_t, o)
            imgs = [o.view(B, 1, H, W)]
        return [first.view(B, 1, H, W)] + imgs                                                        # :1354-1359

    # kept for API parity with the reference (they simply route to sample's two modes)
    def ddim_sample(self, x_input, shape, last=True, **kw):
        assert self.is_ddim_sampling
        return self._sample_normalised(x_input, last, **kw)

    def p_sample_loop(self, x_input, shape, last=True, **kw):
        assert not self.is_ddim_sampling
        return self._sample_normalised(x_input, last, **kw)

    def _sample_normalised(self, x_input, last, **kw):
        # the reference's ddim_sample / p_sample_loop take inputs already mapped to [-1, 1] (:1375)
        return self.sample([(x_input[0] + 1) * 0.5], last=last, **kw)

    @torch.no_grad()
    def model_predictions(self, x_input, x, t, x_input_condition=0, x_self_cond=None, clip_denoised=True):
        """src/DADiff.py:1153-1209, branch 'pred_res' — the per-step parity tap.  x_input, x: (B,1,H,W) in [-1,1];
        t: (B,) long (all equal).  Returns (pred_res, pred_noise, pred_x_start)."""
        from collections import namedtuple
        B, _, H, W = x.shape
        ti = int(t.reshape(-1)[0])
        time = (self.alphas_cumsum[ti] * self.num_timesteps).to(x.device).expand(B)
        out = self.model(torch.cat((x, x_input), dim=1), [time, time])[0]
        pred_res = out.clamp(-1., 1.) if clip_denoised else out
        acs, bcs = self._sched("alphas_cumsum", ti), self._sched("betas_cumsum", ti)
        pred_noise = (x - x_input - (acs - 1) * pred_res) / bcs
        x_start = x_input - pred_res
        if clip_denoised:
            x_start = x_start.clamp(-1., 1.)
        return namedtuple('ModelResPrediction', ['pred_res', 'pred_noise', 'pred_x_start'])(pred_res, pred_noise, x_start)

    # -- internals -------------------------------------------------------------------------------------------
    @staticmethod
    def _buffer(eng, name, n):
        store = eng.__dict__.setdefault("_sampler_bufs", {})
        t = store.get(name)
        if t is None or t.numel() < n:
            t = torch.zeros(n, device=eng.device, dtype=torch.float32)
            store[name] = t
        return t[:n]

    def _graphed(self, eng, fn, with_taps):
        """Capture one timestep (conditioning + Unet + fused final_conv/update) as a CUDA graph; the step's scalars
        (time, coefficients) and noise live in device buffers that are refreshed before each replay."""
        key = (id(eng), with_taps)
        g = self._graphs.get(key)
        if g is None:
            s = torch.cuda.Stream(device=eng.device)
            s.wait_stream(torch.cuda.current_stream(eng.device))
            with torch.cuda.stream(s):
                saved = eng.x_t.clone()
                fn()                                   # warm-up outside capture (lazy module loading, attribute sets)
                eng.x_t.copy_(saved)
            torch.cuda.current_stream(eng.device).wait_stream(s)
            torch.cuda.synchronize(eng.device)
            saved = eng.x_t.clone()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            eng.x_t.copy_(saved)
            self._graphs.clear()
            self._graphs[key] = (g, eng)
        else:
            g = g[0]
        return g.replay
