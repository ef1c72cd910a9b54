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

Objectives: the shipped configuration (train.py:78-82: num_unet=1, objective='pred_res') and the
reference's other three — 'pred_noise' (num_unet=1), 'pred_res_noise' and 'pred_x0_noise' (num_unet=2, train.py:75-77; two
Unet engines per timestep, test_res_or_noise in {"res_noise", "res", "noise"}) — all end in the same fused
final_conv + model_predictions + update kernel (`fd_final_conv_update_obj`).  `q_sample / p_losses / forward`
(:1382-1500) are forward-only mirrors (validation loss; no autograd).  condition=True, input_condition=False only.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Optional

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
        if num_unet not in (1, 2):
            raise ValueError("num_unet must be 1 or 2 (src/DADiff.py:775-815)")
        self.condition = condition
        self.input_condition = input_condition
        self.channels = channels
        self.out_dim = channels
        self.random_or_learned_sinusoidal_cond = False
        self.self_condition = self_condition
        self.num_unet = num_unet
        self.objective = objective
        self.test_res_or_noise = test_res_or_noise
        kw = dict(init_dim=init_dim, out_dim=out_dim, dim_mults=dim_mults, channels=channels,
                  self_condition=self_condition, resnet_block_groups=resnet_block_groups,
                  learned_variance=learned_variance, learned_sinusoidal_cond=learned_sinusoidal_cond,
                  random_fourier_features=random_fourier_features, learned_sinusoidal_dim=learned_sinusoidal_dim,
                  condition=condition, input_condition=input_condition)
        self.unet0 = Unet(dim, seed=seed, **kw)
        if num_unet == 2:
            self.unet1 = Unet(dim, seed=seed + 1, **kw)                 # :789-801 (an independently initialised twin)
        self.compute_dtype = torch.bfloat16
        # storage of the residual stream / pre-GroupNorm conv outputs; None = automatic: fp16 under bf16 sampling (the
        # tensors whose rounding error accumulates get the 11-bit mantissa, the unnormalised ones keep bf16's range),
        # otherwise the compute dtype.  Set to torch.bfloat16 for pure-bf16 storage (misses the 1e-2 parity gate).
        self.trunk_dtype: Optional[torch.dtype] = None
        self._engines: Dict = {}
        self._daclip = None
        self._built_fingerprint = None
        # nn.Module.load_state_dict recurses with _load_from_state_dict: hooks (not a load_state_dict override) see every path
        self._register_load_state_dict_pre_hook(self._drop_dead_keys)
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate())

    # -- engine management ----------------------------------------------------------------------------------
    def invalidate(self):
        """Drops the packed engines, the DA-CLIP encoders and (through the engine identity check in
        ResidualDiffusion._graphed) the captured CUDA graphs.  Runs automatically after load_state_dict on any path, after
        .to() and when a parameter's version counter moved."""
        self._engines.clear()
        self._daclip = None

    # Dead reference members a checkpoint may carry: dropped on load, on ANY load path.
    _DEAD = ("clip_model.", "dose_encoder.clip_model.", "dose_encoder.prompt_learner.")

    def _drop_dead_keys(self, state_dict, prefix, *_):
        """load_state_dict pre-hook: runs on the recursive path too (`Trainer.load` / EMA wrappers load through a PARENT module,
        which never calls a child's `load_state_dict` override), before this module's children look at their keys."""
        dead = tuple(f"{prefix}unet{i}.{d}" for i in range(self.num_unet) for d in self._DEAD)
        own = {prefix + k for k in self.state_dict().keys()}       # `dose_encoder.clip_model.visual.*` is live
        for k in [k for k in state_dict if k.startswith(dead) and k not in own]:
            del state_dict[k]

    def _apply(self, fn, *a, **kw):
        out = super()._apply(fn, *a, **kw)          # .to() / .cuda() / .half(): packed engine weights would go stale
        self.invalidate()
        return out

    def _param_fingerprint(self):
        """Detects in-place parameter edits (optimizer / EMA steps, manual `.copy_`) between sample() calls."""
        return sum(p._version for p in self.parameters()) + sum(b._version for b in self.buffers())

    def _unet(self, idx: int) -> Unet:
        return self.unet0 if idx == 0 else self.unet1

    def _live_sd(self, idx: int = 0):
        return OrderedDict((k, v.detach()) for k, v in self._unet(idx).state_dict().items())

    def engine(self, B, H, W, device, idx: int = 0) -> UnetEngine:
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())       # 'cuda' and 'cuda:0' are one engine
        trunk = self.trunk_dtype
        if trunk is None:
            trunk = torch.float16 if self.compute_dtype == torch.bfloat16 else self.compute_dtype
        key = (B, H, W, self.compute_dtype, trunk, str(device))
        fp = self._param_fingerprint()
        if fp != self._built_fingerprint:              # weights changed in place since the engines were packed
            self.invalidate()
            self._built_fingerprint = fp
        if self._engines and next(iter(self._engines))[0] != key:
            self._engines.clear()          # one resident shape: activations for B=16 at 512^2 are ~25 GB per Unet
        eng = self._engines.get((key, idx))
        if eng is None:
            eng = UnetEngine(self._live_sd(idx), self._unet(idx).cfg, B, H, W, dtype=self.compute_dtype, device=device,
                             trunk_dtype=trunk)
            self._engines[(key, idx)] = eng
        return eng

    def daclip(self, device, idx: int = 0) -> DAClipEncoder:
        """Each Unet owns its dose encoder (`unet{idx}.dose_encoder.*`)."""
        cdt = torch.float32 if self.compute_dtype == torch.float32 else torch.bfloat16
        if self._daclip is None:
            self._daclip = {}
        enc = self._daclip.get(idx)
        if enc is None or str(enc._fd_dev) != str(device) or enc.conv_dtype != cdt:
            enc = DAClipEncoder(self._live_sd(idx), device, conv_dtype=cdt)
            enc._fd_dev = device
            self._daclip[idx] = enc
        return enc

    def _eval_plan(self, objective: Optional[str] = None, test_res_or_noise: Optional[str] = None):
        """Which Unets one model call evaluates, with which entry of `time`, and the FD_OBJ_* branch that consumes the
        outputs: (kernel objective, [(unet idx, time idx), ...]).  src/DADiff.py:817-836 + :1168-1207."""
        objective = self.objective if objective is None else objective
        trn = self.test_res_or_noise if test_res_or_noise is None else test_res_or_noise
        if self.num_unet == 1:
            if objective == "pred_res":
                return "pred_res", [(0, 0)]
            if objective == "pred_noise":
                return "pred_noise", [(0, 1)]
            raise ValueError(f"objective {objective!r} needs num_unet=2 (src/DADiff.py:824-829)")
        if objective == "pred_res_noise":
            if trn == "res_noise":
                return "pred_res_noise", [(0, 0), (1, 1)]
            if trn == "res":
                return "pred_res", [(0, 0)]
            if trn == "noise":
                return "pred_noise", [(1, 1)]
            raise ValueError(f"test_res_or_noise {trn!r} (src/DADiff.py:818-823)")
        if objective == "pred_x0_noise":
            if trn != "res_noise":
                raise NotImplementedError("pred_x0_noise reads both model outputs (src/DADiff.py:1188-1192)")
            return "pred_x0_noise", [(0, 0), (1, 1)]
        raise ValueError(f"objective {objective!r} with num_unet=2: model_output[0] would be a pair (src/DADiff.py:1194-1207)")

    def _run_unet(self, idx, x, t):
        B, _, H, W = x.shape
        eng = self.engine(B, H, W, x.device, idx)
        eng.x_t.copy_(x[:, 0].reshape(B, -1))
        eng.x_input.copy_(x[:, 1].reshape(B, -1))
        dose, ctx = self.daclip(x.device, idx).embed(x[:, 1:2])
        eng.set_condition(dose, ctx)
        eng.time.copy_(t.to(torch.float32).reshape(-1).expand(B))
        feat = eng.forward()
        out = torch.empty(B, H * W, 1, device=x.device, dtype=feat.dtype)
        w = eng.final_w.to(feat.dtype).reshape(1, 1, 1, -1).contiguous()
        ops.Conv(feat, w, out, B=B, Hin=H, Win=W, bias=eng.final_b, prefer_tc=False).run()
        return out.float().reshape(B, 1, H, W)

    @torch.no_grad()
    def forward(self, x, time, x_self_cond=None):
        """Model-call boundary (src/DADiff.py:817-836, 1161-1164): x = cat(x_t, x_input) (B,2,H,W),
        time = [t_res, t_noise].  num_unet=1: returns [pred (B,1,H,W)] (raw, un-clamped); num_unet=2: the pair
        (unet0 output or 0, unet1 output or 0) selected by test_res_or_noise."""
        if not x.is_cuda:
            raise RuntimeError("founddiff_b200 has no CPU path")
        if not isinstance(time, (list, tuple)):
            time = [time, time]
        if self.num_unet == 1:
            _, evals = self._eval_plan()
        else:                                                        # :818-823 depends on test_res_or_noise only
            evals = {"res_noise": [(0, 0), (1, 1)], "res": [(0, 0)], "noise": [(1, 1)]}[self.test_res_or_noise]
        outs = {idx: self._run_unet(idx, x, time[ti]) for idx, ti in evals}
        if self.num_unet == 1:
            return [outs[0]]
        return outs.get(0, 0), outs.get(1, 0)


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
    """src/DADiff.py:908-1500: the sampler, plus forward-only `q_sample / p_losses / forward` (no autograd; training
    itself — optimiser, EMA, Trainer — is out of scope)."""

    def __init__(self, model, *, image_size, timesteps=1000, sampling_timesteps=None, loss_type='l1',
                 objective='pred_res_noise', ddim_sampling_eta=0., condition=False, sum_scale=None,
                 input_condition=False, input_condition_mask=False, test_res_or_noise="None"):
        super().__init__()
        if not condition or input_condition:
            raise NotImplementedError("only condition=True, input_condition=False is implemented (train.py:78-119)")
        if objective not in ops.OBJECTIVES:
            raise ValueError(f"unknown objective {objective}")                    # :1468
        model._eval_plan(objective, test_res_or_noise if model.num_unet == 2 else None)   # raises on combinations the reference cannot run
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
        # num_unet = 2: evaluate the two Unets of a timestep on two streams (fork / join inside the step's CUDA graph)
        self.concurrent_unets = False
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
        """[(t, coef[7])] with coef = {c_xt, c_res, c_x0, c_noise, alphas_cumsum[t], betas_cumsum[t],
        one_minus_alphas_cumsum[t]}."""
        plan = []
        if self.is_ddim_sampling:
            for t, t_next in self._ddim_pairs():
                acs, bcs = self._sched("alphas_cumsum", t), self._sched("betas_cumsum", t)
                if t_next < 0:
                    plan.append((t, [0., 0., 1., 0., acs, bcs, self._sched("one_minus_alphas_cumsum", t)]))   # :1317-1321
                else:
                    import numpy as np
                    alpha = float(np.float32(acs) - np.float32(self._sched("alphas_cumsum", t_next)))  # :1323-1325 (fp32)
                    plan.append((t, [1., -alpha, 0., 0., acs, bcs, self._sched("one_minus_alphas_cumsum", t)]))  # :1344, sigma2 = 0
        else:
            for t in reversed(range(self.num_timesteps)):                                           # :1254
                acs, bcs = self._sched("alphas_cumsum", t), self._sched("betas_cumsum", t)
                cn = float(torch.tensor(0.5 * self._sched("posterior_log_variance_clipped", t), dtype=torch.float32).exp()) if t > 0 else 0.   # :1228-1229
                plan.append((t, [self._sched("posterior_mean_coef1", t), self._sched("posterior_mean_coef2", t),
                                 self._sched("posterior_mean_coef3", t), cn, acs, bcs,
                                 self._sched("one_minus_alphas_cumsum", t)]))
        return plan

    # -- sampling --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def sample(self, x_input=0, batch_size=16, last=True, *, noise=None, trace: Optional[list] = None,
               steps_limit: Optional[int] = None, _record=None):
        """src/DADiff.py:1367-1380.  x_input: list [ldct (B,1,H,W) in [0,1]] on a CUDA device.  Returns the list
        [x_input_plus_noise, denoised] (or every intermediate when last=False), each (B,1,H,W) in [0,1].
        `steps_limit` (measurement only): stop after the first n timesteps of the schedule (a window of the ancestral loop)."""
        if not isinstance(x_input, (list, tuple)):
            raise TypeError("condition=True: x_input must be a list [ldct] (src/DADiff.py:1371-1376)")
        ldct = x_input[0]
        if not ldct.is_cuda:
            raise RuntimeError("founddiff_b200 has no CPU path: pass CUDA tensors")
        B, C, H, W = ldct.shape
        assert C == 1
        dev = ldct.device
        model = self.model
        objective, evals = model._eval_plan(self.objective, self.test_res_or_noise if model.num_unet == 2 else None)
        model._param_fingerprint() == model._built_fingerprint or model.invalidate()
        live = {id(e) for e in model._engines.values()}
        if any(id(e) not in live for _, es in self._graphs.values() for e in es):
            self._graphs.clear()                     # graphs of dropped engines (and the ~23 GB of buffers they pin) go first
        engs = [model.engine(B, H, W, dev, idx) for idx, _ in evals]
        eng = engs[0]
        for other in engs[1:]:                       # both Unets read the same images (UnetRes.forward, :817-820)
            other.x_t, other.x_input = eng.x_t, eng.x_input
        P = H * W
        plan = self._step_plan()
        if steps_limit is not None:
            plan = plan[:max(1, int(steps_limit))]

        def get_noise(kind, idx=None, t=None):
            if noise is None:
                return torch.randn(B, 1, H, W, device=dev)           # same call order as the reference (:1295, :1228)
            if kind == "init":
                return noise["init"].to(dev, torch.float32, non_blocking=True)
            st = noise["steps"]
            return st(t) if callable(st) else st[idx]

        main = torch.cuda.current_stream(dev)
        h2d = {"stream": None, "stage": None, "free": None}

        def put_step_noise(src):
            """Step noise -> the graph's noise buffer.  Host tensors (north_star: host-supplied noise) go through one device
            staging buffer on a copy stream, so the H2D of step k+1 runs under step k's kernels; the host tensor is free for
            its producer when this returns."""
            if src.is_cuda:
                noise_buf.copy_(src.to(torch.float32).reshape(B, P))
                return
            if h2d["stream"] is None:
                h2d["stream"] = self._side_stream(dev, "h2d")
                h2d["stage"] = self._buffer(eng, "noise_stage", B * P).view(B, P)
                h2d["free"] = torch.cuda.Event()
                h2d["free"].record(main)
            cs = h2d["stream"]
            cs.wait_event(h2d["free"])                   # the previous step's copy out of the staging buffer is done
            with torch.cuda.stream(cs):
                h2d["stage"].copy_(src.to(torch.float32).reshape(B, P), non_blocking=True)
            cs.synchronize()
            noise_buf.copy_(h2d["stage"])
            h2d["free"].record(main)

        ldct32 = ldct.to(torch.float32).contiguous().view(B, P)
        first = torch.empty(B, P, device=dev, dtype=torch.float32)
        ops.sampler_init(ldct32, get_noise("init").contiguous().view(B, P), math.sqrt(self.sum_scale), eng.x_input, eng.x_t, first)
        for e, (idx, _) in zip(engs, evals):         # once per slice and per Unet (cached across timesteps)
            with ops.nvtx_range("fd.sample.daclip (once per slice)"):
                dose, ctx = model.daclip(dev, idx).embed(eng.x_input.view(B, 1, H, W))
                e.set_condition(dose, ctx)

        coef = self._buffer(eng, "coef", 8)
        noise_buf = self._buffer(eng, "noise", B * P).view(B, P)
        taps = None
        if trace is not None:
            taps = [self._buffer(eng, n, B * P).view(B, P) for n in ("pred_res", "pred_noise", "x_start")]
        coef_host = torch.tensor([[*c, 0.] for _, c in plan], dtype=torch.float32).pin_memory()
        # model time arguments (:1161-1163): [alphas_cumsum[t], betas_cumsum[t]] * num_timesteps, one row per Unet evaluated
        time_rows = []
        for _, ti in evals:
            name = ("alphas_cumsum", "betas_cumsum")[ti]
            th = torch.tensor([self._sched(name, t) for t, _ in plan], dtype=torch.float32) * self.num_timesteps
            time_rows.append(th[:, None].expand(-1, B).contiguous().pin_memory())
        two = len(engs) == 2
        side = self._side_stream(dev) if (two and self.concurrent_unets) else None

        def one_step():
            if side is not None:
                cur = torch.cuda.current_stream(dev)
                side.wait_stream(cur)
                with torch.cuda.stream(side):
                    engs[1].forward()
                eng.forward()
                cur.wait_stream(side)
            else:
                for e in engs:
                    e.forward()
            ops.final_conv_update(eng.feat, eng.final_w, eng.final_b, eng.x_input, eng.x_t, noise_buf, coef, eng.x_t,
                                  *(taps if taps is not None else (None, None, None)), objective=objective,
                                  feat1=engs[1].feat if two else None, w1=engs[1].final_w if two else None,
                                  bias1=engs[1].final_b if two else None)

        step_fn = one_step
        if self.use_cuda_graph and _record is None:
            step_fn = self._graphed(engs, one_step, (taps is not None, objective, side is not None))
        if _record is not None:                      # program.export_step_program: the first timestep runs eagerly under the recorder
            if two or taps is not None:
                raise NotImplementedError("step programs are recorded for the one-Unet objectives without taps")
            for t_, n_ in ((eng.x_t, "x_t"), (eng.x_input, "x_input"), (eng.time, "time"), (coef, "coef"), (noise_buf, "noise"),
                           (eng.prompt_emb, "prompt_emb"), (eng.feat, "feat")):
                _record.name(t_, n_)
            for j, lv in enumerate(eng._local_views):
                _record.name(lv.dense(), f"local.{j}")

            def step_fn():
                if _record.calls:
                    return one_step()
                with _record:
                    eng.forward()
                    _record.mark_unet_done()
                    ops.final_conv_update(eng.feat, eng.final_w, eng.final_b, eng.x_input, eng.x_t, noise_buf, coef, eng.x_t,
                                          None, None, None, objective=objective)

        imgs = []
        for i, (t, c) in enumerate(plan):
            coef.copy_(coef_host[i], non_blocking=True)
            for e, rows in zip(engs, time_rows):
                e.time.copy_(rows[i], non_blocking=True)
            if c[3] != 0.:
                put_step_noise(get_noise("step", i, t))
            with ops.nvtx_range(f"fd.sample.step t={t}"):
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
        """src/DADiff.py:1153-1209, every objective branch — the per-step parity tap.  x_input, x: (B,1,H,W) in [-1,1];
        t: (B,) long.  Returns (pred_res, pred_noise, pred_x_start)."""
        from collections import namedtuple
        ex = lambda name: getattr(self, name).to(x.device)[t.reshape(-1)].view(-1, 1, 1, 1)      # noqa: E731  extract(), :841-844
        acs, bcs, omacs = ex("alphas_cumsum"), ex("betas_cumsum"), ex("one_minus_alphas_cumsum")
        out = self.model(torch.cat((x, x_input), dim=1),
                         [acs.reshape(-1) * self.num_timesteps, bcs.reshape(-1) * self.num_timesteps], x_self_cond)
        clip = (lambda v: v.clamp(-1., 1.)) if clip_denoised else (lambda v: v)
        objective, _ = self.model._eval_plan(self.objective, self.test_res_or_noise if self.model.num_unet == 2 else None)
        if objective == "pred_res":                                                   # :1176-1182, 1202-1207
            pred_res = clip(out[0])
            pred_noise = (x - x_input - (acs - 1) * pred_res) / bcs                   # :1120-1124
            x_start = clip(x_input - pred_res)
        elif objective == "pred_noise":                                               # :1183-1189, 1194-1201
            pred_noise = out[1] if self.model.num_unet == 2 else out[0]
            x_start = clip((x - acs * x_input - bcs * pred_noise) / omacs)            # :1126-1130
            pred_res = clip(x_input - x_start)
        elif objective == "pred_res_noise":                                           # :1169-1175
            pred_res, pred_noise = clip(out[0]), out[1]
            x_start = clip(x - acs * pred_res - bcs * pred_noise)                     # :1132-1136
        else:                                                                         # pred_x0_noise :1188-1192
            pred_res, pred_noise, x_start = clip(x_input - out[0]), out[1], clip(out[0])
        return namedtuple('ModelResPrediction', ['pred_res', 'pred_noise', 'pred_x_start'])(pred_res, pred_noise, x_start)

    # -- forward-only training-side members (src/DADiff.py:1382-1500) ------------------------------------------
    def q_sample(self, x_start, x_res, t, noise=None):
        """:1382-1388  x_t = x_start + alphas_cumsum[t] x_res + betas_cumsum[t] noise."""
        if noise is None:
            noise = torch.randn_like(x_start)
        ex = lambda name: getattr(self, name).to(x_start.device)[t.reshape(-1)].view(-1, 1, 1, 1)  # noqa: E731
        return x_start + ex("alphas_cumsum") * x_res + ex("betas_cumsum") * noise

    @property
    def loss_fn(self):
        if self.loss_type == 'l1':
            return F.l1_loss
        if self.loss_type == 'l2':
            return F.mse_loss
        raise ValueError(f'invalid loss type {self.loss_type}')

    @torch.no_grad()
    def p_losses(self, imgs, t, noise=None):
        """:1399-1482 without autograd: imgs = [x_start (gt), x_input] in [-1,1]; t (B,) long; returns the loss list
        (one entry per model output).  The Unet evaluations run on the kernel engine."""
        if not isinstance(imgs, (list, tuple)):
            raise TypeError("condition=True: imgs must be the list [gt, input] (:1400-1406)")
        x_start, x_input = imgs[0], imgs[1]
        if noise is None:
            noise = torch.randn_like(x_start)
        x_res = x_input - x_start
        x = self.q_sample(x_start, x_res, t, noise=noise)
        ex = lambda name: getattr(self, name).to(x.device)[t.reshape(-1)]                            # noqa: E731
        model_out = self.model(torch.cat((x, x_input), dim=1),
                               [ex("alphas_cumsum") * self.num_timesteps, ex("betas_cumsum") * self.num_timesteps], None)
        target = {"pred_res_noise": [x_res, noise], "pred_x0_noise": [x_start, noise], "pred_noise": [noise],
                  "pred_res": [x_res]}[self.objective]                                               # :1444-1466
        losses = []
        for out, tgt in zip(model_out, target):                                                      # :1476-1482
            if not torch.is_tensor(out):
                raise ValueError("p_losses needs every model output (test_res_or_noise='res_noise' with num_unet=2)")
            loss = self.loss_fn(out, tgt, reduction='none')
            losses.append(loss.reshape(loss.shape[0], -1).mean(dim=1).mean())
        return losses

    def forward(self, img, *args, **kwargs):
        """:1484-1500: draws t ~ U{0..T-1} per sample, maps [gt, input] from [0,1] to [-1,1], returns p_losses."""
        b, device = img[0].shape[0], img[0].device
        t = torch.randint(0, self.num_timesteps, (b,), device=device).long()
        img = [i * 2 - 1 for i in img]
        return self.p_losses(img, t, *args, **kwargs)

    # -- internals -------------------------------------------------------------------------------------------
    @staticmethod
    def _buffer(eng, name, n):
        store = eng.__dict__.setdefault("_sampler_bufs", {})
        t = store.get(name)
        if t is None or t.numel() < n:
            t = torch.zeros(n, device=eng.device, dtype=torch.float32)
            store[name] = t
        return t[:n]

    def _side_stream(self, dev, name: str = "unet1"):
        st = self.__dict__.setdefault("_side_streams", {})
        key = (str(dev), name)
        if key not in st:
            st[key] = torch.cuda.Stream(device=dev)
        return st[key]

    def _graphed(self, engs, fn, variant):
        """Capture one timestep (conditioning + Unet(s) + fused final_conv/update) as a CUDA graph; the step's scalars
        (time, coefficients) and noise live in device buffers that are refreshed before each replay."""
        eng = engs[0]
        key = (tuple(id(e) for e in engs), variant)
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
            self._graphs[key] = (g, list(engs))
        else:
            g = g[0]
        return g.replay
