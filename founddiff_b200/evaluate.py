"""Batched mirror of `Trainer.test` (src/DADiff.py:1817-1920): load low-dose `.npy` slices, `ema_model.init()`,
`sample([ldct], last=True)[-1]`, PSNR / SSIM / RMSE against the normal-dose slices on the device, write `.npy` results.
Checkpoints saved by the reference (`Trainer.save`, src/DADiff.py:1630-1646: keys 'step', 'model', 'ema', ...) are ingested
with `load_reference_checkpoint` (the EMA weights are what `Trainer.test` samples with, :1865-1869)."""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import metrics
from .io import SliceStream, save_slices
from .weights import extract_live_weights


def load_reference_checkpoint(diffusion, checkpoint, prefer_ema: bool = True, unsafe_pickle: bool = False) -> Dict[str, int]:
    """checkpoint: path to `model-<milestone>.pt` or the loaded dict.  Loads the live denoiser + DA-CLIP weights into
    `diffusion.model` (dead `clip_model.*` / text tower / `perceploss.*` entries are dropped).
    `Trainer.save` files hold tensors, ints and dicts only, so the file is read with `weights_only=True`; `unsafe_pickle=True`
    is the explicit opt-in for files that need the full unpickler (arbitrary code execution: trusted files only)."""
    data = (torch.load(checkpoint, map_location="cpu", weights_only=not unsafe_pickle)
            if isinstance(checkpoint, (str, os.PathLike)) else checkpoint)
    sd = None
    if isinstance(data, dict):
        if prefer_ema and "ema" in data:
            sd = data["ema"]
        elif "model" in data:
            sd = data["model"]
    if sd is None:
        sd = data
    merged, n = {}, 0
    for i in range(getattr(diffusion.model, "num_unet", 1)):          # num_unet = 2 checkpoints carry unet0.* and unet1.*
        live = extract_live_weights(sd, unet=i)
        merged.update({f"unet{i}." + k: v for k, v in live.items()})
        n += len(live)
    diffusion.model.load_state_dict(merged)
    return {"loaded": n, "step": int(data.get("step", -1)) if isinstance(data, dict) else -1}


@torch.no_grad()
def evaluate(diffusion, ldct_paths: Sequence[str], ndct_paths: Optional[Sequence[str]] = None, out_dir: Optional[str] = None,
             batch: int = 16, device="cuda", noise_seed: Optional[int] = None) -> Dict[str, List[float]]:
    """Returns per-slice lists {'psnr', 'ssim', 'rmse'} (empty without `ndct_paths`) — what the reference accumulates in
    `test_running_{psnr,ssim,rmse}` — and writes `<out_dir>/<name>.npy` per slice when `out_dir` is given."""
    diffusion.init()
    res: Dict[str, List[float]] = {"psnr": [], "ssim": [], "rmse": []}
    gen = torch.Generator(device="cpu").manual_seed(noise_seed) if noise_seed is not None else None
    for ldct, ndct, paths in SliceStream(ldct_paths, ndct_paths, batch=batch, device=device):
        noise = None
        if gen is not None and diffusion.is_ddim_sampling:
            noise = {"init": torch.randn(ldct.shape, generator=gen)}
        pred = diffusion.sample([ldct], batch_size=ldct.shape[0], last=True, noise=noise)[-1]
        if ndct is not None:
            p, s, r = metrics.slice_metrics(pred, ndct)
            res["psnr"] += p.tolist()
            res["ssim"] += s.tolist()
            res["rmse"] += r.tolist()
        if out_dir is not None:
            save_slices([os.path.join(out_dir, os.path.basename(q)[:-4] + ".npy") for q in paths], pred)
    return res
