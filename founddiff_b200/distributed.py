"""Multi-GPU sampling: slices are independent chains, so a global batch is split contiguously over ranks (one process per
GPU, weights replicated), each rank samples its shard with no per-step communication, and ONE NCCL all_gather of the
final (B/W, 1, H, W) fp32 slices assembles the result.  Noise is drawn once for the GLOBAL batch from a
seeded host generator and sliced by rank, so results do not depend on the world size.

The reference has no multi-GPU inference at all (train.py:162-165 runs test() on the local main process, batch 1).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_global: int, rank: int, world_size: int):
    """Contiguous split; the first (n_global % world_size) ranks get one extra slice."""
    base, rem = divmod(n_global, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard(x: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    a, b = shard_range(x.shape[0], rank, world_size)
    return x[a:b]


def global_noise(n_global: int, shape, seed: int = 4321, steps: int = 0):
    """init (n_global, *shape) [+ steps (steps, n_global, *shape)] from ONE host generator."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    init = torch.randn(n_global, *shape, generator=g)
    st = torch.randn(steps, n_global, *shape, generator=g) if steps else None
    return init, st


def gather_slices(local: torch.Tensor, n_global: Optional[int] = None) -> torch.Tensor:
    """all_gather of the final denoised slices.  Equal shards use all_gather_into_tensor (one NCCL call over
    NVLink/NVSwitch); ragged shards are padded to the largest shard and trimmed."""
    rank, ws = world()
    if ws == 1:
        return local
    local = local.contiguous()
    n_global = n_global if n_global is not None else local.shape[0] * ws
    sizes = [shard_range(n_global, r, ws) for r in range(ws)]
    mx = max(b - a for a, b in sizes)
    if all(b - a == mx for a, b in sizes):
        out = torch.empty(ws * mx, *local.shape[1:], device=local.device, dtype=local.dtype)
        dist.all_gather_into_tensor(out, local)
        return out
    pad = torch.zeros(mx, *local.shape[1:], device=local.device, dtype=local.dtype)
    pad[: local.shape[0]] = local
    out = torch.empty(ws * mx, *local.shape[1:], device=local.device, dtype=local.dtype)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx: r * mx + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


def sample_sharded(diffusion, ldct_global: torch.Tensor, *, noise_seed: int = 4321, last: bool = True, device=None):
    """Drop-in multi-GPU `sample()`: every rank passes the same global batch (host or device tensor); returns the
    gathered (B_global, 1, H, W) denoised slices on every rank."""
    rank, ws = world()
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = ldct_global.shape[0]
    a, b = shard_range(n, rank, ws)
    n_steps = 0 if diffusion.is_ddim_sampling else diffusion.num_timesteps - 1
    init, steps = global_noise(n, tuple(ldct_global.shape[1:]), noise_seed, n_steps)
    noise = {"init": init[a:b]}
    if steps is not None:
        noise["steps"] = steps[:, a:b]
    out = diffusion.sample([ldct_global[a:b].to(device)], batch_size=b - a, last=last, noise=noise)
    return gather_slices(out[-1], n)
