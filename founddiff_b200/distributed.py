"""Multi-GPU sampling: slices are independent chains, so a global batch is split contiguously over ranks (one process per
GPU, weights replicated), each rank samples its shard with no per-step communication, and ONE NCCL all_gather of the
final (B/W, 1, H, W) fp32 slices assembles the result.

Noise is host-supplied (north_star) and drawn from ONE GENERATOR PER GLOBAL SLICE INDEX (seed, slice) — every slice sees
the same numbers whatever the world size or the shard it lands in, and no rank ever draws (or holds) more than its own
shard: `SliceNoise`.  For ancestral sampling the per-step tensors are produced by a background thread into a small ring of
pinned buffers (`SliceNoise.steps`), never materialised as a (T, B, 1, H, W) tensor.

The reference has no multi-GPU inference at all (train.py:162-165 runs test() on the local main process, batch 1).
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Sequence

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


class SliceNoise:
    """Host noise for the slices `ids` (GLOBAL slice indices) of a batch: slice i owns the generator seeded with
    seed * 1000003 + i and draws, in this order, its init tensor and then one tensor per ancestral step (t = T-1 .. 1).
    Two runs that give a slice the same global index get bit-identical noise for it, on any rank layout."""

    def __init__(self, seed: int, ids: Sequence[int], shape, pin: bool = True):
        self.ids, self.shape = list(ids), tuple(shape)
        self.gens = [torch.Generator(device="cpu").manual_seed(int(seed) * 1000003 + int(i)) for i in self.ids]
        self.pin = pin and torch.cuda.is_available()
        self._pool = ThreadPoolExecutor(max_workers=max(1, min(8, len(self.ids)))) if len(self.ids) > 1 else None

    def _buf(self):
        t = torch.empty(len(self.ids), *self.shape, dtype=torch.float32)
        return t.pin_memory() if self.pin else t

    def _draw(self, out: torch.Tensor) -> torch.Tensor:
        def one(j):
            torch.randn(self.shape, generator=self.gens[j], out=out[j])
        if self._pool is None:
            for j in range(len(self.ids)):
                one(j)
        else:                               # ATen releases the GIL: the slices' generators advance in parallel
            list(self._pool.map(one, range(len(self.ids))))
        return out

    def init(self) -> torch.Tensor:
        return self._draw(self._buf())

    def steps(self, n_steps: int, depth: int = 3):
        """Callable t -> (n, *shape) pinned tensor for the ancestral loop; must be called once per step, in loop order.
        A producer thread stays `depth` steps ahead (16 slices of 512^2 take ~10 ms of host time per step against ~37 ms
        of device time), so the host RNG never stalls the sampler and at most `depth + 2` tensors exist at any time.
        The tensor returned for step i stays valid until the call for step i + 1 (ring of depth + 2: `depth` queued, one
        being drawn, one with the consumer)."""
        ring = [self._buf() for _ in range(depth + 2)]
        q: "queue.Queue" = queue.Queue(maxsize=depth)

        def produce():
            for i in range(n_steps):
                q.put(self._draw(ring[i % len(ring)]))
        th = threading.Thread(target=produce, daemon=True)
        th.start()
        return lambda t: q.get()


def slice_noise(seed: int, a: int, b: int, shape, steps: int = 0):
    """Materialised form of `SliceNoise` for the global slices [a, b): (init (b-a, *shape), steps (steps, b-a, *shape) | None)."""
    sn = SliceNoise(seed, range(a, b), shape, pin=False)
    init = sn.init()
    st = torch.stack([sn._draw(torch.empty(b - a, *shape)) for _ in range(steps)]) if steps else None
    return init, st


def global_noise(n_global: int, shape, seed: int = 4321, steps: int = 0):
    """init (n_global, *shape) [+ steps (steps, n_global, *shape)] for the WHOLE batch (tests, single-process use):
    identical, slice by slice, to what each rank's `SliceNoise` draws for its shard."""
    return slice_noise(seed, 0, n_global, shape, steps)


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
    gathered (B_global, 1, H, W) denoised slices on every rank.  A rank draws noise for ITS slices only (per-slice
    generators, see SliceNoise), per step and ahead of the device for ancestral sampling."""
    rank, ws = world()
    device = device or torch.device("cuda", torch.cuda.current_device())
    n = ldct_global.shape[0]
    if n < ws:
        raise ValueError(f"sample_sharded: {n} slices cannot be split over {ws} ranks (every rank must take part in the gather)")
    a, b = shard_range(n, rank, ws)
    sn = SliceNoise(noise_seed, range(a, b), tuple(ldct_global.shape[1:]))
    noise = {"init": sn.init()}
    if not diffusion.is_ddim_sampling:
        noise["steps"] = sn.steps(diffusion.num_timesteps - 1)
    out = diffusion.sample([ldct_global[a:b].to(device, non_blocking=True)], batch_size=b - a, last=last, noise=noise)
    return gather_slices(out[-1], n)
