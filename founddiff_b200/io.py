"""Slice I/O of the reference's inference loop, batched.

* input: one `.npy` file per CT slice (raw detector units), `Normalize(min_value=-1000, max_value=2000)`
  (data/transforms.py:577-587): `clip(((m - 1024) - min) / (max - min), 0, 1)`; the dataset returns `[ndct, ldct]`
  (data/pdf_dataset.py:424-466).  The reference feeds batch 1; `SliceStream` stacks `batch` slices and double-buffers the
  host->device copy through pinned memory on a side stream so that the sampling graph is never starved.
* output: `np.save(name, img.reshape(H, W))` in [0, 1] (src/DADiff.py:1912-1915).
"""
from __future__ import annotations

import os
from typing import Iterable, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

HU_OFFSET = 1024.0
MIN_VALUE, MAX_VALUE = -1000.0, 2000.0


def normalize_hu(m: np.ndarray, min_value: float = MIN_VALUE, max_value: float = MAX_VALUE) -> np.ndarray:
    """data/transforms.py:582-587."""
    assert max_value > min_value
    m = m.astype(np.float32) - HU_OFFSET
    return np.clip((m - min_value) / (max_value - min_value), 0, 1).astype(np.float32)


def load_slices(paths: Sequence[str], normalize: bool = True) -> torch.Tensor:
    """(B, 1, H, W) fp32 in [0, 1], pinned when CUDA is available."""
    arrs = []
    for p in paths:
        a = np.load(p).astype(np.float32)
        if a.ndim != 2:
            a = a.reshape(a.shape[-2], a.shape[-1])
        arrs.append(normalize_hu(a) if normalize else a)
    t = torch.from_numpy(np.stack(arrs, axis=0)[:, None])
    return t.pin_memory() if torch.cuda.is_available() else t


def save_slices(paths: Sequence[str], images: torch.Tensor) -> None:
    """One `.npy` per slice, `(H, W)` float32 in [0, 1], as `np.save(npy_name, img.reshape(512, 512))` does."""
    imgs = images.detach().to("cpu", torch.float32).numpy()
    assert imgs.shape[0] == len(paths)
    for p, im in zip(paths, imgs):
        os.makedirs(os.path.dirname(os.path.abspath(p)), exist_ok=True)
        np.save(p, im.reshape(im.shape[-2], im.shape[-1]))


class SliceStream:
    """Iterates over `(ldct_paths[, ndct_paths])` in batches; yields `(ldct, ndct_or_None, batch_paths)` with the tensors
    already on `device`.  Batch i+1 is read, normalised and copied (pinned memory, side stream) while batch i is sampled."""

    def __init__(self, ldct_paths: Sequence[str], ndct_paths: Optional[Sequence[str]] = None, batch: int = 16, device="cuda"):
        assert ndct_paths is None or len(ndct_paths) == len(ldct_paths)
        self.ldct, self.ndct, self.batch, self.device = list(ldct_paths), (list(ndct_paths) if ndct_paths is not None else None), batch, torch.device(device)
        self._side = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None

    def __len__(self):
        return (len(self.ldct) + self.batch - 1) // self.batch

    def _stage(self, i: int):
        sl = slice(i * self.batch, (i + 1) * self.batch)
        host = [load_slices(self.ldct[sl])] + ([load_slices(self.ndct[sl])] if self.ndct is not None else [])
        if self._side is None:
            return host, self.ldct[sl], None
        with torch.cuda.stream(self._side):
            dev = [h.to(self.device, non_blocking=True) for h in host]
            ev = torch.cuda.Event()
            ev.record(self._side)
        return dev, self.ldct[sl], (ev, host)                # keep the pinned staging alive until the copy is done

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, Optional[torch.Tensor], List[str]]]:
        n = len(self)
        nxt = self._stage(0) if n else None
        for i in range(n):
            cur = nxt
            nxt = self._stage(i + 1) if i + 1 < n else None
            tensors, paths, sync = cur
            if sync is not None:
                torch.cuda.current_stream(self.device).wait_event(sync[0])
            yield tensors[0], (tensors[1] if len(tensors) > 1 else None), paths
