"""Step programs: record ONE sampling timestep of the engine into a plan file that the C runtime replays without Python.

    path = export_step_program(diffusion, ldct, "plan.fdp", noise=...)      # once per (batch, H, W, storage type)
    # C / C++ host (include/founddiff_b200.h):  fd_program_load -> fd_program_buffer("x_t" ...) -> fd_sample_step(prog, stream)
    prog = StepProgram(path)                                                # the same runtime through ctypes, on a torch arena

A timestep (`Unet.forward`, src/DADiff.py:685-740, + `model_predictions` and the update, :1153-1209, 1221-1230, 1323-1344) is a
fixed list of launches over fixed buffers.  `Recorder` intercepts the C-ABI calls `ops` makes while one step runs eagerly:
every launch is stored with its scalar arguments and its pointer arguments as (allocation, byte offset); allocations that hold
packed weights (made by `engine.upload`) and small host-prepared tensors are stored with their bytes, activations are not.
The file layout is the one csrc/fd_program.cu reads."""
from __future__ import annotations

import bisect
import ctypes
import struct
from typing import Dict, List, Optional

import torch

from . import _lib, gen_dispatch, ops

MAGIC = b"FDPROG2\0"
MAX_ARGS = 40
A_INT, A_LONG, A_FLOAT, A_PTR, A_NULL, A_STREAM = range(6)
OP_CALL, OP_CONV, OP_MEMSET = range(3)
SMALL = 1 << 20                      # allocations up to 1 MiB are stored with their contents (conditioning vectors, schedules)
_KINDS: Optional[Dict[str, str]] = None


def _kinds() -> Dict[str, str]:
    global _KINDS
    if _KINDS is None:
        _KINDS = {name: "".join(k for _, k in plist) for name, plist in gen_dispatch.parse()}
    return _KINDS


class _LibProxy:
    """Forwards every attribute to the loaded library; calls of recordable entry points are also appended to the recorder."""

    def __init__(self, lib, rec):
        self._lib, self._rec = lib, rec

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name not in _kinds():
            return fn

        def call(*args):
            self._rec.calls.append((OP_CALL, name, args))
            return fn(*args)
        return call


class Recorder:
    def __init__(self):
        self.calls: List = []
        self.tensors: Dict[int, torch.Tensor] = {}        # data_ptr -> tensor (kept alive)
        self.names: Dict[int, str] = {}                   # storage ptr -> name
        self.unet_ops: Optional[int] = None

    # -- hooks used by ops ---------------------------------------------------------------------------------
    def note(self, t):
        if not isinstance(t, torch.Tensor):
            t = getattr(t, "t", None)                     # engine._view_ptr wraps a strided view
        if isinstance(t, torch.Tensor) and t.is_cuda:
            self.tensors[t.data_ptr()] = t

    def name(self, t: torch.Tensor, name: str):
        self.note(t)
        self.names[t.untyped_storage().data_ptr()] = name

    def conv(self, conv):
        for t in conv._keep + (conv._w4,):
            if t is not None:
                self.note(t)
        self.calls.append((OP_CONV, "conv", conv.params))

    def memset(self, t: torch.Tensor):
        self.note(t)
        self.calls.append((OP_MEMSET, "memset", (t.data_ptr(), t.numel() * t.element_size())))

    def mark_unet_done(self):
        self.unet_ops = len(self.calls)

    def __enter__(self):
        real = _lib.load()
        ops._REC = self
        _lib.PROXY = _LibProxy(real, self)
        return self

    def __exit__(self, *exc):
        ops._REC = None
        _lib.PROXY = None
        return False

    # -- serialisation -------------------------------------------------------------------------------------
    def save(self, path: str) -> str:
        torch.cuda.synchronize()
        storages: Dict[int, torch.Tensor] = {}            # storage base -> a tensor on it
        for t in self.tensors.values():
            storages.setdefault(t.untyped_storage().data_ptr(), t)
        bases = sorted(storages)
        sizes = [storages[b].untyped_storage().nbytes() for b in bases]

        def locate(ptr: int):
            i = bisect.bisect_right(bases, ptr) - 1
            if i < 0 or ptr >= bases[i] + max(sizes[i], 1):
                raise RuntimeError(f"step program: pointer {ptr:#x} is not inside any tensor the recorder saw")
            return i, ptr - bases[i]

        allocs, blobs, arena, data_off = [], [], 0, 0
        for b, n in zip(bases, sizes):
            t = storages[b]
            keep = b in ops.CONST_STORAGES or n <= SMALL
            raw = b""
            if keep:
                flat = torch.empty(0, dtype=torch.uint8, device=t.device).set_(t.untyped_storage(), 0, (n,), (1,))
                raw = flat.cpu().numpy().tobytes()
            allocs.append(struct.pack("<QQQII48s", n, arena, data_off if keep else 0, 1 if keep else 0, 0,
                                      self.names.get(b, "").encode()[:47]))
            blobs.append(raw)
            data_off += len(raw)
            arena += (n + 255) // 256 * 256

        def arg(kind, alloc=0, value=0):
            return struct.pack("<IIQ", kind, alloc, value & 0xFFFFFFFFFFFFFFFF)

        def fbits(x: float) -> int:
            return struct.unpack("<I", struct.pack("<f", float(x)))[0]

        def ptr_arg(v):
            v = getattr(v, "value", v)
            if v is None or v == 0:
                return arg(A_NULL)
            i, off = locate(int(v))
            return arg(A_PTR, i, off)

        recs = []
        for kind, name, payload in self.calls:
            args = []
            if kind == OP_CALL:
                kinds = _kinds()[name]
                assert len(kinds) == len(payload), (name, len(kinds), len(payload))
                for k, v in zip(kinds, payload):
                    if k == "p":
                        args.append(ptr_arg(v))
                    elif k == "s":
                        args.append(arg(A_STREAM))
                    elif k == "f":
                        args.append(arg(A_FLOAT, 0, fbits(getattr(v, "value", v))))
                    else:
                        args.append(arg(A_LONG if k == "l" else A_INT, 0, int(getattr(v, "value", v))))
            elif kind == OP_CONV:
                p = payload
                for f in ("src0", "src1", "weight", "bias", "gate", "addend", "out", "gn_sums", "weight_up4", "gn_ws", "ln_v", "ln_rstd"):
                    args.append(ptr_arg(getattr(p, f)))
                for f in ("c0", "c1", "ld0", "B", "Hin", "Win", "Cout", "KH", "KW", "stride", "pad", "upsample", "silu_from",
                          "gate_stride", "gn_groups", "per_batch_weight", "dtype", "relu_out", "ab_dtype_p1"):
                    args.append(arg(A_INT, 0, int(getattr(p, f))))
                args.append(arg(A_FLOAT, 0, fbits(p.ln_eps)))
            else:
                args = [ptr_arg(payload[0]), arg(A_LONG, 0, payload[1])]
            assert len(args) <= MAX_ARGS
            body = b"".join(args) + arg(A_INT) * (MAX_ARGS - len(args))
            recs.append(struct.pack("<II40s", kind, len(args), name.encode()[:39]) + body)
        n_unet = self.unet_ops if self.unet_ops is not None else len(recs)
        with open(path, "wb") as f:
            f.write(struct.pack("<8sIIIIQQ", MAGIC, len(allocs), len(recs), n_unet, 0, arena, data_off))
            f.write(b"".join(allocs))
            f.write(b"".join(recs))
            for raw in blobs:
                f.write(raw)
        return path


@torch.no_grad()
def export_step_program(diffusion, ldct: torch.Tensor, path: str, noise=None) -> str:
    """Runs `diffusion.sample([ldct])` with its first timestep recorded, and writes the plan file.  The recorded buffers keep the
    conditioning of THESE slices (DA-CLIP embeddings -> prompt / local vectors); a host that samples other slices overwrites
    "x_input", "x_t" and the conditioning buffers before stepping."""
    rec = Recorder()
    diffusion.sample([ldct], batch_size=ldct.shape[0], last=True, noise=noise, _record=rec)
    return rec.save(path)


class StepProgram:
    """The C runtime (fd_program_*) driven through ctypes on a torch-owned arena: what examples/c_host.c does with cudaMalloc."""

    def __init__(self, path: str, device="cuda"):
        lib = _lib.load()
        n = lib.fd_program_arena_bytes(path.encode())
        if n < 0:
            raise _lib.FdError(f"{path}: not a founddiff_b200 plan file")
        self.arena = torch.empty(int(n), dtype=torch.uint8, device=device)
        self._h = ctypes.c_void_p()
        _lib.check(lib.fd_program_load(path.encode(), ctypes.c_void_p(self.arena.data_ptr()), int(n), ctypes.byref(self._h)), "fd_program_load")
        self._lib = lib

    def buffer(self, name: str, dtype=torch.float32) -> torch.Tensor:
        nb = ctypes.c_long()
        p = self._lib.fd_program_buffer(self._h, name.encode(), ctypes.byref(nb))
        if not p:
            raise KeyError(name)
        off = int(p) - self.arena.data_ptr()
        return self.arena[off:off + nb.value].view(dtype)

    def num_launches(self) -> int:
        return int(self._lib.fd_program_num_launches(self._h))

    def unet_step(self):
        _lib.check(self._lib.fd_unet_step(self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "fd_unet_step")

    def sample_step(self):
        _lib.check(self._lib.fd_sample_step(self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)), "fd_sample_step")

    def __del__(self):
        try:
            if self._h:
                self._lib.fd_program_destroy(self._h)
        except Exception:
            pass
