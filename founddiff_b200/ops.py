"""Thin tensor-level wrappers over the C ABI (include/founddiff_b200.h).  PyTorch here is only the owner of device
memory and streams: every function passes raw pointers + the current CUDA stream to a hand-written kernel.
No function in this file has a PyTorch/CPU fallback; a non-CUDA tensor raises."""
from __future__ import annotations

import ctypes
from ctypes import byref, c_void_p
from typing import Optional

import torch

from . import _lib
from ._lib import FD_BF16, FD_F16, FD_F32, ConvParams, check

_DT = {torch.float32: FD_F32, torch.bfloat16: FD_BF16, torch.float16: FD_F16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"unsupported activation dtype {dt}") from None


_REC = None                      # program.Recorder while a step is being recorded
CONST_STORAGES = set()           # storage pointers of packed weights (engine.upload): stored with their bytes in a plan file


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if _REC is not None:
        _REC.note(t)
    if not t.is_cuda:
        raise _lib.FdError("founddiff_b200 ops need CUDA tensors (there is no CPU fallback)")
    if not t.is_contiguous():
        raise ValueError("tensor must be contiguous")
    return c_void_p(t.data_ptr())


def _f32(t: Optional[torch.Tensor]):
    if t is not None and t.dtype != torch.float32:
        raise TypeError("expected a float32 tensor")
    return _p(t)


def _stream():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def version() -> str:
    return _lib.load().fd_version().decode()


# ---------------------------------------------------------------------------------------------------------
# NVTX ranges: FD_NVTX=1 brackets sample() / DA-CLIP / every timestep / every Unet block, so that a
# timeline (ncu --nvtx, nsys where available) is labelled with the reference's own structure.  Off by default: a push / pop
# pair per block is ~1 us of host time per launch group.
import os as _os
NVTX = _os.environ.get("FD_NVTX", "0") == "1"


class nvtx_range:
    __slots__ = ("name",)

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if NVTX:
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if NVTX:
            torch.cuda.nvtx.range_pop()
        return False


# ---------------------------------------------------------------------------------------------------------
# Launch accounting / per-kernel timing (bench.py).  LAUNCHES counts kernels launched through this module;
# when PROFILE is a list every op is bracketed by CUDA events on the launching stream and appended as
# (name, detail, start_event, end_event).
LAUNCHES = 0
PROFILE = None


def _launched(name: str, detail: str = "", n: int = 1):
    """Context manager used by every wrapper: counts `n` kernel launches and optionally times them."""
    return _Launch(name, detail, n)


class _Launch:
    __slots__ = ("name", "detail", "n", "ev")

    def __init__(self, name, detail, n):
        self.name, self.detail, self.n, self.ev = name, detail, n, None

    def __enter__(self):
        global LAUNCHES
        LAUNCHES += self.n
        if PROFILE is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None:
            self.ev[1].record()
            PROFILE.append((self.name, self.detail, self.ev[0], self.ev[1]))
        return False


# ---------------------------------------------------------------------------------------------------------
def selective_scan_fwd(u, delta, A, B, C, D=None, delta_bias=None, delta_softplus=True, out=None):
    """u, delta: (b, KD, L) [fp32/bf16/fp16]; A: (KD, N) fp32; B, C: (b, K, N, L) fp32; D, delta_bias: (KD,) fp32."""
    lib = _lib.load()
    b, kd, L = u.shape
    n, g = A.shape[1], B.shape[1]
    assert delta.shape == u.shape and delta.dtype == u.dtype and A.shape[0] == kd
    assert B.shape == (b, g, n, L) and C.shape == B.shape
    y = torch.empty_like(u) if out is None else out
    with _launched("selective_scan", f"{b}x{kd}x{L} N{n}"):
        check(lib.fd_selective_scan_fwd(_p(u), _p(delta), _f32(A), _f32(B), _f32(C), _f32(D), _f32(delta_bias), _p(y),
                                        b, kd, L, n, g, int(bool(delta_softplus)), dtype_code(u.dtype), _stream()),
              "fd_selective_scan_fwd")
    return y


def selective_scan_fwd_merge(u, delta, A, B, C, D, delta_bias, delta_softplus, y_nhwc, H, W):
    """Scan + EfficientMerge: u, delta (b, 4*Dg, L); y_nhwc (b, H*W, Dg) channels-last."""
    b, kd, L = u.shape
    n = A.shape[1]
    assert L == (H // 2) * (W // 2) and B.shape == (b, 4, n, L)
    with _launched("selective_scan_merge", f"{b}x{kd}x{L} N{n}"):
        check(_lib.load().fd_selective_scan_fwd_merge(_p(u), _p(delta), _f32(A), _f32(B), _f32(C), _f32(D), _f32(delta_bias),
                                                      _p(y_nhwc), b, kd, H, W, n, int(bool(delta_softplus)), dtype_code(u.dtype),
                                                      _stream()), "fd_selective_scan_fwd_merge")
    return y_nhwc


def row_rstd(x, rstd, rows, C, eps):
    """rstd[row] = 1 / sqrt(var(x[row, :C]) + eps): the statistics of a LayerNorm folded into the GEMM that reads x (Conv(ln_rstd=...))."""
    with _launched("row_rstd", f"{rows}x{C}"):
        check(_lib.load().fd_row_rstd(_p(x), _f32(rstd), rows, C, float(eps), dtype_code(x.dtype), _stream()), "fd_row_rstd")


def ln_gate(y, xz, ld, z_off, gamma, beta, local, out, B, P, C, eps=1e-5):
    with _launched("ln_gate", f"{B}x{P}x{C}"):
        check(_lib.load().fd_ln_gate(_p(y), _p(xz), ld, z_off, _f32(gamma), _f32(beta), _f32(local), _p(out), B, P, C,
                                     float(eps), dtype_code(y.dtype), _stream()), "fd_ln_gate")


def ln_gate_out_proj_supported(P, D, Cout, ld, z_off, io_dtype, out_dtype) -> bool:
    if io_dtype == torch.float32 or out_dtype == torch.float32:
        return False
    return bool(_lib.load().fd_ln_gate_out_proj_supported(P, D, Cout, ld, z_off, dtype_code(io_dtype), dtype_code(out_dtype)))


def ln_gate_out_proj(y, xz, ld, z_off, gamma, beta, local, w, gate, gate_stride, addend, out, B, P, D, Cout, eps=1e-5):
    """out = addend + gate * (w . ((LN(y) * gamma + beta) * z + local)): fd_ln_gate + out_proj + gated residual in one pass."""
    with _launched("ln_gate_out_proj", f"{B}x{P}x{D}->{Cout}"):
        check(_lib.load().fd_ln_gate_out_proj(_p(y), _p(xz), ld, z_off, _f32(gamma), _f32(beta), _f32(local), _p(w), _f32(gate), gate_stride,
                                              _p(addend), _p(out), B, P, D, Cout, float(eps), dtype_code(y.dtype), dtype_code(out.dtype),
                                              _stream()), "fd_ln_gate_out_proj")


def pack_upsample_phases(weight: torch.Tensor, cout: int, cin: int) -> torch.Tensor:
    """nearest-x2 upsample followed by a 3x3 conv == 4 output phases, each a 2x2 conv over the low-resolution input
    whose taps are sums of the 3x3 taps that land on the same low-res pixel.  weight: (Cout, 3, 3, Cin) ->
    (4, Cout, 2, 2, Cin), phase = 2*row_parity + col_parity."""
    w = weight.reshape(cout, 3, 3, cin).float()
    # row/col combination per parity: parity 0 -> taps {0}, {1,2};  parity 1 -> taps {0,1}, {2}
    comb = {0: ([0], [1, 2]), 1: ([0, 1], [2])}
    out = torch.zeros(4, cout, 2, 2, cin, device=w.device, dtype=torch.float32)
    for a in (0, 1):
        for b in (0, 1):
            for dh in (0, 1):
                for dw in (0, 1):
                    acc = 0
                    for kh in comb[a][dh]:
                        for kw in comb[b][dw]:
                            acc = acc + w[:, kh, kw, :]
                    out[2 * a + b, :, dh, dw, :] = acc
    return out.to(weight.dtype).contiguous()


class Conv:
    """One convolution / 1x1 GEMM call site (fd_conv_params).  `run()` launches it; buffers are bound at
    construction so that the call is CUDA-graph friendly (and so that the tcgen05 path can bake TMA descriptors)."""

    def __init__(self, src0, weight, out, *, B, Hin, Win, KH=1, KW=1, stride=1, pad=0, upsample=False, src1=None,
                 bias=None, gate=None, gate_stride=0, addend=None, silu_from=None, gn_sums=None, gn_groups=0,
                 per_batch_weight=False, prefer_tc=True, c0=None, ld0=0, relu_out=False, gn_ws=None, weight_up4=None,
                 ln_v=None, ln_eps=1e-5, ln_rstd=None):
        """`c0` / `ld0`: read only the first c0 channels of rows of pitch ld0 starting at src0's data pointer (src0 may
        be a strided channel-slice view).
        `gn_sums` (B, G, 2): GroupNorm statistics of the output.  With `gn_ws` (zeroed before every run: conv_gn_ws_floats(B) for
        the tcgen05 kernel, gn_stats_ws_floats(B, P, G) otherwise — pass max of the two) they are reproducible: the tcgen05 kernel
        reduces fixed per-sample slots in order; any other shape runs the convolution without statistics and a separate
        reproducible fd_gn_stats pass over its output.  Without gn_ws: float atomics (gn_sums zeroed by the caller)."""
        lib = _lib.load()
        c0 = src0.shape[-1] if c0 is None else c0
        c1 = src1.shape[-1] if src1 is not None else 0
        cout = out.shape[-1]
        p = ConvParams()
        p.src0 = c_void_p(src0.data_ptr()) if ld0 else _p(src0)
        p.src1, p.weight, p.out = _p(src1), _p(weight), _p(out)
        p.bias, p.gate, p.addend, p.gn_sums = _f32(bias), _f32(gate), _p(addend), _f32(gn_sums)
        p.gn_ws = _f32(gn_ws)
        # LayerNorm of the input rows folded into this GEMM (tcgen05 path only): weight = ln_fold()'s per-sample W', see the header
        p.ln_v, p.ln_eps = _f32(ln_v), float(ln_eps)
        p.ln_rstd = _f32(ln_rstd)           # per-pixel rstd from row_rstd(): the GEMM then needs no statistics warps
        self._gn = None
        p.c0, p.c1, p.B, p.Hin, p.Win, p.Cout = c0, c1, B, Hin, Win, cout
        p.ld0 = ld0
        p.KH, p.KW, p.stride, p.pad, p.upsample = KH, KW, stride, pad, int(upsample)
        self._w4 = weight_up4                    # (4, Cout, 2, 2, Cin) phase-summed kernels; packed here unless the caller did (on the host)
        if upsample and prefer_tc and out.dtype != torch.float32:
            if self._w4 is None:
                self._w4 = pack_upsample_phases(weight, cout, c0 + c1)
            p.weight_up4 = _p(self._w4)
        p.silu_from = cout if silu_from is None else silu_from
        p.gate_stride, p.gn_groups, p.per_batch_weight = gate_stride, gn_groups, int(per_batch_weight)
        p.dtype = dtype_code(out.dtype)
        p.relu_out = int(bool(relu_out))
        # operands (src0 / src1 / weight) share one storage type; out / addend may use another 16-bit type
        p.ab_dtype_p1 = 0 if src0.dtype == out.dtype else dtype_code(src0.dtype) + 1
        assert src0.dtype == weight.dtype and (src1 is None or src1.dtype == src0.dtype), (src0.dtype, weight.dtype)
        assert addend is None or addend.dtype == out.dtype
        assert weight.numel() == (B if per_batch_weight else 1) * cout * KH * KW * (c0 + c1), (weight.shape, cout, KH, KW, c0, c1)
        self.params = p
        self._keep = (src0, src1, weight, out, bias, gate, addend, gn_sums, gn_ws, ln_v, ln_rstd)
        self._lib = lib
        self._plan = c_void_p()
        self.uses_tc = False
        if prefer_tc and lib.fd_conv2d_tc_supported(byref(p)):
            check(lib.fd_conv2d_tc_plan_create(byref(p), byref(self._plan)), "fd_conv2d_tc_plan_create")
            self.uses_tc = True
        elif ln_v is not None:
            raise _lib.FdError("Conv: the LayerNorm fold needs the tcgen05 path (1x1, stride 1, tiling geometry): " + self.describe())
        elif gn_sums is not None and gn_ws is not None:
            # CUDA-core path: statistics by a separate reproducible pass over the stored output (the kernel's own epilogue sums
            # use float atomics)
            Pout = out.shape[-2] if out.dim() >= 2 else 0
            assert gn_ws.numel() >= gn_stats_ws_floats(B, Pout, gn_groups), "gn_ws too small for the CUDA-core path"
            p.gn_sums, p.gn_ws = None, None
            self._gn = (out, gn_sums, gn_ws, B, Pout, cout, gn_groups)

    def describe(self) -> str:
        p = self.params
        return (f"{p.B}x{p.Hin}x{p.Win} {p.c0}+{p.c1}->{p.Cout} k{p.KH} s{p.stride}" + (" up2" if p.upsample else "")
                + (" wb" if p.per_batch_weight else ""))

    def flops(self) -> float:
        """Dense FLOPs of the convolution as the reference computes it (2*MAC; upsample counted at 3x3 on the
        upsampled grid, as FlopCounterMode does for the reference — SURVEY.md Appendix A)."""
        p = self.params
        up = 2 if p.upsample else 1
        ho = (p.Hin * up + 2 * p.pad - p.KH) // p.stride + 1
        wo = (p.Win * up + 2 * p.pad - p.KW) // p.stride + 1
        return 2.0 * p.B * ho * wo * p.Cout * p.KH * p.KW * (p.c0 + p.c1)

    def run(self):
        if _REC is not None:
            _REC.conv(self)
        with _launched("conv_tc" if self.uses_tc else "conv_simt", self.describe(),
                       2 if (self.uses_tc and self.params.gn_sums and self.params.gn_ws) else 1):
            if self.uses_tc:
                check(self._lib.fd_conv2d_tc_run(self._plan, _stream()), "fd_conv2d_tc_run")
            else:
                check(self._lib.fd_conv2d_simt(byref(self.params), _stream()), "fd_conv2d_simt")
        if self._gn is not None:
            o, sums, ws, B, P, C, G = self._gn
            gn_stats(o, sums, B, P, C, G, ws=ws)

    def __del__(self):
        try:
            if self._plan:
                self._lib.fd_conv2d_tc_plan_destroy(self._plan)
        except Exception:
            pass


def slice_metrics(pred, target, out, B, H, W, max_val=1.0):
    """out (B, 2) fp32: per-slice sum of squared errors and sum of the SSIM map (see metrics.py for PSNR / SSIM / RMSE)."""
    with _launched("slice_metrics", f"{B}x{H}x{W}", 1):
        check(_lib.load().fd_slice_metrics(_f32(pred), _f32(target), _f32(out), B, H, W, float(max_val), _stream()), "fd_slice_metrics")


def avgpool2x2_nhwc(x, out, B, H, W, C):
    with _launched("avgpool2x2_nhwc", f"{B}x{H}x{W}x{C}", 1):
        check(_lib.load().fd_avgpool2x2_nhwc(_p(x), _p(out), B, H, W, C, dtype_code(x.dtype), _stream()), "fd_avgpool2x2_nhwc")


def init_conv7x7(x_t, x_input, weight, bias, out, B, H, W):
    with _launched("init_conv7x7", f"{B}x{H}x{W}", 1):
        check(_lib.load().fd_init_conv7x7(_f32(x_t), _f32(x_input), _f32(weight), _f32(bias), _p(out), B, H, W,
                                          out.shape[-1], dtype_code(out.dtype), _stream()), "fd_init_conv7x7")


def pack_init_conv_weights(weight: torch.Tensor) -> torch.Tensor:
    """(64, 2, 7, 7) fp32 -> (64, 256) fp16 for fd_init_conv7x7_tc: K index = ci*56 + ky*8 + kx (tap rows padded to 8 with a
    zero weight), 112 of 128 used, stored twice (the hi and the lo half of the fp16-split images)."""
    co = weight.shape[0]
    w = torch.zeros(co, 2, 7, 8, device=weight.device, dtype=torch.float16)
    w[..., :7] = weight.detach().to(torch.float16)
    out = torch.zeros(co, 256, device=weight.device, dtype=torch.float16)
    out[:, :112] = w.reshape(co, 112)
    out[:, 128:240] = w.reshape(co, 112)
    return out.contiguous()


def init_conv7x7_tc(x_t, x_input, w16, bias, out, B, H, W):
    with _launched("init_conv7x7_tc", f"{B}x{H}x{W}", 1):
        check(_lib.load().fd_init_conv7x7_tc(_f32(x_t), _f32(x_input), _p(w16), _f32(bias), _p(out), B, H, W, out.shape[-1],
                                             dtype_code(out.dtype), _stream()), "fd_init_conv7x7_tc")


def ln_modulate(x, out, gamma, beta, shift, scale, mod_stride, B, P, C, eps):
    with _launched("ln_modulate", f"{B}x{P}x{C}", 1):
        check(_lib.load().fd_ln_modulate_io(_p(x), _p(out), _f32(gamma), _f32(beta), _f32(shift), _f32(scale), mod_stride,
                                            B, P, C, float(eps), dtype_code(x.dtype), dtype_code(out.dtype), _stream()),
              "fd_ln_modulate_io")


def ln_fold(W, gamma, beta, shift, scale, mod_stride, Wf, v, B, Cout, C):
    """Per-sample folded weights of LayerNorm + adaLN modulate + 1x1 GEMM: Wf (B, Cout, C) 16-bit (zero row sums), v (B, Cout) fp32."""
    with _launched("ln_fold", f"{B}x{Cout}x{C}", 1):
        check(_lib.load().fd_ln_fold(_f32(W), _f32(gamma), _f32(beta), _f32(shift), _f32(scale), mod_stride, _p(Wf), _f32(v),
                                     B, Cout, C, dtype_code(Wf.dtype), _stream()), "fd_ln_fold")


def dwconv3x3_silu_scan(xz, ld, w, bias, xs, B, H, W, D):
    with _launched("dwconv_scan", f"{B}x{H}x{W}x{D}", 1):
        check(_lib.load().fd_dwconv3x3_silu_scan(_p(xz), ld, _f32(w), _f32(bias), _p(xs), B, H, W, D,
                                                 dtype_code(xz.dtype), _stream()), "fd_dwconv3x3_silu_scan")


def xdt_proj(xs, x_proj_w, dt_w, dts, Bs, Cs, B, D, L, R, N):
    with _launched("xdt_proj", f"{B}x{D}x{L} R{R} N{N}", 1):
        check(_lib.load().fd_xdt_proj(_p(xs), _f32(x_proj_w), _f32(dt_w), _p(dts), _f32(Bs), _f32(Cs), B, D, L, R, N,
                                      dtype_code(xs.dtype), _stream()), "fd_xdt_proj")


def pack_xdt_weights(x_proj_w: torch.Tensor, dt_w: torch.Tensor, dtype: torch.dtype):
    """(4, R+2N, D) / (4, D, R) fp32 -> zero-padded 16-bit copies for fd_xdt_proj_tc: (4, CCp, D), (4, D, Rp), Rp."""
    K, CC, D = x_proj_w.shape
    R = dt_w.shape[-1]
    CCp = (CC + 15) // 16 * 16
    Rp = 16 if R <= 16 else 32
    xw = torch.zeros(K, CCp, D, device=x_proj_w.device, dtype=dtype)
    xw[:, :CC] = x_proj_w.to(dtype)
    dw = torch.zeros(K, D, Rp, device=dt_w.device, dtype=dtype)
    dw[:, :, :R] = dt_w.to(dtype)
    return xw.contiguous(), dw.contiguous(), Rp


def xdt_proj_tc(xs, xw16, dw16, Rp, dts, Bs, Cs, B, D, L, R, N, time_major=False, dt_bias=None, delta_softplus=False):
    """time_major: Bs / Cs written as (B, 4, L, N) (for selective_scan_fwd_merge_cl) instead of (B, 4, N, L).
    dt_bias (4*D,) fp32: dts = [softplus](dt_proj(...) + dt_bias), so that the scan takes delta as is."""
    with _launched("xdt_proj_tc", f"{B}x{D}x{L} R{R} N{N}", 1):
        check(_lib.load().fd_xdt_proj_tc(_p(xs), _p(xw16), _p(dw16), _p(dts), _f32(Bs), _f32(Cs), B, D, L, R, N, Rp,
                                         int(bool(time_major)), _f32(dt_bias), int(bool(delta_softplus)), dtype_code(xs.dtype),
                                         _stream()), "fd_xdt_proj_tc")


def x_proj_tc(xs, xw16, x_dbl, B, D, L, R, N):
    """x_dbl (B, 4, R+2N, L) fp32 = x_proj(xs) on the tensor cores; consumed by selective_scan_fwd_merge_xdbl."""
    assert x_dbl.dtype == torch.float32 and x_dbl.shape == (B, 4, R + 2 * N, L)
    with _launched("x_proj_tc", f"{B}x{D}x{L} R{R} N{N}", 1):
        check(_lib.load().fd_x_proj_tc(_p(xs), _p(xw16), _f32(x_dbl), B, D, L, R, N, dtype_code(xs.dtype), _stream()), "fd_x_proj_tc")


def dwconv3x3_silu_tm(xz, ld, w_tap_major, bias, xs_tm, B, H, W, D):
    """Depthwise 3x3 + bias + SiLU on the x half of xz -> time-major scan input (B, 4, L, D)."""
    assert tuple(w_tap_major.shape) == (9, D)
    with _launched("dwconv_tm", f"{B}x{H}x{W}x{D}", 1):
        check(_lib.load().fd_dwconv3x3_silu_tm(_p(xz), ld, _f32(w_tap_major), _f32(bias), _p(xs_tm), B, H, W, D,
                                               dtype_code(xz.dtype), _stream()), "fd_dwconv3x3_silu_tm")


def x_proj_tm(xs_tm, xw16, xdbl, dw16, dts_tm, dt_bias, B, D, L, R, N, Rp, fuse_dt):
    """x_proj (+ dt_proj + bias + softplus when not fused into the scan) on time-major rows; xdbl (B,4,L,R+2N | 2N) fp32."""
    assert xdbl.dtype == torch.float32 and xdbl.shape == (B, 4, L, (R + 2 * N) if fuse_dt else 2 * N)
    with _launched("x_proj_tm", f"{B}x{D}x{L} R{R} N{N}" + (" dt-fused" if fuse_dt else ""), 1):
        check(_lib.load().fd_x_proj_tm(_p(xs_tm), _p(xw16), _f32(xdbl), _p(dw16), _p(dts_tm), _f32(dt_bias), B, D, L, R, N, Rp,
                                       int(bool(fuse_dt)), dtype_code(xs_tm.dtype), _stream()), "fd_x_proj_tm")


def scan_tm_segments(B, D, H, W) -> int:
    return int(_lib.load().fd_scan_tm_segments(B, D, H, W))


def scan_tm_plan(B, D, H, W, N, R_fused) -> int:
    """> 0: segments of the segmented scan; -8 / -4: time-sliced cooperative scan with that many warps per block."""
    return int(_lib.load().fd_scan_tm_plan(B, D, H, W, N, R_fused))


def selective_scan_tm(u_tm, dts_tm, xdbl, A, dt_w, dt_bias, D_skip, carry, y_nhwc, B, D, H, W, N, R_fused, segments=0):
    """Time-major scan + EfficientMerge.  segments: 0 = automatic, > 0 = segmented channel-per-lane scan (carry pass + forward
    pass), -8 / -4 = time-sliced cooperative scan (one launch)."""
    L = (H // 2) * (W // 2)
    S = segments or scan_tm_plan(B, D, H, W, N, R_fused)
    with _launched("scan_tm", f"{B}x{4 * D}x{L} N{N}" + (f" R{R_fused}" if R_fused else "") + (f" S{S}" if S > 0 else f" TW{-S}"),
                   2 if S > 1 else 1):
        check(_lib.load().fd_selective_scan_tm(_p(u_tm), _p(dts_tm), _f32(xdbl), _f32(A), _f32(dt_w), _f32(dt_bias), _f32(D_skip),
                                               _f32(carry), carry.numel() if carry is not None else 0, _p(y_nhwc), B, D, H, W, N,
                                               R_fused, segments, dtype_code(u_tm.dtype), _stream()), "fd_selective_scan_tm")
    return y_nhwc


def scan_tm_chain_plan(B, D, H, W, N, R_fused):
    """(segments, workspace floats) of the chained time-sliced scan for this geometry; (0, 0) where it does not apply."""
    n = ctypes.c_int(0)
    nseg = int(_lib.load().fd_scan_tm_chain_plan(B, D, H, W, N, R_fused, ctypes.byref(n)))
    return nseg, int(n.value)


def selective_scan_tm_chained(u_tm, xdbl, A, dt_w, dt_bias, D_skip, chain_ws, y_nhwc, B, D, H, W, N, R_fused):
    """Time-sliced scan + EfficientMerge with chained segments (fd_selective_scan_tm_chained).  chain_ws: fp32 workspace of
    scan_tm_chain_plan(...)[1] floats, zero-filled once by the caller; same bits as selective_scan_tm."""
    L = (H // 2) * (W // 2)
    tw, (nseg, _) = -scan_tm_plan(B, D, H, W, N, R_fused), scan_tm_chain_plan(B, D, H, W, N, R_fused)
    with _launched("scan_tm", f"{B}x{4 * D}x{L} N{N} R{R_fused} TW{tw} x{nseg}", 1):
        check(_lib.load().fd_selective_scan_tm_chained(_p(u_tm), _f32(xdbl), _f32(A), _f32(dt_w), _f32(dt_bias), _f32(D_skip),
                                                       _f32(chain_ws), chain_ws.numel(), _p(y_nhwc), B, D, H, W, N, R_fused,
                                                       dtype_code(u_tm.dtype), _stream()), "fd_selective_scan_tm_chained")
    return y_nhwc


class LinearAttention:
    """lucidrains LinearAttention between to_qkv and the end of to_out (src/denoising_diffusion_pytorch.py:238-255) as a
    call site with pre-allocated workspace and a persistent per-sample GEMM plan (graph-capturable, no allocation per call):
    qkv (B, H*W, 3*heads*32) -> out (B, H*W, dim) = LayerNorm_c(Conv1x1(linear-attention(q, k, v)) + bias) * g."""

    def __init__(self, qkv, wout, bias, g, out, B, H, W, heads, dim, scale=32 ** -0.5, prefer_tc=True):
        N, HC, dev, dt = H * W, heads * 32, qkv.device, qkv.dtype
        self.args = (B, N, heads, dim, float(scale), H, W)
        self.qkv, self.wout, self.bias, self.g, self.out = qkv, wout, bias, g, out
        self.kmax = torch.empty(B, HC, device=dev)
        self.ksum = torch.empty(B, HC, device=dev)
        self.ctx = torch.empty(B, heads, 32, 32, device=dev)
        self.weff = torch.empty(B, dim, HC, device=dev, dtype=dt)
        self.qhat = torch.empty(B, N, HC, device=dev, dtype=dt)
        self.y = torch.empty(B, N, dim, device=dev, dtype=dt)
        self.zb = torch.zeros_like(g)
        self.zeros = torch.zeros(B, dim, device=dev)
        self.conv = Conv(self.qhat, self.weff, self.y, B=B, Hin=H, Win=W, bias=bias, per_batch_weight=True, prefer_tc=prefer_tc)

    def run(self):
        lib = _lib.load()
        B, N, heads, dim, scale, H, W = self.args
        dt = self.qkv.dtype
        self.kmax.fill_(float("-inf"))
        self.ksum.zero_()
        self.ctx.zero_()
        with _launched("linattn_context", f"{B}x{N}x{heads}", 2):
            check(lib.fd_linattn_context(_p(self.qkv), _f32(self.kmax), _f32(self.ksum), _f32(self.ctx), B, N, heads, dtype_code(dt),
                                         _stream()), "fd_linattn_context")
        with _launched("linattn_weff", f"{B}x{dim}"):
            check(lib.fd_linattn_weff(_f32(self.ctx), _f32(self.ksum), _f32(self.wout), _p(self.weff), B, N, heads, dim, scale,
                                      dtype_code(dt), _stream()), "fd_linattn_weff")
        with _launched("softmax_d32", f"{B}x{N}x{heads}"):
            check(lib.fd_softmax_d32(_p(self.qkv), _p(self.qhat), B, N, heads, dtype_code(dt), _stream()), "fd_softmax_d32")
        self.conv.run()
        ln_modulate(self.y, self.out, self.g, self.zb, self.zeros, self.zeros, dim, B, N, dim, 1e-5)


def selective_scan_fwd_merge_cl(u, delta, A, Bt, Ct, D, delta_bias, delta_softplus, y_nhwc, H, W):
    """Channel-per-lane scan + EfficientMerge (deep levels): Bt, Ct time-major (b, 4, L, N) fp32."""
    b, kd, L = u.shape
    n = A.shape[1]
    assert L == (H // 2) * (W // 2) and Bt.shape == (b, 4, L, n) and Ct.shape == (b, 4, L, n)
    with _launched("selective_scan_merge", f"{b}x{kd}x{L} N{n} cl"):
        check(_lib.load().fd_selective_scan_fwd_merge_cl(_p(u), _p(delta), _f32(A), _f32(Bt), _f32(Ct), _f32(D), _f32(delta_bias),
                                                         _p(y_nhwc), b, kd, H, W, n, int(bool(delta_softplus)), dtype_code(u.dtype),
                                                         _stream()), "fd_selective_scan_fwd_merge_cl")
    return y_nhwc


def selective_scan_fwd_merge_xdbl(u, x_dbl, dt_w, A, D, delta_bias, delta_softplus, y_nhwc, H, W):
    """Scan + EfficientMerge with dt_proj fused: u (b, 4*Dg, L) 16-bit, x_dbl (b, 4, R+2N, L) fp32, dt_w (4*Dg, R) fp32."""
    b, kd, L = u.shape
    n, r = A.shape[1], dt_w.shape[1]
    assert L == (H // 2) * (W // 2) and x_dbl.shape == (b, 4, r + 2 * n, L) and dt_w.shape[0] == kd
    with _launched("selective_scan_merge", f"{b}x{kd}x{L} N{n} dt-fused"):
        check(_lib.load().fd_selective_scan_fwd_merge_xdbl(_p(u), _f32(x_dbl), _f32(dt_w), _f32(A), _f32(D), _f32(delta_bias),
                                                           _p(y_nhwc), b, kd, H, W, n, r, int(bool(delta_softplus)),
                                                           dtype_code(u.dtype), _stream()), "fd_selective_scan_fwd_merge_xdbl")
    return y_nhwc


def merge_ln_gate(ys, xz, ld, z_off, gamma, beta, local, stats_ws, out, B, H, W, D, eps=1e-5):
    with _launched("merge_ln_gate", f"{B}x{H}x{W}x{D}", 2):
        check(_lib.load().fd_merge_ln_gate(_p(ys), _p(xz), ld, z_off, _f32(gamma), _f32(beta), _f32(local), _f32(stats_ws),
                                           _p(out), B, H, W, D, float(eps), dtype_code(ys.dtype), _stream()),
              "fd_merge_ln_gate")


def gram_ws_floats(B, H, W, C, dtype) -> int:
    """Size of the zeroed fp32 workspace of dwconv3x3_qkv_gram (fp32 storage) / gram_qk (16-bit storage)."""
    return int(_lib.load().fd_gram_ws_floats(B, H, W, C, dtype_code(dtype)))


def dwconv3x3_qkv_gram(qkv, w, v, gram, qk_sq, B, H, W, C, ws=None):
    if ws is None:
        ws = torch.zeros(gram_ws_floats(B, H, W, C, qkv.dtype), device=qkv.device)
    with _launched("dwconv_qkv_gram", f"{B}x{H}x{W}x{C}", 1):
        check(_lib.load().fd_dwconv3x3_qkv_gram(_p(qkv), _f32(w), _p(v), _f32(gram), _f32(qk_sq), _f32(ws), B, H, W, C,
                                                dtype_code(qkv.dtype), _stream()), "fd_dwconv3x3_qkv_gram")


def dwconv3x3_nhwc(x, w, bias, out, B, H, W, C, silu=False):
    """w: TAP-MAJOR (9, C) fp32 (use `w.reshape(C, 9).t().contiguous()` on a (C,1,3,3) PyTorch weight)."""
    assert tuple(w.shape) == (9, C), w.shape
    with _launched("dwconv3x3_nhwc", f"{B}x{H}x{W}x{C}", 1):
        check(_lib.load().fd_dwconv3x3_nhwc(_p(x), _f32(w), _f32(bias), _p(out), B, H, W, C, int(silu), dtype_code(x.dtype),
                                            _stream()), "fd_dwconv3x3_nhwc")


def gram_qk(qkv, ld, gram, qk_sq, B, P, C, ws=None):
    """ws: zeroed fp32 scratch of gram_ws_floats(B, 1, P, C, dtype) elements (allocated here when omitted)."""
    if ws is None:
        ws = torch.zeros(gram_ws_floats(B, 1, P, C, qkv.dtype), device=qkv.device)
    with _launched("gram_qk", f"{B}x{P}x{C}", 1):
        check(_lib.load().fd_gram_qk(_p(qkv), ld, _f32(gram), _f32(qk_sq), _f32(ws), B, P, C, dtype_code(qkv.dtype), _stream()),
              "fd_gram_qk")


def attn_weff(gram, qk_sq, temperature, proj_w, weff, B, C):
    with _launched("attn_weff", f"{B}x{C}", 1):
        check(_lib.load().fd_attn_weff(_f32(gram), _f32(qk_sq), _f32(temperature), _f32(proj_w), _p(weff), B, C,
                                       dtype_code(weff.dtype), _stream()), "fd_attn_weff")


def gn_stats_ws_floats(B, P, G) -> int:
    return int(_lib.load().fd_gn_stats_ws_floats(B, P, G))


def conv_gn_ws_floats(B) -> int:
    """Size of the `gn_ws` workspace of a Conv with gn_sums (reproducible GroupNorm statistics)."""
    return int(_lib.load().fd_conv_gn_ws_floats(B))


def gn_stats(y, sums, B, P, C, G, ws=None):
    """sums (B, G, 2) = per-group (sum, sum of squares) of y (B, P, C); reproducible (no float atomics).  ws: ZEROED fp32 scratch of
    gn_stats_ws_floats(B, P, G) elements (allocated here when omitted)."""
    if ws is None:
        ws = torch.zeros(gn_stats_ws_floats(B, P, G), device=y.device, dtype=torch.float32)
    with _launched("gn_stats", f"{B}x{P}x{C}", 1):
        check(_lib.load().fd_gn_stats(_p(y), _f32(sums), _f32(ws), B, P, C, G, dtype_code(y.dtype), _stream()), "fd_gn_stats")


def gn_silu_add(y, sums, gamma, beta, skip, out, B, P, C, G, eps=1e-5):
    with _launched("gn_silu_add", f"{B}x{P}x{C}", 1):
        check(_lib.load().fd_gn_silu_add(_p(y), _f32(sums), _f32(gamma), _f32(beta), _p(skip), _p(out), B, P, C, G,
                                         float(eps), dtype_code(y.dtype), _stream()), "fd_gn_silu_add")


def gn_scale_shift_silu(y, sums, gamma, beta, scale, shift, ss_stride, skip, out, B, P, C, G, eps=1e-5):
    """lucidrains Block with time scale/shift (src/denoising_diffusion_pytorch.py:183-199)."""
    with _launched("gn_scale_shift_silu", f"{B}x{P}x{C}"):
        check(_lib.load().fd_gn_scale_shift_silu(_p(y), _f32(sums), _f32(gamma), _f32(beta), _f32(scale), _f32(shift), ss_stride,
                                                 _p(skip), _p(out), B, P, C, G, float(eps), dtype_code(y.dtype), _stream()),
              "fd_gn_scale_shift_silu")


def flash_attn_d32(qkv, out, B, N, heads, scale, impl=None):
    """lucidrains bottleneck Attention (src/denoising_diffusion_pytorch.py:257-279); qkv (B,N,3*heads*32), out (B,N,heads*32).
    impl: "tc" = tcgen05 / tensor-memory kernel (default), "mma" = the mma.sync kernel of round 1 (FD_FLASH_TC=0 selects it)."""
    if impl is None:
        impl = "mma" if _os.environ.get("FD_FLASH_TC", "1") == "0" else "tc"
    lib = _lib.load()
    fn, name = (lib.fd_flash_attn_d32_tc, "fd_flash_attn_d32_tc") if impl == "tc" else (lib.fd_flash_attn_d32, "fd_flash_attn_d32")
    with _launched("flash_attn_d32" + ("_tc" if impl == "tc" else ""), f"{B}x{N}x{heads}"):
        check(fn(_p(qkv), _p(out), B, N, heads, float(scale), dtype_code(qkv.dtype), _stream()), name)


def linear_attention(qkv, wout, bias, g, out, B, H, W, heads, dim, scale=32 ** -0.5, prefer_tc=True):
    """lucidrains LinearAttention between to_qkv and the end of to_out (src/denoising_diffusion_pytorch.py:238-255):
    qkv (B, H*W, 3*heads*32) -> out (B, H*W, dim) = LayerNorm_c(Conv1x1(linear-attention(q, k, v)) + bias) * g."""
    lib = _lib.load()
    N, HC, dev, dt = H * W, heads * 32, qkv.device, qkv.dtype
    kmax = torch.full((B, HC), float("-inf"), device=dev)
    ksum = torch.zeros(B, HC, device=dev)
    ctx = torch.zeros(B, heads, 32, 32, device=dev)
    with _launched("linattn_context", f"{B}x{N}x{heads}", 2):
        check(lib.fd_linattn_context(_p(qkv), _f32(kmax), _f32(ksum), _f32(ctx), B, N, heads, dtype_code(dt), _stream()), "fd_linattn_context")
    weff = torch.empty(B, dim, HC, device=dev, dtype=dt)
    with _launched("linattn_weff", f"{B}x{dim}"):
        check(lib.fd_linattn_weff(_f32(ctx), _f32(ksum), _f32(wout), _p(weff), B, N, heads, dim, float(scale), dtype_code(dt), _stream()),
              "fd_linattn_weff")
    qhat = torch.empty(B, N, HC, device=dev, dtype=dt)
    with _launched("softmax_d32", f"{B}x{N}x{heads}"):
        check(lib.fd_softmax_d32(_p(qkv), _p(qhat), B, N, heads, dtype_code(dt), _stream()), "fd_softmax_d32")
    y = torch.empty(B, N, dim, device=dev, dtype=dt)
    Conv(qhat, weff, y, B=B, Hin=H, Win=W, bias=bias, per_batch_weight=True, prefer_tc=prefer_tc).run()
    zeros = torch.zeros(B, dim, device=dev)
    ln_modulate(y, out, g, torch.zeros_like(g), zeros, zeros, dim, B, N, dim, 1e-5)
    return out


def zero_(t: torch.Tensor):
    """t.zero_() that a step recording sees (cudaMemsetAsync in the replayed program)."""
    if _REC is not None:
        _REC.memset(t)
    t.zero_()


def linear_small(x, W, bias, out, *, add=None, act_in=0, act_out=0):
    B, K = x.shape
    N = W.shape[0]
    assert W.shape[1] == K and out.shape == (B, N)
    with _launched("linear_small", f"{B}x{K}x{N}"):
        check(_lib.load().fd_linear_small(_f32(x), _f32(W), _f32(bias), _f32(add), _f32(out), B, K, N, act_in, act_out,
                                          _stream()), "fd_linear_small")


def time_sinusoid(time, out):
    with _launched("time_sinusoid", "", 1):
        B, dim = out.shape
        check(_lib.load().fd_time_sinusoid(_f32(time), _f32(out), B, dim, _stream()), "fd_time_sinusoid")


def sampler_init(ldct, noise, noise_scale, x_input, x_t, first):
    with _launched("sampler_init", "", 1):
        check(_lib.load().fd_sampler_init(_f32(ldct), _f32(noise), float(noise_scale), _f32(x_input), _f32(x_t), _f32(first),
                                          ldct.numel(), _stream()), "fd_sampler_init")


OBJECTIVES = {"pred_res": 0, "pred_noise": 1, "pred_res_noise": 2, "pred_x0_noise": 3}     # FD_OBJ_* (include/founddiff_b200.h)


def final_conv_update(feat, w, bias, x_input, x_t, noise, coef, x_next, pred_res=None, pred_noise=None, x_start=None,
                      objective: str = "pred_res", feat1=None, w1=None, bias1=None):
    """final_conv + model_predictions (any objective, src/DADiff.py:1168-1207) + posterior / DDIM update in one kernel.
    `feat1 / w1 / bias1`: the second Unet's final_conv operands for the two-output objectives."""
    npix = x_input.numel()
    C = feat.shape[-1]
    if feat1 is not None and (feat1.dtype != feat.dtype or feat1.shape[-1] != C):
        raise ValueError("both Unets must share the feature dtype and width")
    with _launched("final_conv_update", f"{npix}x{C}"):
        check(_lib.load().fd_final_conv_update_obj(_p(feat), _f32(w), _f32(bias), _p(feat1), _f32(w1), _f32(bias1), _f32(x_input),
                                                   _f32(x_t), _f32(noise), _f32(coef), _f32(x_next), _f32(pred_res),
                                                   _f32(pred_noise), _f32(x_start), npix, C, dtype_code(feat.dtype),
                                                   OBJECTIVES[objective], _stream()), "fd_final_conv_update_obj")


def unnormalize(x, out):
    with _launched("unnormalize", "", 1):
        check(_lib.load().fd_unnormalize(_f32(x), _f32(out), x.numel(), _stream()), "fd_unnormalize")


def ddpm_update(x_t, eps, noise, coef, x_next, x_start=None):
    """lucidrains GaussianDiffusion p_sample / ddim_sample update (src/denoising_diffusion_pytorch.py:588-595, 612-646)."""
    with _launched("ddpm_update", ""):
        check(_lib.load().fd_ddpm_update(_f32(x_t), _f32(eps), _f32(noise), _f32(coef), _f32(x_next), _f32(x_start), x_t.numel(),
                                         _stream()), "fd_ddpm_update")
