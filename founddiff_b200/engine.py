"""The per-timestep denoiser engine: FoundDiff's `Unet.forward` (src/DADiff.py:685-740) as a fixed sequence of
hand-written sm_100a kernels over pre-allocated channels-last buffers.

One `UnetEngine` is built for a fixed (batch, H, W, dtype): weights are packed once (weight standardisation
folded, conv weights in (Cout, KH, KW, Cin) order, -exp(A_logs) precomputed, the nine adaLN projections
concatenated), every buffer is allocated once, every convolution call site is bound to its buffers once
(`ops.Conv`) — so one timestep is pointer-stable and can be captured in a CUDA graph by the sampler.

Data layout in HBM (B slices, level l has H_l x W_l pixels, C_l channels, D = 2C, L = H_l*W_l/4):
    trunk / block-internal activations : (B, H_l, W_l, C)            channels-last, `dtype`
    xz                                 : (B, H_l, W_l, 4C)           [x | silu(z)]
    xs, dts, ys (scan layout)          : (B, 4, D, L)                emamba2.py:207-210 direction order
    Bs, Cs                             : (B, 4, N, L)  fp32
    conditioning                       : t (B,256), mods (B, sum 6C), locals (B, D) fp32
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import os

import torch

from . import ops
from .weights import UnetConfig

GN_GROUPS = 8
# arena names that hold the residual / skip stream or feed GroupNorm (UnetEngine.trunk_dtype); everything else is `dtype`
TRUNK_BUFFERS = ("R", "TA", "TB", "H", "SK", "Y", "A")        # "A": the adaLN-modulated LayerNorm output (normalisation-bounded)
BASE_MID_STATE = 32            # int(base_d_state * 2 ** 3), src/DADiff.py:649


def ws_fold(w: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """WeightStandardizedConv2d (src/DADiff.py:145-152) is a constant at inference: fold it once (fp32 eps)."""
    w = w.float()
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    var = w.var(dim=(1, 2, 3), unbiased=False, keepdim=True)
    return (w - mean) * (var + eps).rsqrt()


def upload(t: torch.Tensor, device, dtype=None) -> torch.Tensor:
    """Host tensor -> device, converted and made contiguous ON THE HOST: weight packing launches no device kernel (one memcpy per
    packed tensor), so an engine build is a few hundred copies instead of ~1000 elementwise launches."""
    t = t.detach()
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    out = t.contiguous().to(device)
    if out.is_cuda:
        ops.CONST_STORAGES.add(out.untyped_storage().data_ptr())      # a plan file stores these with their bytes (program.py)
    return out


def pack_conv(w: torch.Tensor, dtype, device) -> torch.Tensor:
    """(Cout, Cin, KH, KW) -> (Cout, KH, KW, Cin) contiguous in the activation dtype (packed on the host)."""
    return upload(w.detach().float().cpu().permute(0, 2, 3, 1), device, dtype)


class UnetEngine:
    def __init__(self, sd: Dict[str, torch.Tensor], cfg: UnetConfig, B: int, H: int, W: int,
                 dtype: torch.dtype = torch.bfloat16, device="cuda", prefer_tc: bool = True,
                 trunk_dtype: Optional[torch.dtype] = None):
        """`dtype`: storage of the block-internal tensors (xz, scan layout, qkv, LN outputs ...) and of the projections
        that read them.  `trunk_dtype`: storage of the residual / skip stream, of the pre-GroupNorm conv outputs and of the
        convolutions that read the stream.  bf16 sampling uses dtype = bf16 (range for the unnormalised projections and
        scan tensors) with an fp16 trunk (11-bit mantissa where the rounding error accumulates from block to block): the
        oracle ablation in DESIGN.md section 2 puts 80 % of the pure-bf16 error on exactly these tensors."""
        self.cfg, self.B, self.H, self.W, self.dtype, self.device = cfg, B, H, W, dtype, torch.device(device)
        self.trunk_dtype = trunk_dtype or dtype
        if (self.trunk_dtype == torch.float32) != (dtype == torch.float32):
            raise ValueError("fp32 storage does not mix with 16-bit storage")
        self.prefer_tc = prefer_tc
        self._bufs: Dict[str, torch.Tensor] = {}
        # every weight transformation below (fp32 casts, weight standardisation, permutes, concatenations, -exp(A_logs), 16-bit
        # conversion, zero padding) runs on HOST copies; the device only ever receives finished tensors
        sd = {k: v.detach().to("cpu", torch.float32) for k, v in sd.items() if v.is_floating_point()}
        f32 = lambda k: upload(sd[k], self.device)  # noqa: E731
        self.f32 = f32
        self.sd = sd
        d, td = cfg.dim, cfg.time_dim
        dev = self.device

        # ---- conditioning weights (fp32, tiny) --------------------------------------------------------
        self.time_w1, self.time_b1 = f32("time_mlp.1.weight"), f32("time_mlp.1.bias")
        self.time_w2, self.time_b2 = f32("time_mlp.3.weight"), f32("time_mlp.3.bias")
        self.text_w0, self.text_b0 = f32("text_mlp.0.weight"), f32("text_mlp.0.bias")
        self.text_w2, self.text_b2 = f32("text_mlp.2.weight"), f32("text_mlp.2.bias")
        self.prompt = f32("prompt")
        self.pm_w, self.pm_b = f32("prompt_mlp.weight"), f32("prompt_mlp.bias")
        blocks = cfg.mamba_blocks()
        self.mod_total = sum(6 * C for _, C, _ in blocks)
        self.adaln_w = upload(torch.cat([sd[f"{p}.adaLN_modulation.1.weight"] for p, _, _ in blocks], dim=0), self.device)
        self.adaln_b = upload(torch.cat([sd[f"{p}.adaLN_modulation.1.bias"] for p, _, _ in blocks], dim=0), self.device)
        self.local_w = upload(torch.cat([sd[f"{p}.mamba.attn.0.weight"] for p, _, _ in blocks], dim=0), self.device)
        self.local_total = sum(2 * C for _, C, _ in blocks)

        # ---- per-step conditioning buffers ---------------------------------------------------------------
        z = lambda *s: torch.zeros(*s, device=dev, dtype=torch.float32)  # noqa: E731
        self.time = z(B)                    # alphas_cumsum[t] * num_timesteps, written by the sampler
        self.t_sin, self.t_hid, self.t_emb = z(B, d), z(B, td), z(B, td)
        self.prompt_emb = z(B, td)
        self.mods = z(B, self.mod_total)
        self.locals = z(B, self.local_total)
        self.txt_hid, self.txt_out = z(B, td), z(B, td)

        # ---- accumulators zeroed once per forward ----------------------------------------------------------
        self._acc_size = 0
        self._acc_slices = []

        def acc(n):
            off = self._acc_size
            self._acc_size += n
            self._acc_slices.append((off, n))
            return len(self._acc_slices) - 1
        self._acc = acc

        mod_off = loc_off = 0
        self._mod_offsets, self._loc_offsets = {}, {}
        for p, C, _N in blocks:
            self._mod_offsets[p], self._loc_offsets[p] = mod_off, loc_off
            mod_off += 6 * C
            loc_off += 2 * C
        self._local_views = []

        # ---- build the layer list -----------------------------------------------------------------------
        self.steps: List = []          # list of callables executed in order by forward()
        self.step_names: List[str] = []   # NVTX label of each entry (filled after the build from the block prefixes)
        self.paths: Dict[str, str] = {}   # Mamba block -> kernel variant its SS2D core runs (asserted by the B = 16 parity test)
        self._acc_users = []
        self._build(sd)
        self.step_names = [getattr(fn, "_fd_name", f"fd.unet.step{i}") for i, fn in enumerate(self.steps)]
        self.acc_buf = torch.zeros(max(self._acc_size, 1), device=dev, dtype=torch.float32)
        for fn in self._acc_users:
            fn()

    # -------------------------------------------------------------------------------------------------------
    def buf(self, name: str, *shape, dtype=None) -> torch.Tensor:
        """Named arena: one allocation per name, grown to the largest request, returned as a view."""
        if dtype is None:
            dtype = self.trunk_dtype if name.rstrip("0123456789") in TRUNK_BUFFERS else self.dtype
        n = int(math.prod(shape))
        t = self._bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            assert t is None, f"buffer {name} requested again with a larger size; allocate the maximum first"
            t = torch.empty(n, device=self.device, dtype=dtype)
            self._bufs[name] = t
        return t[:n].view(*shape)

    def _acc_view(self, idx, *shape):
        off, n = self._acc_slices[idx]
        return self.acc_buf[off:off + n].view(*shape)

    # -------------------------------------------------------------------------------------------------------
    @staticmethod
    def check_geometry(H: int, W: int, dtype: torch.dtype) -> None:
        """Raises for slice sizes the full Unet cannot (or is not validated to) run."""
        if H % 16 or W % 16:
            raise ValueError("H and W must be multiples of 16 (three 2x downsamplings + the stride-2 scan sub-grids)")
        # 16-bit storage on odd deepest-level scan lengths (48x80 -> 3x5 = 15 steps) was refused in round 1 after a misaligned store;
        # round 2 ran the whole 16-bit path on that geometry under compute-sanitizer memcheck (tools/sanitize.sh,
        # profiles/r2_sanitizer_summary.txt: 0 errors; fp16 1.6e-3, bf16 4.8e-3 against the reference fixture) and lifted it.

    def _build(self, sd):
        cfg, B, H, W, dt, dev = self.cfg, self.B, self.H, self.W, self.dtype, self.device
        self.check_geometry(H, W, dt)
        f32 = self.f32
        d = cfg.dim
        sizes = [(H >> i, W >> i) for i in range(4)]
        up_c = [co for (_ci, co) in reversed(cfg.in_out)]          # Mamba width of ups.0..3 (512, 256, 128, 64)
        # widest trunk at each level (SURVEY.md section 3.2 "Levels")
        cmax = [max(cfg.in_out[0][0], up_c[3]), max(cfg.in_out[1][0], up_c[2]), max(cfg.in_out[2][0], up_c[1]),
                max(cfg.in_out[3][0], cfg.mid_dim, up_c[0])]
        nmax = [max(cfg.down_states[i], cfg.up_states[3 - i]) for i in range(4)]
        nmax[3] = max(nmax[3], 8 * 4)
        for l, ((h, w), C) in enumerate(zip(sizes, cmax)):
            P = h * w
            for name, ch in (("TA", C), ("TB", C), ("A", C), ("XZ", 4 * C), ("XS", 2 * C), ("DTS", 2 * C), ("YS", 2 * C),
                             ("G", 2 * C), ("QKV", 3 * C), ("QKV2", 3 * C), ("V", C), ("Y", C), ("SK", C)):
                self.buf(f"{name}{l}", B, P, ch)
            self.buf(f"H{l}", B, P, cfg.in_out[l][0])
            self.buf(f"WEFF{l}", B, C, C)
            self.buf(f"STAT{l}", B, P, 2, dtype=torch.float32)
            self.buf(f"BS{l}", B, 4, nmax[l], P // 4, dtype=torch.float32)
            self.buf(f"CS{l}", B, 4, nmax[l], P // 4, dtype=torch.float32)

        # init conv (fp32 images -> level-0 trunk); `r = x.clone()` (src/DADiff.py:701) lives in its own buffer R
        self.init_w, self.init_b = f32("init_conv.weight"), f32("init_conv.bias")
        init_w_host = sd["init_conv.weight"]
        self.x_t = torch.zeros(B, H * W, device=dev, dtype=torch.float32)
        self.x_input = torch.zeros(B, H * W, device=dev, dtype=torch.float32)
        R = self.buf("R", B, H * W, d)
        if R.dtype != torch.float32 and d == 64 and H % 8 == 0 and W % 16 == 0 and self.prefer_tc:
            self.init_w16 = upload(ops.pack_init_conv_weights(init_w_host), dev)
            self.steps.append(lambda: ops.init_conv7x7_tc(self.x_t, self.x_input, self.init_w16, self.init_b, R, B, H, W))
        else:
            self.steps.append(lambda: ops.init_conv7x7(self.x_t, self.x_input, self.init_w, self.init_b, R, B, H, W))

        def conv_plain(key_w, key_b, src, dst, h, w, k, stride=1, upsample=False):
            w_host = sd[key_w].permute(0, 2, 3, 1).contiguous()                   # (Cout, KH, KW, Cin)
            up4 = None
            if upsample and self.prefer_tc and dst.dtype != torch.float32:        # phase-summed 2x2 kernels, packed on the host
                up4 = upload(ops.pack_upsample_phases(w_host.to(self.trunk_dtype), w_host.shape[0], w_host.shape[-1]), dev)
            c = ops.Conv(src, upload(w_host, dev, self.trunk_dtype), dst, B=B, Hin=h, Win=w, KH=k, KW=k, stride=stride,
                         pad=1, upsample=upsample, bias=f32(key_b), prefer_tc=self.prefer_tc, weight_up4=up4)
            self.steps.append(c.run)

        cur = R
        for i, (ci, co) in enumerate(cfg.in_out):                                   # src/DADiff.py:712-719
            h, w = sizes[i]
            P = h * w
            ta = self.buf(f"TA{i}", B, P, ci)
            self._mamba(f"downs.{i}.1", i, cur, ta, ci, cfg.down_states[i], h, w)
            hs = self.buf(f"H{i}", B, P, ci)
            self._resblock(f"downs.{i}.0", i, [ta], hs, ci, h, w)
            if i < 3:
                nh, nw = sizes[i + 1]
                cur = self.buf(f"TA{i + 1}", B, nh * nw, co)
                conv_plain(f"downs.{i}.2.weight", f"downs.{i}.2.bias", hs, cur, h, w, 4, stride=2)
            else:
                cur = self.buf(f"TB{i}", B, P, co)
                conv_plain(f"downs.{i}.2.weight", f"downs.{i}.2.bias", hs, cur, h, w, 3)
        h, w = sizes[3]
        md = cfg.mid_dim
        x = self.buf("TA3", B, h * w, md)
        self._resblock("mid_block", 3, [cur], x, md, h, w)                         # :721
        self._mamba("mid_attn", 3, x, x, md, BASE_MID_STATE, h, w)                 # :722
        cur = x
        for i, (ci, co) in enumerate(reversed(cfg.in_out)):                         # :725-731
            l = 3 - i
            h, w = sizes[l]
            P = h * w
            hs = self.buf(f"H{l}", B, P, ci)
            tb = self.buf(f"TB{l}", B, P, co)
            self._resblock(f"ups.{i}.0", l, [cur, hs], tb, co, h, w)
            self._mamba(f"ups.{i}.1", l, tb, tb, co, cfg.up_states[i], h, w)
            if i < 3:
                nh, nw = sizes[l - 1]
                cur = self.buf(f"TA{l - 1}", B, nh * nw, ci)
                conv_plain(f"ups.{i}.2.1.weight", f"ups.{i}.2.1.bias", tb, cur, h, w, 3, upsample=True)
            else:
                cur = self.buf(f"TA{l}", B, P, ci)
                conv_plain(f"ups.{i}.2.weight", f"ups.{i}.2.bias", tb, cur, h, w, 3)
        h, w = sizes[0]
        self.feat = self.buf("TB0", B, h * w, d)
        self._resblock("final_res_block", 0, [cur, R], self.feat, d, h, w)         # :733-735
        self.final_w = upload(sd["final_conv.weight"].reshape(-1), dev)
        self.final_b = f32("final_conv.bias")

    # -------------------------------------------------------------------------------------------------------
    def _resblock(self, p, l, srcs, out, cout, h, w):
        """SiLU(GN8(WSConv3x3(x))) + res_conv(x)  (src/DADiff.py:213-229, 397-430); x = cat(srcs)."""
        sd, B, dt, dev = self.sd, self.B, self.dtype, self.device
        f32 = self.f32
        P = h * w
        cin = sum(s.shape[-1] for s in srcs)
        dt = self.trunk_dtype                  # reads and writes the residual stream only
        wc = pack_conv(ws_fold(sd[p + ".block1.proj.weight"]), dt, dev)
        bc = f32(p + ".block1.proj.bias")
        gamma, beta = f32(p + ".block1.norm.weight"), f32(p + ".block1.norm.bias")
        y = self.buf(f"Y{l}", B, P, cout)
        src1 = srcs[1] if len(srcs) > 1 else None
        acc = self._acc(B * GN_GROUPS * 2)
        # zeroed every forward with the other accumulators: slots + arrival counters of the reproducible GroupNorm sums
        acc_ws = self._acc(max(ops.conv_gn_ws_floats(B), ops.gn_stats_ws_floats(B, P, GN_GROUPS)))
        has_res = (p + ".res_conv.weight") in sd
        if has_res:
            sk = self.buf(f"SK{l}", B, P, cout)
            wr, br = pack_conv(sd[p + ".res_conv.weight"], dt, dev), f32(p + ".res_conv.bias")
        else:
            assert len(srcs) == 1 and cin == cout
            sk = srcs[0]
        for t in srcs:
            assert t.data_ptr() not in (y.data_ptr(), out.data_ptr()) and (not has_res or t.data_ptr() != sk.data_ptr())
        holder = {}

        def bind():
            sums = self._acc_view(acc, B, GN_GROUPS, 2)
            holder["sums"] = sums
            holder["conv"] = ops.Conv(srcs[0], wc, y, B=B, Hin=h, Win=w, KH=3, KW=3, pad=1, src1=src1, bias=bc,
                                      gn_sums=sums, gn_groups=GN_GROUPS, prefer_tc=self.prefer_tc,
                                      gn_ws=self._acc_view(acc_ws, self._acc_slices[acc_ws][1]))
            if has_res:
                holder["rconv"] = ops.Conv(srcs[0], wr, sk, B=B, Hin=h, Win=w, src1=src1, bias=br, prefer_tc=self.prefer_tc)
        self._acc_users.append(bind)

        def run():
            holder["conv"].run()
            if has_res:
                holder["rconv"].run()
            ops.gn_silu_add(y, holder["sums"], gamma, beta, sk, out, B, P, cout, GN_GROUPS)
        run._fd_name = f"fd.unet.{p} ResnetBlock {cin}->{cout} @{h}x{w}"
        self.steps.append(run)

    # -------------------------------------------------------------------------------------------------------
    def _mamba(self, p, l, x_in, x, C, N, h, w):
        """Mamba_block (src/DADiff.py:477-488).  Reads trunk `x_in`, leaves the result in trunk `x` (may alias)."""
        sd, B, dt, dev = self.sd, self.B, self.dtype, self.device
        f32 = self.f32
        P, D, L = h * w, 2 * C, h * w // 4
        R = math.ceil(C / 16)
        heads = C // 32
        mo, lo = self._mod_offsets[p], self._loc_offsets[p]
        MS = self.mod_total
        mods = self.mods
        sh1, sc1, g1, sh2, sc2, g2 = (_view_ptr(mods[:, mo + j * C:mo + (j + 1) * C]) for j in range(6))
        n1w, n1b = f32(p + ".norm1.weight"), f32(p + ".norm1.bias")
        a = self.buf(f"A{l}", B, P, C)
        xz = self.buf(f"XZ{l}", B, P, 4 * C)
        xs, dts, ys = (self.buf(f"{n}{l}", B, 4, D, L) for n in ("XS", "DTS", "YS"))
        Bs, Cs = self.buf(f"BS{l}", B, 4, N, L, dtype=torch.float32), self.buf(f"CS{l}", B, 4, N, L, dtype=torch.float32)
        g = self.buf(f"G{l}", B, P, D)
        stat = self.buf(f"STAT{l}", B, P, 2, dtype=torch.float32)
        qkv = self.buf(f"QKV{l}", B, P, 3 * C)
        v = self.buf(f"V{l}", B, P, C)
        weff = self.buf(f"WEFF{l}", B, C, C)
        to_dt = lambda t: upload(t, dev, dt)  # noqa: E731
        to_tdt = lambda t: upload(t, dev, self.trunk_dtype)  # noqa: E731  (operands of convs reading `a`)
        in_w = to_tdt(sd[p + ".mamba.in_proj.weight"])                                                 # (4C, C)
        out_w = to_dt(sd[p + ".mamba.out_proj.weight"])                                                # (C, 2C)
        dw_host = sd[p + ".mamba.conv2d.weight"].reshape(D, 9)
        dw_w, dw_b = upload(dw_host, dev), f32(p + ".mamba.conv2d.bias")
        xp_w, dtp_w = f32(p + ".mamba.x_proj_weight"), f32(p + ".mamba.dt_projs_weight")
        use_xdt_tc = dt != torch.float32 and R <= 32 and R + 2 * N <= 96
        if use_xdt_tc:
            xw16, dw16, Rp = ops.pack_xdt_weights(sd[p + ".mamba.x_proj_weight"], sd[p + ".mamba.dt_projs_weight"], dt)
            xw16, dw16 = upload(xw16, dev), upload(dw16, dev)
        dt_bias = upload(sd[p + ".mamba.dt_projs_bias"].reshape(-1), dev)
        A_neg = upload(-torch.exp(sd[p + ".mamba.A_logs"]), dev)                                      # emamba2.py:344
        Ds = f32(p + ".mamba.Ds")
        on_w, on_b = f32(p + ".mamba.out_norm.weight"), f32(p + ".mamba.out_norm.bias")
        qkv_w = to_tdt(sd[p + ".attn_blk.qkv.weight"].reshape(3 * C, C))
        qdw_host = sd[p + ".attn_blk.qkv_dwconv.weight"].reshape(3 * C, 9)
        qdw_w = upload(qdw_host, dev)
        qdw_wt = upload(qdw_host.t(), dev)                   # tap-major copy for the streaming dwconv kernel
        proj_w = upload(sd[p + ".attn_blk.project_out.weight"].reshape(C, C), dev)
        temp = upload(sd[p + ".attn_blk.temperature"].reshape(-1), dev)
        acc_g = self._acc(B * heads * 32 * 32)
        acc_q = self._acc(B * 2 * C)
        acc_gw = self._acc(ops.gram_ws_floats(B, h, w, C, dt))      # per-chunk Gram records + arrival counters (zeroed per forward)
        tc = self.prefer_tc
        holder = {}
        # LayerNorm + adaLN modulate folded into the 1x1 GEMM that reads it (fd_ln_fold + the statistics warps of conv_tc): the
        # GEMM reads the residual stream itself, the `a` tensor and both ln_modulate passes disappear.  The folded weights are
        # per sample and per step (B x Cout x C), so the fold only pays where a level has many more pixels than output channels.
        use_fold = (dt != torch.float32 and tc and P >= 8 * 4 * C and os.environ.get("FD_LN_FOLD", "1") == "1")
        c_in = c_qkv = None
        ext_rstd = False
        if use_fold:
            tdt = self.trunk_dtype
            in_w32 = upload(sd[p + ".mamba.in_proj.weight"], dev)
            qkv_w32 = upload(sd[p + ".attn_blk.qkv.weight"].reshape(3 * C, C), dev)
            wf_in = self.buf(f"WFIN.{p}", B, 4 * C, C, dtype=tdt)
            wf_qkv = self.buf(f"WFQKV.{p}", B, 3 * C, C, dtype=tdt)
            v_in = torch.zeros(B, 4 * C, device=dev, dtype=torch.float32)
            v_q = torch.zeros(B, 3 * C, device=dev, dtype=torch.float32)
            # Row statistics: the GEMM's own statistics warps (default) or a separate one-read pass (fd_row_rstd, 4 bytes per pixel,
            # FD_LN_RSTD=1: at C >= 128, =all: everywhere).  The pass won while the statistics warps accumulated in scalar fp32 and
            # were the critical path of the block (64 -> 256: 759 us against 665 + a 110 us pass); on packed fp32 pairs they are not
            # (627 us), and per level the two forms now measure equal or better for the in-kernel one, which needs no extra launch.
            ext_rstd = {"1": C in (128, 256, 512), "all": C in (64, 128, 256, 512)}.get(os.environ.get("FD_LN_RSTD", "0"), False)
            rstd = self.buf(f"RSTD{l}", B, P, dtype=torch.float32) if ext_rstd else None
            try:
                c_in = ops.Conv(x_in, wf_in, xz, B=B, Hin=h, Win=w, silu_from=2 * C, per_batch_weight=True, prefer_tc=True,
                                ln_v=v_in, ln_eps=1e-5, ln_rstd=rstd)
                c_qkv = ops.Conv(x, wf_qkv, qkv, B=B, Hin=h, Win=w, per_batch_weight=True, prefer_tc=True, ln_v=v_q, ln_eps=1e-6,
                                 ln_rstd=rstd)
            except Exception:                           # geometry does not tile for the tensor-core kernel: separate passes
                use_fold = False
        if not use_fold:
            c_in = ops.Conv(a, in_w, xz, B=B, Hin=h, Win=w, silu_from=2 * C, prefer_tc=tc)
            c_qkv = ops.Conv(a, qkv_w, qkv, B=B, Hin=h, Win=w, prefer_tc=tc)
        c_out = ops.Conv(g, out_w, x, B=B, Hin=h, Win=w, gate=g1, gate_stride=MS, addend=x_in, prefer_tc=tc)

        def ln1_in_proj():
            if use_fold:
                ops.ln_fold(in_w32, n1w, n1b, sh1, sc1, MS, wf_in, v_in, B, 4 * C, C)
                if ext_rstd:
                    ops.row_rstd(x_in, rstd, B * P, C, 1e-5)
            else:
                ops.ln_modulate(x_in, a, n1w, n1b, sh1, sc1, MS, B, P, C, 1e-5)
            c_in.run()

        def ln2_qkv():
            if use_fold:
                ops.ln_fold(qkv_w32, None, None, sh2, sc2, MS, wf_qkv, v_q, B, 3 * C, C)
                if ext_rstd:
                    ops.row_rstd(x, rstd, B * P, C, 1e-6)
            else:
                ops.ln_modulate(x, a, None, None, sh2, sc2, MS, B, P, C, 1e-6)
            c_qkv.run()
        # fused scan+merge needs 16-bit io, 16/32-byte aligned rows (L % 8 == 0) and d_state in {4, 8, 16, 32}
        fuse_merge = dt != torch.float32 and L % 8 == 0 and N in (4, 8, 16, 32) and D % 8 == 0
        # levels with a small dt_rank: x_proj alone, dt_proj applied inside the scan (no (B, 4D, L) delta tensor at all)
        fuse_dt = (fuse_merge and use_xdt_tc and (N, R) in ((4, 4), (8, 8)) and D % 32 == 0
                   and os.environ.get("FD_FUSE_DT", "0") == "1")
        # deep levels (many short rows, d_state >= 16): channel-per-lane scan, B / C time-major
        scan_cl = (fuse_merge and use_xdt_tc and N in (16, 32) and D % 32 == 0 and B * 4 * D >= 16384
                   and os.environ.get("FD_SCAN_CL", "1") == "1")       # d_state 8 measured slower this way (too few rows)
        if scan_cl:
            Bs_t, Cs_t = Bs.view(B, 4, L, N), Cs.view(B, 4, L, N)
        # bias + softplus finished by the x_proj/dt_proj kernel: measured negative for the warp-shuffle scan levels, twice (again
        # after the producer went to 3 blocks per SM: scans 18.44 -> 18.17 ms, producer 4.02 -> 4.81 ms per call) — the scan is bound
        # by its dependent shuffle / MUFU chains, not by issue slots — so only the channel-per-lane levels use it (FD_XDT_SOFTPLUS=1
        # switches it on for the others)
        xdt_softplus = fuse_merge and use_xdt_tc and not scan_cl and os.environ.get("FD_XDT_SOFTPLUS", "0") == "1"
        # 16-bit modes: the TIME-MAJOR core (fd_ss2d_tm.cu) — channel-per-lane segmented scan, dt_proj fused where the rank is small
        tm_fuse = (N, R) in ((4, 4), (8, 4), (8, 8), (16, 8))
        use_tm = (dt != torch.float32 and use_xdt_tc and D % 128 == 0 and h % 2 == 0 and w % 2 == 0 and (tm_fuse or N in (4, 8, 16, 32))
                  and os.environ.get("FD_SS2D_TM", "1") != "0")
        # Measured per level at B = 16 (profiles/r2_scan_tm_variants.json, profiles/r2_scan_tw2_notes.txt): the time-major chain wins at
        # every level since the time-sliced scan moved to packed arithmetic with its fix-up rows in tensor memory (the d_state-8,
        # D = 128 level was the last one on the round-1 warp-shuffle chain: 743 us against 705 now, and a cheaper dwconv / x_proj).
        # out_norm + gate + local + out_proj + gated residual in one pass (fd_ln_gate_gemm.cu) where the level is large and the
        # kernel exists (D = 128 -> C = 64: the full-resolution level, 4 of the 9 Mamba blocks and 2/3 of their tail traffic)
        fuse_tail = (use_tm and os.environ.get("FD_FUSE_TAIL", "1") == "1" and out_w.dtype == dt
                     and ops.ln_gate_out_proj_supported(P, D, C, 4 * C, 2 * C, dt, x.dtype))
        if use_tm:
            scan_cl = fuse_dt = False
            dw_wt = upload(dw_host.t(), dev)                               # tap-major (9, D)
            xs_tm, dts_tm = xs.view(B, 4, L, D), dts.view(B, 4, L, D)
            XR = (R + 2 * N) if tm_fuse else 2 * N
            xdbl_tm = self.buf(f"XDBLTM.{p}", B, 4, L, XR, dtype=torch.float32)
            S_tm = ops.scan_tm_plan(B, D, h, w, N, R if tm_fuse else 0)      # > 0 segments | -8 / -4 time-sliced kernel
            carry = self.buf(f"CARRY.{p}", B * 4 * max(S_tm, 1) * 2 * N * D, dtype=torch.float32)
            # time-sliced levels: rows cut into chained segments (short blocks refill the SMs; bit-identical results).  The workspace
            # holds the ticket counter, the hand-over flags and the carried states: zero-filled here once, every launch leaves it zeroed
            chain_n, chain_floats = ops.scan_tm_chain_plan(B, D, h, w, N, R) if (tm_fuse and S_tm < 0) else (0, 0)
            chain_ws = torch.zeros(chain_floats, device=dev, dtype=torch.float32) if chain_n > 1 else None
            dtw_tm = upload(sd[p + ".mamba.dt_projs_weight"].reshape(4 * D, R), dev)
        if fuse_dt:
            xdbl = self.buf(f"XDBL{l}", B, 4, R + 2 * N, L, dtype=torch.float32)
            dtw_flat = upload(sd[p + ".mamba.dt_projs_weight"].reshape(4 * D, R), dev)
        split_attn = dt != torch.float32            # 16-bit modes: streaming dwconv + tensor-core Gram, v read in place
        if split_attn:
            qkv2 = self.buf(f"QKV2{l}", B, P, 3 * C)
            v_view = qkv2[:, :, 2 * C:]
            c_att = ops.Conv(v_view, weff, x, B=B, Hin=h, Win=w, gate=g2, gate_stride=MS, addend=x, per_batch_weight=True,
                             prefer_tc=tc, c0=C, ld0=3 * C)
        else:
            c_att = ops.Conv(v, weff, x, B=B, Hin=h, Win=w, gate=g2, gate_stride=MS, addend=x, per_batch_weight=True,
                             prefer_tc=tc)
        local_c = _LocalView(self.locals, lo, D)
        self._local_views.append(local_c)

        def bind():
            holder["gram"] = self._acc_view(acc_g, B, heads, 32, 32)
            holder["qk"] = self._acc_view(acc_q, B, 2, C)
            holder["gws"] = self._acc_view(acc_gw, self._acc_slices[acc_gw][1])
        self._acc_users.append(bind)

        self.paths[p] = (((f"time-major scan, {S_tm} segment(s)" if S_tm > 0 else f"time-major scan, time-sliced x{-S_tm}") + (f", {chain_n} chained segments" if chain_n > 1 else ""))
                         + (", dt_proj fused" if tm_fuse else "") + (", LayerNorm folded into in_proj / qkv" if use_fold else "") + (", fused out_norm / out_proj tail" if fuse_tail else "") if use_tm else
                         "scan_cl time-major B/C" if scan_cl else "dt-fused warp scan" if fuse_dt else
                         "warp scan + merge" if fuse_merge else "reference-layout" if dt == torch.float32 else "warp scan, unfused merge")

        def run():
            ln1_in_proj()
            if use_tm:
                ops.dwconv3x3_silu_tm(xz, 4 * C, dw_wt, dw_b, xs_tm, B, h, w, D)
                if tm_fuse:
                    ops.x_proj_tm(xs_tm, xw16, xdbl_tm, None, None, None, B, D, L, R, N, Rp, True)
                    if chain_ws is not None:
                        ops.selective_scan_tm_chained(xs_tm, xdbl_tm, A_neg, dtw_tm, dt_bias, Ds, chain_ws, ys.view(B, P, D), B, D, h, w, N, R)
                    else:
                        ops.selective_scan_tm(xs_tm, None, xdbl_tm, A_neg, dtw_tm, dt_bias, Ds, carry, ys.view(B, P, D), B, D, h, w, N, R, S_tm)
                else:
                    ops.x_proj_tm(xs_tm, xw16, xdbl_tm, dw16, dts_tm, dt_bias, B, D, L, R, N, Rp, False)
                    ops.selective_scan_tm(xs_tm, dts_tm, xdbl_tm, A_neg, None, None, Ds, carry, ys.view(B, P, D), B, D, h, w, N, 0, S_tm)
                if fuse_tail:
                    ops.ln_gate_out_proj(ys.view(B, P, D), xz, 4 * C, 2 * C, on_w, on_b, local_c.dense(), out_w, g1, MS, x_in, x,
                                         B, P, D, C)
                else:
                    ops.ln_gate(ys.view(B, P, D), xz, 4 * C, 2 * C, on_w, on_b, local_c.dense(), g, B, P, D)
                    c_out.run()
                ln2_qkv()
                ops.dwconv3x3_nhwc(qkv, qdw_wt, None, qkv2, B, h, w, 3 * C)
                ops.gram_qk(qkv2, 3 * C, holder["gram"], holder["qk"], B, P, C, ws=holder["gws"])
                ops.attn_weff(holder["gram"], holder["qk"], temp, proj_w, weff, B, C)
                c_att.run()
                return
            ops.dwconv3x3_silu_scan(xz, 4 * C, dw_w, dw_b, xs, B, h, w, D)
            if scan_cl:
                ops.xdt_proj_tc(xs, xw16, dw16, Rp, dts, Bs_t, Cs_t, B, D, L, R, N, time_major=True, dt_bias=dt_bias, delta_softplus=True)
                ops.selective_scan_fwd_merge_cl(xs.view(B, 4 * D, L), dts.view(B, 4 * D, L), A_neg, Bs_t, Cs_t, Ds, None, False,
                                                ys.view(B, P, D), h, w)
                ops.ln_gate(ys.view(B, P, D), xz, 4 * C, 2 * C, on_w, on_b, local_c.dense(), g, B, P, D)
            elif fuse_dt:
                ops.x_proj_tc(xs, xw16, xdbl, B, D, L, R, N)
                ops.selective_scan_fwd_merge_xdbl(xs.view(B, 4 * D, L), xdbl, dtw_flat, A_neg, Ds, dt_bias, True, ys.view(B, P, D), h, w)
                ops.ln_gate(ys.view(B, P, D), xz, 4 * C, 2 * C, on_w, on_b, local_c.dense(), g, B, P, D)
            elif xdt_softplus:       # delta finished (bias + softplus) by the HBM-bound producer; the issue-bound scan takes it as is
                ops.xdt_proj_tc(xs, xw16, dw16, Rp, dts, Bs, Cs, B, D, L, R, N, dt_bias=dt_bias, delta_softplus=True)
            elif use_xdt_tc:
                ops.xdt_proj_tc(xs, xw16, dw16, Rp, dts, Bs, Cs, B, D, L, R, N)
            else:
                ops.xdt_proj(xs, xp_w, dtp_w, dts, Bs, Cs, B, D, L, R, N)
            if fuse_dt or scan_cl:
                pass
            elif fuse_merge:      # scan writes channels-last directly (EfficientMerge fused), then a row-wise LN + gate
                ops.selective_scan_fwd_merge(xs.view(B, 4 * D, L), dts.view(B, 4 * D, L), A_neg, Bs, Cs, Ds,
                                             None if xdt_softplus else dt_bias, not xdt_softplus, ys.view(B, P, D), h, w)
                ops.ln_gate(ys.view(B, P, D), xz, 4 * C, 2 * C, on_w, on_b, local_c.dense(), g, B, P, D)
            else:
                ops.selective_scan_fwd(xs.view(B, 4 * D, L), dts.view(B, 4 * D, L), A_neg, Bs, Cs, Ds, dt_bias, True,
                                       out=ys.view(B, 4 * D, L))
                ops.merge_ln_gate(ys, xz, 4 * C, 2 * C, on_w, on_b, local_c.dense(), stat, g, B, h, w, D)
            c_out.run()
            ln2_qkv()
            if split_attn:
                ops.dwconv3x3_nhwc(qkv, qdw_wt, None, qkv2, B, h, w, 3 * C)
                ops.gram_qk(qkv2, 3 * C, holder["gram"], holder["qk"], B, P, C, ws=holder["gws"])
            else:
                ops.dwconv3x3_qkv_gram(qkv, qdw_w, v, holder["gram"], holder["qk"], B, h, w, C, ws=holder["gws"])
            ops.attn_weff(holder["gram"], holder["qk"], temp, proj_w, weff, B, C)
            c_att.run()
        run._fd_name = f"fd.unet.{p} Mamba_block C{C} N{N} @{h}x{w}"
        self.steps.append(run)
        return x

    # -------------------------------------------------------------------------------------------------------
    def set_condition(self, dose_emb: torch.Tensor, ctx_emb: torch.Tensor):
        """Per-slice constants (computed once per sample() call, never per step): the prompt embedding
        prompt_mlp(softmax(text_mlp(dose)) * prompt) (src/DADiff.py:706-707) and every SS2D `local` vector
        SiLU(Linear(256 -> 2C)(ctx)) (src/emamba2.py:522-525, 715)."""
        B = self.B
        ops.linear_small(dose_emb.float().contiguous(), self.text_w0, self.text_b0, self.txt_hid, act_out=1)
        ops.linear_small(self.txt_hid, self.text_w2, self.text_b2, self.txt_out)
        sm = torch.softmax(self.txt_out, dim=1) * self.prompt
        ops.linear_small(sm.contiguous(), self.pm_w, self.pm_b, self.prompt_emb)
        ops.linear_small(ctx_emb.float().contiguous(), self.local_w, None, self.locals, act_out=1)
        for lv in self._local_views:
            lv.refresh()

    def conditioning(self):
        """Per-step: time embedding (+ prompt) and the 9 adaLN modulation vectors (src/DADiff.py:703-709, 484)."""
        ops.time_sinusoid(self.time, self.t_sin)
        ops.linear_small(self.t_sin, self.time_w1, self.time_b1, self.t_hid, act_out=2)
        ops.linear_small(self.t_hid, self.time_w2, self.time_b2, self.t_emb, add=self.prompt_emb)
        ops.linear_small(self.t_emb, self.adaln_w, self.adaln_b, self.mods, act_in=1)

    def forward(self):
        """One Unet evaluation on (self.x_t, self.x_input, self.time); the result is `self.feat`, the input of
        final_conv, which the sampler fuses with the update (ops.final_conv_update)."""
        ops.zero_(self.acc_buf)
        with ops.nvtx_range("fd.unet.conditioning"):
            self.conditioning()
        if ops.NVTX:
            for name, fn in zip(self.step_names, self.steps):
                with ops.nvtx_range(name):
                    fn()
        else:
            for fn in self.steps:
                fn()
        return self.feat


class _LocalView:
    """Dense (B, D) copy of one block's slice of the concatenated `locals` tensor (the kernel wants row stride D)."""

    def __init__(self, full, off, D):
        self.full, self.off, self.D = full, off, D
        self._dense = torch.zeros(full.shape[0], D, device=full.device, dtype=torch.float32)

    def refresh(self):
        self._dense.copy_(self.full[:, self.off:self.off + self.D])

    def dense(self):
        return self._dense


class _view_ptr:
    """Marks a strided fp32 view whose base pointer is handed to a kernel together with an explicit row stride."""

    def __init__(self, t: torch.Tensor):
        assert t.dtype == torch.float32 and t.is_cuda
        self.t = t
        self.dtype = torch.float32
        self.is_cuda = True

    def is_contiguous(self):
        return True

    def data_ptr(self):
        return self.t.data_ptr()
