"""DA-CLIP conditioning (dose / anatomy embeddings), computed ONCE per slice and cached across timesteps.

The reference recomputes `CLIPIQA.forward` inside every `Unet.forward` (src/DADiff.py:692), including a 12-layer
text transformer whose only consumer (`probs`) is discarded.  Both embeddings are pure functions of the low-dose
input slice (channel 1 of the Unet input), so the sampler calls `embed()` once per `sample()` call and the
per-step path never sees the encoder (SURVEY.md "Five facts" #4).

This is off the per-step hot path; it runs the RN50 `ModifiedResNet` visual tower + `AttentionPool2d`
(pos_embedding=False) + head1/head2 (src/DACLIP.py:168-349, 1180-1211).

* fp32 mode: library convolutions (cuDNN, TF32 disabled) — the validation path.
* 16-bit modes: the 16 bottleneck blocks (94 % of the tower's FLOPs: 1x1 / 3x3 convolutions over 64..2048 channels) run
  on this repo's tcgen05 implicit-GEMM kernel (fd_conv2d_tc) with BatchNorm folded into weights + bias, ReLU and the
  residual add fused into the epilogue and channels-last activations; only the 3-convolution stem (3 / 32 input
  channels, below the 64-channel K block) and the attention pool stay on the library path.  At 16 x 512^2 this takes the
  embedding from 7.9 ms (cuDNN bf16) to the figure in DESIGN.md.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import torch
import torch.nn.functional as F

from . import ops

PREFIX = "dose_encoder."


def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


class DAClipEncoder:
    """Holds the frozen visual-tower weights on one device and maps (B,1,H,W) slices in [-1,1] to
    (dose_embedding (B,1024), context_embedding (B,256))."""

    def __init__(self, state_dict: Dict[str, torch.Tensor], device, layers: Sequence[int] = (3, 4, 6, 3), heads: int = 32,
                 conv_dtype: torch.dtype = torch.float32):
        # conv_dtype: storage/compute type of the RN50 convolutions.  bf16 (tensor-core cuDNN, channels_last) is used
        # by the 16-bit sampling modes: it perturbs the embeddings by ~6e-3 and the Unet output by ~1.6e-4 rel-L2
        # (measured with the oracle, DESIGN.md "Precision"), far below the 16-bit path's own rounding.
        self.conv_dtype = conv_dtype
        self.device = torch.device(device)
        # BatchNorm folding, layout permutes and 16-bit conversion run on HOST copies (one upload per finished tensor)
        self.sd_host = {k: v.detach().to("cpu", torch.float32) for k, v in state_dict.items()
                        if k.startswith(PREFIX) and v.is_floating_point()}
        self.sd = {k: v.to(self.device) for k, v in self.sd_host.items() if ".visual.attnpool." in k or ".head" in k}
        self.layers, self.heads = tuple(layers), heads
        # BatchNorm (eval) folded into the preceding bias-free convolution once, at load time
        self._folded = {}
        self._towers = {}
        self.use_tc = conv_dtype != torch.float32      # tcgen05 bottleneck tower (16-bit storage only)

    def _conv_bn(self, x, conv, bn, stride=1, padding=0):
        key = conv
        if key not in self._folded:
            sd = self.sd_host
            w = sd[conv + ".weight"]
            s = sd[bn + ".weight"] * torch.rsqrt(sd[bn + ".running_var"] + 1e-5)
            wf = (w * s[:, None, None, None]).to(self.conv_dtype).to(self.device).contiguous(memory_format=torch.channels_last)
            self._folded[key] = (wf, (sd[bn + ".bias"] - sd[bn + ".running_mean"] * s).to(self.conv_dtype).contiguous().to(self.device))
        w, b = self._folded[key]
        return F.conv2d(x, w, b, stride=stride, padding=padding)

    def _bottleneck(self, p, x, stride):
        out = F.relu(self._conv_bn(x, p + "conv1", p + "bn1"))
        out = F.relu(self._conv_bn(out, p + "conv2", p + "bn2", padding=1))
        if stride > 1:
            out = F.avg_pool2d(out, stride)
        out = self._conv_bn(out, p + "conv3", p + "bn3")
        if (p + "downsample.0.weight") in self.sd_host:
            idt = F.avg_pool2d(x, stride) if stride > 1 else x
            idt = self._conv_bn(idt, p + "downsample.0", p + "downsample.1")
        else:
            idt = x
        return F.relu(out + idt)

    # ---------------------------------------------------------------------------------------------------
    # tensor-core tower (16-bit modes)
    def _folded_nhwc(self, conv, bn):
        """BN-folded weight as (Cout, KH, KW, Cin) 16-bit + fp32 bias, the layout fd_conv_params takes."""
        key = "nhwc:" + conv
        if key not in self._folded:
            sd = self.sd_host
            w = sd[conv + ".weight"]
            s = sd[bn + ".weight"] * torch.rsqrt(sd[bn + ".running_var"] + 1e-5)
            wf = (w * s[:, None, None, None]).permute(0, 2, 3, 1).to(self.conv_dtype).contiguous().to(self.device)
            self._folded[key] = (wf, (sd[bn + ".bias"] - sd[bn + ".running_mean"] * s).float().contiguous().to(self.device))
        return self._folded[key]

    def _build_tower(self, B: int, H: int, W: int, dev):
        """Buffers + convolution plans of layer1..layer4 for a (B, 64, H, W) stem output (H, W = input / 4)."""
        v = PREFIX + "clip_model.visual."
        dt = self.conv_dtype
        steps, keep = [], []
        x = torch.empty(B, H * W, 64, device=dev, dtype=dt)
        tower_in = x
        cin, h, w = 64, H, W
        for li, blocks in enumerate(self.layers):
            planes = 64 * 2 ** li
            for bi in range(blocks):
                p = f"{v}layer{li + 1}.{bi}."
                stride = 2 if (li > 0 and bi == 0) else 1
                ho, wo = h // stride, w // stride
                new = lambda hh, ww, c: torch.empty(B, hh * ww, c, device=dev, dtype=dt)  # noqa: E731
                t1, t2 = new(h, w, planes), new(h, w, planes)
                out = new(ho, wo, planes * 4)
                w1, b1 = self._folded_nhwc(p + "conv1", p + "bn1")
                w2, b2 = self._folded_nhwc(p + "conv2", p + "bn2")
                w3, b3 = self._folded_nhwc(p + "conv3", p + "bn3")
                c1 = ops.Conv(x, w1, t1, B=B, Hin=h, Win=w, bias=b1, relu_out=True)
                c2 = ops.Conv(t1, w2, t2, B=B, Hin=h, Win=w, KH=3, KW=3, pad=1, bias=b2, relu_out=True)
                seq = [c1.run, c2.run]
                t2p = t2
                if stride > 1:
                    t2p = new(ho, wo, planes)
                    seq.append(lambda a=t2, o=t2p, hh=h, ww=w, c=planes: ops.avgpool2x2_nhwc(a, o, B, hh, ww, c))
                idt = x
                if (p + "downsample.0.weight") in self.sd_host:
                    wd, bd = self._folded_nhwc(p + "downsample.0", p + "downsample.1")
                    src = x
                    if stride > 1:
                        src = new(ho, wo, cin)
                        seq.append(lambda a=x, o=src, hh=h, ww=w, c=cin: ops.avgpool2x2_nhwc(a, o, B, hh, ww, c))
                    idt = new(ho, wo, planes * 4)
                    cd = ops.Conv(src, wd, idt, B=B, Hin=ho, Win=wo, bias=bd)
                    seq.append(cd.run)
                    keep.append(cd)
                c3 = ops.Conv(t2p, w3, out, B=B, Hin=ho, Win=wo, bias=b3, addend=idt, relu_out=True)
                seq.append(c3.run)
                keep += [c1, c2, c3]
                for c in (c1, c2, c3):
                    if not c.uses_tc:
                        raise RuntimeError("DA-CLIP tower: convolution not eligible for the tcgen05 path: " + c.describe())
                steps += seq
                x, cin, h, w = out, planes * 4, ho, wo
        return dict(inp=tower_in, out=x, steps=steps, keep=keep, h=h, w=w, c=cin)

    def _tower_tc(self, x: torch.Tensor):
        """x: un-pooled stem output (B, 64, 2H, 2W), channels-last -> layer4 output as (B, C, h, w) view of channels-last
        data, or None when the size does not tile for the tensor-core kernel."""
        B, C, H2, W2 = x.shape
        H, W = H2 // 2, W2 // 2
        key = (B, H, W, str(x.device))
        tw = self._towers.get(key)
        if tw is None:
            self._towers.clear()
            try:
                tw = self._build_tower(B, H, W, x.device)
            except RuntimeError:
                tw = False                                 # some level does not tile (small / odd sizes): library path
            self._towers[key] = tw
        if tw is False:
            return None
        xl = x.permute(0, 2, 3, 1)
        if not xl.is_contiguous():
            xl = xl.contiguous()
        ops.avgpool2x2_nhwc(xl, tw["inp"], B, H2, W2, C)      # the stem's AvgPool2d(2), written into the tower's input buffer
        for fn in tw["steps"]:
            fn()
        return tw["out"].view(B, tw["h"], tw["w"], tw["c"]).permute(0, 3, 1, 2)

    # Library kernels (cuDNN stem, cuBLAS attention pool / heads) are launched on FIXED-size groups of slices, the last group
    # zero-padded: cuDNN / cuBLAS choose algorithms (tilings, split-K) from the problem size, so a slice's embedding could change
    # in the last bits with the batch it is in.  With a fixed group size every launch has the same shape, and within one launch a
    # slice's result does not depend on its position or on the other slices.  The tcgen05 tower is per-sample by construction.
    # Measured on B200 (tools/probes/batch_invariance.py): the bf16 stem + the fp32 pool / heads are position-independent inside a
    # group of 4; the fp32 cuDNN tower of the validation mode is not (4e-8 on the embeddings), so fp32 mode runs slice by slice.
    LIB_GROUP = 4

    def _groups(self, x: torch.Tensor):
        import os
        n = x.shape[0]
        g = int(os.environ.get("FD_DACLIP_GROUP", "0")) or (1 if self.conv_dtype == torch.float32 else self.LIB_GROUP)
        for a in range(0, n, g):
            c = x[a:a + g]
            if c.shape[0] < g:
                c = torch.cat([c, c.new_zeros((g - c.shape[0],) + tuple(c.shape[1:]))], dim=0)
            yield min(g, n - a), c

    def _stem(self, x_input: torch.Tensor) -> torch.Tensor:
        v = PREFIX + "clip_model.visual."
        outs = []
        for n, c in self._groups(x_input):
            x = c.to(self.conv_dtype).repeat(1, 3, 1, 1).contiguous(memory_format=torch.channels_last)   # src/DADiff.py:692
            x = F.relu(self._conv_bn(x, v + "conv1", v + "bn1", stride=2, padding=1))
            x = F.relu(self._conv_bn(x, v + "conv2", v + "bn2", padding=1))
            x = F.relu(self._conv_bn(x, v + "conv3", v + "bn3", padding=1))
            outs.append(x[:n])
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def _tower_lib(self, x: torch.Tensor) -> torch.Tensor:
        v = PREFIX + "clip_model.visual."
        outs = []
        for n, c in self._groups(x):
            c = F.avg_pool2d(c, 2)
            for li, blocks in enumerate(self.layers):
                for bi in range(blocks):
                    c = self._bottleneck(f"{v}layer{li + 1}.{bi}.", c, 2 if (li > 0 and bi == 0) else 1)
            outs.append(c[:n])
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    def _pool_heads(self, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        sd = self.sd
        a = PREFIX + "clip_model.visual.attnpool."
        B, C, H, W = x.shape
        tok = x.reshape(B, C, H * W).permute(2, 0, 1)
        tok = torch.cat([tok.mean(dim=0, keepdim=True), tok], dim=0)          # (HW+1, B, C)
        # AttentionPool2d (src/DACLIP.py:168-211) with ONE query (the mean token).  k_proj / v_proj are linear, so they
        # are applied to the query / to the attention-weighted token sum instead of to all HW+1 tokens (softmax weights
        # sum to 1, so the v bias passes through): identical result, ~70x fewer FLOPs than projecting every token —
        # as fp32 SIMT GEMMs (TF32 off) those two projections were 1.3 ms of the 4.6 ms embedding.
        hd = C // self.heads
        q = F.linear(tok[0], sd[a + "q_proj.weight"], sd[a + "q_proj.bias"]).reshape(B, self.heads, hd) * (hd ** -0.5)
        wk = sd[a + "k_proj.weight"].reshape(self.heads, hd, C)
        wv = sd[a + "v_proj.weight"].reshape(self.heads, hd, C)
        qk = torch.einsum("bhd,hdc->bhc", q, wk)                               # W_k^T q per head
        qb = torch.einsum("bhd,hd->bh", q, sd[a + "k_proj.bias"].reshape(self.heads, hd))
        att = (torch.einsum("bhc,tbc->bht", qk, tok) + qb[..., None]).softmax(dim=-1)
        xbar = torch.einsum("bht,tbc->bhc", att, tok)                          # attention-weighted token sum per head
        o = (torch.einsum("bhc,hdc->bhd", xbar, wv) + sd[a + "v_proj.bias"].reshape(self.heads, hd)).reshape(1, B, C)
        feat = F.linear(o, sd[a + "c_proj.weight"], sd[a + "c_proj.bias"])[0]
        h1 = F.linear(F.relu(F.linear(feat, sd[PREFIX + "head1.0.weight"], sd[PREFIX + "head1.0.bias"])),
                      sd[PREFIX + "head1.2.weight"], sd[PREFIX + "head1.2.bias"])
        h2 = F.linear(F.relu(F.linear(feat, sd[PREFIX + "head2.0.weight"], sd[PREFIX + "head2.0.bias"])),
                      sd[PREFIX + "head2.2.weight"], sd[PREFIX + "head2.2.bias"])
        dose = h1 / h1.norm(dim=-1, keepdim=True).clamp_min(1e-30)             # src/DACLIP.py:1210 (the clamp only guards zero padding)
        ctx = F.normalize(h2, dim=1)                                           # src/DACLIP.py:1207
        return dose, ctx

    @torch.no_grad()
    def embed(self, x_input: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            x = self._stem(x_input)
            y = None
            if self.use_tc and x.shape[2] % 16 == 0 and x.shape[3] % 16 == 0:
                y = self._tower_tc(x)                     # pools the (channels-last) stem output itself
            x = y if y is not None else self._tower_lib(x)
            x = x.float()
            dose, ctx = [], []
            for n, c in self._groups(x):
                d, cx = self._pool_heads(c)
                dose.append(d[:n])
                ctx.append(cx[:n])
            dose = dose[0] if len(dose) == 1 else torch.cat(dose, dim=0)
            ctx = ctx[0] if len(ctx) == 1 else torch.cat(ctx, dim=0)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
        return dose.contiguous(), ctx.contiguous()
