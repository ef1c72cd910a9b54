"""Round-2 feasibility probe (CPU, oracle only): how far does the carried state of the selective scan reach?

Splitting a row of L steps into S segments that are scanned independently from h = 0 (so that the 8 K rows of the full-resolution
levels become enough rows for the channel-per-lane kernel) needs a fix-up  y_l += sum_n C_n,l * exp(A_n * sum_{j<=l} dt_j) * H_n  for
the steps of a segment where the decay factor of the carried state H is still above fp32 noise.  This script measures, on the
reference-initialised weights (A_logs = log(1..N), dt bias per the reference initialiser) and the synthetic CT-like input, the
number of steps after which exp(A_n * cumsum(dt)) < 1e-7, per (channel, state), for the first Mamba block at 256^2 (L = 16384).

    python tools/probes/scan_carry_study.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from founddiff_b200 import weights  # noqa: E402
from oracle import founddiff_oracle as O  # noqa: E402
from oracle.gen_golden import synth_slices  # noqa: E402

torch.manual_seed(0)
sd = weights.random_state_dict(10)
H = W = 256
_, ldct = synth_slices(1, H, W)
x_input = ldct * 2 - 1
x_t = x_input + 0.1 * torch.randn(1, 1, H, W, generator=torch.Generator().manual_seed(1))
sched = O.make_schedule(1000, "init")
taps = {}
with torch.no_grad():
    O.unet_forward(sd, torch.cat((x_t, x_input), 1), (sched["alphas_cumsum"][999] * 1000).expand(1), taps=taps)
p = "downs.0.1.mamba"
ss = taps.get("ss2d." + p) or taps.get(p)
if ss is None:                                   # the Unet taps do not keep SS2D internals: evaluate the block directly
    x0 = taps["init_conv"]
    c = torch.nn.functional.normalize(torch.randn(1, 1, 256, generator=torch.Generator().manual_seed(2)), dim=-1)
    t = O.time_embedding(sd, (sched["alphas_cumsum"][999] * 1000).expand(1), 64)
    ss = {}
    with torch.no_grad():
        O.mamba_block(sd, "downs.0.1", x0, c, t, taps=ss)
dts = ss["dts"]                                                       # (1, 4, D, L) before bias / softplus
bias = sd[p + ".dt_projs_bias"].reshape(1, 4, -1, 1)
dt = torch.nn.functional.softplus(dts + bias)
A = -torch.exp(sd[p + ".A_logs"].float()).reshape(4, -1, sd[p + ".A_logs"].shape[1])      # (4, D, N)
L = dt.shape[-1]
print(f"L = {L}, dt: median {dt.median():.4f}, 1% {dt.flatten().kthvalue(int(0.01 * dt.numel())).values:.5f}, "
      f"99% {dt.flatten().kthvalue(int(0.99 * dt.numel())).values:.4f};  A in [{A.min():.2f}, {A.max():.2f}]")
S = 8
seg = L // S
horizons = []
for s in range(1, S):                                                 # carry into segments 1..S-1
    cs = torch.cumsum(dt[0, :, :, s * seg:(s + 1) * seg], dim=-1)      # (4, D, seg)
    need = (-16.1 / A).unsqueeze(-1)                                  # cumsum(dt) at which exp(A * cumsum) = 1e-7   (4, D, N, 1)
    reach = (cs.unsqueeze(2) < need).sum(-1)                          # steps of the segment the carry still matters in
    horizons.append(reach.float())
hz = torch.stack(horizons)                                            # (S-1, 4, D, N)
q = lambda f: hz.flatten().kthvalue(max(1, int(f * hz.numel()))).values.item()  # noqa: E731
print(f"carry horizon in steps (segment length {seg}): median {q(0.5):.0f}, 90% {q(0.9):.0f}, 99% {q(0.99):.0f}, max {hz.max():.0f}")
for n in range(A.shape[-1]):
    print(f"  state {n} (A = {A[0, 0, n]:.2f} ...): mean horizon {hz[..., n].mean():.0f} steps = {100 * hz[..., n].mean() / seg:.1f} % of a segment")
print(f"fix-up work = {100 * hz.mean() / seg * (S - 1) / S:.1f} % of the state updates of the plain scan (1 MUFU + 2 FMA each)")
