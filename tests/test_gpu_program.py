"""Step programs (SURVEY 8b `unet_step`): a recorded timestep replayed by the C runtime — through ctypes on a torch arena and by the
plain-C host of examples/c_host.c on cudaMalloc'ed memory — reproduces `ResidualDiffusion.sample` (src/DADiff.py:1367-1380) BIT FOR
BIT: same kernels, same launch order, same buffers, no Python between the launches."""
import math
import os
import struct
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _diffusion(state_dict, H, S, dt):
    from founddiff_b200.diffusion import ResidualDiffusion, UnetRes
    m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, input_condition=False, objective='pred_res', test_res_or_noise='res')
    m.load_state_dict({"unet0." + k: v for k, v in state_dict.items()})
    m.compute_dtype = dt
    d = ResidualDiffusion(m, image_size=H, timesteps=1000, sampling_timesteps=S, objective='pred_res', loss_type='l2', condition=True,
                          sum_scale=0.01).cuda()
    d.init()
    return d


@pytest.mark.parametrize("cfg", [(64, 2, torch.bfloat16, 2), (128, 3, torch.float16, 3), (64, 1000, torch.bfloat16, 2)])
def test_step_program_replays_sample_bit_for_bit(state_dict, tmp_path, cfg):
    from founddiff_b200 import distributed as fdist
    from founddiff_b200.program import StepProgram, export_step_program
    H, S, dt, B = cfg
    d = _diffusion(state_dict, H, S, dt)
    if S == 1000:
        d.num_timesteps = 5                                  # a 5-step ancestral chain (noise injected at every step but the last)
    ldct = torch.rand(B, 1, H, H, generator=torch.Generator().manual_seed(H + S)).cuda()
    sn = fdist.SliceNoise(7, range(B), (1, H, H), pin=False)
    noise = {"init": sn.init()}
    plan = d._step_plan()
    if S == 1000:
        noise["steps"] = torch.stack([sn._draw(torch.empty(B, 1, H, H)) for _ in range(len(plan) - 1)])
    ref = d.sample([ldct], batch_size=B, last=True, noise=noise)[-1]
    path = str(tmp_path / "plan.fdp")
    export_step_program(d, ldct, path, noise=noise)
    assert os.path.getsize(path) > 1 << 20

    prog = StepProgram(path)
    assert prog.num_launches() > 150
    from founddiff_b200 import ops
    P = H * H
    # x_input = 2 ldct - 1, x_T = x_input + sqrt(sum_scale) noise (src/DADiff.py:1294-1296, 1375) from the same kernel sample() uses
    x_in, x_t0, first = (torch.empty(B, P, device="cuda") for _ in range(3))
    ops.sampler_init(ldct.reshape(B, P).contiguous(), noise["init"].cuda().reshape(B, P).contiguous(), math.sqrt(d.sum_scale), x_in, x_t0, first)
    prog.buffer("x_input").copy_(x_in.reshape(-1))
    prog.buffer("x_t").copy_(x_t0.reshape(-1))
    steps = []
    for i, (t, c) in enumerate(plan):
        time = float(torch.tensor(d._sched("alphas_cumsum", t), dtype=torch.float32) * d.num_timesteps)      # as sample() forms it
        prog.buffer("time").copy_(torch.full((B,), time, device="cuda"))
        prog.buffer("coef").copy_(torch.tensor([*c, 0.], device="cuda"))
        nz = None
        if c[3] != 0.:
            nz = noise["steps"][i].reshape(-1)
            prog.buffer("noise").copy_(nz.cuda())
        steps.append((time, c, nz))
        prog.sample_step()
    got = torch.empty(B, P, device="cuda")
    ops.unnormalize(prog.buffer("x_t").reshape(B, P), got)
    diff = float((got.reshape(B, 1, H, H) - ref).abs().max())
    print(f"step program vs sample(): max abs diff {diff:.3e}")
    assert torch.equal(got.reshape(B, 1, H, H), ref), diff

    # ---- the same chain from plain C on cudaMalloc'ed memory
    exe = str(tmp_path / "c_host")
    libdir = os.path.join(ROOT, "founddiff_b200")
    r = subprocess.run(["gcc", "-std=c99", "-DFD_HOST_WITH_CUDA", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include",
                        os.path.join(ROOT, "examples", "c_host.c"), "-o", exe, "-L", libdir, "-lfounddiff_b200", "-L", "/usr/local/cuda/lib64",
                        "-lcudart", f"-Wl,-rpath,{libdir}", "-Wl,-rpath,/usr/local/cuda/lib64"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    inp, outp = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(inp, "wb") as f:
        f.write(struct.pack("<iii", len(steps), B, P))
        f.write(x_in.cpu().numpy().astype(np.float32).tobytes())
        f.write(x_t0.cpu().numpy().astype(np.float32).tobytes())
        for time, c, nz in steps:
            f.write(struct.pack("<f8fi", np.float32(time), *[np.float32(v) for v in c], 0., 1 if nz is not None else 0))
            if nz is not None:
                f.write(nz.numpy().astype(np.float32).tobytes())
    r = subprocess.run([exe, path, inp, outp], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    c_out = torch.from_numpy(np.fromfile(outp, dtype=np.float32)).reshape(B, 1, H, H)
    assert torch.equal(c_out, prog.buffer("x_t").reshape(B, 1, H, H).cpu()), "C host and ctypes replay differ"
