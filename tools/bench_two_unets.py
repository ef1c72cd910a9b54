"""num_unet = 2 ('pred_res_noise', train.py:75-77) sample() timing at the bench geometry: sequential vs two-stream Unets.
Run on the GPU box:  python tools/bench_two_unets.py [--batch 16 --size 512]"""
import argparse
import json
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import weights  # noqa: E402
from founddiff_b200.diffusion import ResidualDiffusion, UnetRes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--iters", type=int, default=5)
a = ap.parse_args()

m = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=2, condition=True, objective='pred_res_noise', test_res_or_noise='res_noise')
sd = {"unet0." + k: v for k, v in weights.random_state_dict(10).items()}
sd.update({"unet1." + k: v for k, v in weights.random_state_dict(11).items()})
m.load_state_dict(sd)
d = ResidualDiffusion(m, image_size=a.size, sampling_timesteps=2, objective='pred_res_noise', condition=True, sum_scale=0.01,
                      test_res_or_noise='res_noise').cuda()
d.init()
g = torch.Generator().manual_seed(1)
ldct = torch.rand(a.batch, 1, a.size, a.size, generator=g).cuda()
noise = {"init": torch.randn(a.batch, 1, a.size, a.size, generator=g).cuda()}
res = {}
for conc in (False, True):
    d.concurrent_unets = conc
    for _ in range(3):
        out = d.sample([ldct], last=True, noise=noise)[-1]
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        out = d.sample([ldct], last=True, noise=noise)[-1]
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    res["two_streams" if conc else "sequential"] = dict(ms_per_call=ms, slices_per_s=a.batch / ms * 1e3)
    assert torch.isfinite(out).all()
res["peak_mem_GB"] = torch.cuda.max_memory_allocated() / 1e9
print(json.dumps(res))
