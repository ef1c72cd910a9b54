"""Prints the measured parity numbers quoted in DESIGN.md section 2 (Unet forward on the 64x96 golden fixture)."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from founddiff_b200 import weights
from founddiff_b200.diffusion import UnetRes

g = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(ROOT, "tests", "golden", "unet_64x96.npz")).items()}
sd = weights.random_state_dict(10)
model = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, input_condition=False, objective='pred_res',
                test_res_or_noise='res')
model.load_state_dict({"unet0." + k: v for k, v in sd.items()})
model = model.cuda()
rel = lambda a, b: float((a.float().cpu() - b).norm() / b.norm())
time = g["t999.time"].cuda()
for name, cd, td in (("fp32", torch.float32, None), ("fp16", torch.float16, None), ("bf16 (fp16 residual stream, default)", torch.bfloat16, None),
                     ("pure bf16", torch.bfloat16, torch.bfloat16)):
    model.compute_dtype, model.trunk_dtype = cd, td
    errs = [rel(model(g["x_in"].cuda(), [time, time])[0], g["t999.out"]) for _ in range(3)]
    print(f"{name:40s} rel-L2 vs reference: " + ", ".join(f"{e:.3e}" for e in errs), flush=True)
