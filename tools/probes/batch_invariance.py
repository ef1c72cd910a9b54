"""Where does a slice's result start to depend on the batch it is in?  Compares DA-CLIP embeddings, conditioning vectors and every
engine buffer after one Unet evaluation between a batch of 6 and the sub-batch [2:5] (same slices, same noise)."""
import sys

import torch

sys.path.insert(0, ".")
from founddiff_b200 import distributed as fdist, weights  # noqa: E402
from founddiff_b200.diffusion import ResidualDiffusion, UnetRes  # noqa: E402

dtn = sys.argv[1] if len(sys.argv) > 1 else "fp32"
dt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[dtn]
H = 128
sd = weights.random_state_dict(10)
model = UnetRes(dim=64, dim_mults=(1, 2, 4, 8), num_unet=1, condition=True, input_condition=False, objective='pred_res', test_res_or_noise='res')
model.load_state_dict({"unet0." + k: v for k, v in sd.items()})
model.compute_dtype = dt
d = ResidualDiffusion(model, image_size=H, timesteps=1000, sampling_timesteps=2, objective='pred_res', loss_type='l2', condition=True, sum_scale=0.01).cuda()
d.init()
d.use_cuda_graph = False
ldct = torch.rand(6, 1, H, H, generator=torch.Generator().manual_seed(5))
noise = fdist.global_noise(6, (1, H, H), 99)[0]


def run(sl):
    x = ldct[sl].cuda()
    B = x.shape[0]
    out = d.sample([x], batch_size=B, last=True, noise={"init": noise[sl]})[-1]
    eng = model.engine(B, H, H, x.device)
    snap = {k: v.clone() for k, v in eng._bufs.items()}
    snap["mods"], snap["locals"], snap["t_emb"], snap["prompt_emb"] = eng.mods.clone(), eng.locals.clone(), eng.t_emb.clone(), eng.prompt_emb.clone()
    dose, ctx = model.daclip(x.device).embed(eng.x_input.view(B, 1, H, H))
    snap["dose"], snap["ctx"], snap["out"] = dose.clone(), ctx.clone(), out.clone()
    return B, snap


B6, s6 = run(slice(0, 6))
B3, s3 = run(slice(2, 5))
for k in s3:
    a, b = s3[k], s6[k]
    if a.dim() == 1 or a.numel() % B3:
        na, nb = a.numel() // B3, b.numel() // B6
        if na != nb:
            continue
        a, b = a.view(B3, na), b.view(B6, nb)
    same = torch.equal(a.reshape(B3, -1), b.reshape(B6, -1)[2:5])
    if not same:
        diff = float((a.reshape(B3, -1).float() - b.reshape(B6, -1)[2:5].float()).abs().max())
        print(f"DIFF {k:24s} max abs {diff:.3e}")
print("done", dtn)
