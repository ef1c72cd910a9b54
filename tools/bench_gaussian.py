"""Times one Unet evaluation and a DDIM-2 sample() of the SECONDARY path (lucidrains Unet + GaussianDiffusion) at the
BASELINE geometry (batch 16 of 512 x 512, 3 channels, fp16 storage).  Informational: the judged metric is bench.py."""
import sys
import time

import torch

sys.path.insert(0, ".")
from founddiff_b200 import ops  # noqa: E402
from founddiff_b200.gaussian import GaussianDiffusion, Unet  # noqa: E402

B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 512
model = Unet(dim=64, dim_mults=(1, 2, 4, 8)).cuda()
model.compute_dtype = torch.float16
diff = GaussianDiffusion(model, image_size=S, timesteps=1000, sampling_timesteps=2, loss_type='l1').cuda()
for _ in range(2):
    diff.sample(batch_size=B)
torch.cuda.synchronize()
eng = model.engine(B, S, S, diff.betas.device)
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
n0 = ops.LAUNCHES
e0.record()
for _ in range(5):
    eng.forward()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"secondary path Unet evaluation, B={B} {S}x{S} fp16: {ms:.2f} ms ({ms / B * 1e3:.0f} us per slice-step, "
      f"{(ops.LAUNCHES - n0) // 5} launches; 951.5 GFLOP per slice-step -> {951.5e9 * B / ms / 1e9:.0f} TFLOP/s)")
e0.record()
for _ in range(3):
    diff.sample(batch_size=B)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"secondary path DDIM-2 sample(): {ms:.1f} ms per call, {B / ms * 1e3:.1f} images/s; peak memory "
      f"{torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
# breakdown: graph replay alone vs eager step vs full sample() in both modes
entry = diff._graphs.get("step") if diff._graphs else None          # (graph, engine) since the round-1 review's fix
replay = entry[0].replay if entry else None
if replay is not None:
    e0.record()
    for _ in range(5):
        replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"graph replay of one timestep: {e0.elapsed_time(e1) / 5:.2f} ms")
for mode in (False, True):
    diff.use_cuda_graph = mode
    diff.sample(batch_size=B)
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for _ in range(3):
        diff.sample(batch_size=B)
    e1.record()
    torch.cuda.synchronize()
    print(f"sample() graph={mode}: {e0.elapsed_time(e1) / 3:.1f} ms (wall {1e3 * (time.time() - t0) / 3:.1f} ms)")
