"""Times DAClipEncoder.embed at the bench shape (16 x 512^2) for the library path and the tensor-core tower."""
import torch
from founddiff_b200 import weights
from founddiff_b200.daclip import DAClipEncoder

sd = weights.random_state_dict(seed=10)
x = torch.rand(16, 1, 512, 512, device="cuda")
for tc in (False, True):
    enc = DAClipEncoder(sd, torch.device("cuda"), conv_dtype=torch.bfloat16)
    enc.use_tc = tc
    for _ in range(3):
        enc.embed(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        enc.embed(x)
    e1.record()
    torch.cuda.synchronize()
    print(f"daclip embed, tensor-core tower={tc}: {e0.elapsed_time(e1) / 10:.3f} ms", flush=True)

from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        enc.embed(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
