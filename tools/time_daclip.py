import torch, time
from founddiff_b200 import weights
from founddiff_b200.daclip import DAClipEncoder
sd = weights.random_state_dict(seed=10)
enc = DAClipEncoder({k: v for k, v in sd.items()}, torch.device("cuda"), conv_dtype=torch.bfloat16)
x = torch.rand(16, 1, 512, 512, device="cuda")
for _ in range(3): enc.embed(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10): enc.embed(x)
e1.record(); torch.cuda.synchronize()
print("daclip embed ms", e0.elapsed_time(e1) / 10)
