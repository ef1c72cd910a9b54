import torch, sys
sys.path.insert(0, '.')
from founddiff_b200 import ops
B, KD, L, N, H, W = 16, 512, 65536, 4, 512, 512
g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s, d=torch.bfloat16: torch.randn(*s, device="cuda", generator=g).to(d)
u, delta = rn(B, KD, L), rn(B, KD, L) * 0.5
A = -torch.exp(torch.randn(KD, N, device="cuda", generator=g) * 0.3)
Bm, Cm = rn(B, 4, N, L, d=torch.float32), rn(B, 4, N, L, d=torch.float32)
D, bias = rn(KD, d=torch.float32), rn(KD, d=torch.float32)
y = torch.empty(B, H * W, KD // 4, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.selective_scan_fwd_merge(u, delta, A, Bm, Cm, D, bias, True, y, H, W)
torch.cuda.synchronize()
