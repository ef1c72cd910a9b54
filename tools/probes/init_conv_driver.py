import torch, sys
sys.path.insert(0, '.')
from founddiff_b200 import ops
B, H, W = 16, 512, 512
g = torch.Generator(device="cuda").manual_seed(0)
x_t = torch.randn(B, H * W, device="cuda", generator=g)
x_in = torch.rand(B, H * W, device="cuda", generator=g)
w = torch.randn(64, 2, 7, 7, device="cuda", generator=g) / 10
bias = torch.randn(64, device="cuda", generator=g)
out = torch.empty(B, H * W, 64, device="cuda", dtype=torch.float16)
w16 = ops.pack_init_conv_weights(w)
for _ in range(3):
    ops.init_conv7x7_tc(x_t, x_in, w16, bias, out, B, H, W)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(20):
    ops.init_conv7x7_tc(x_t, x_in, w16, bias, out, B, H, W)
e1.record()
torch.cuda.synchronize()
print("init_conv7x7_tc us per launch:", e0.elapsed_time(e1) / 20 * 1e3)
