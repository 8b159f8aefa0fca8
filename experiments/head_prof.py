"""kernel durations (CUPTI) of the 3 -> 64 head convolution at 16 x 64 x 64: eager launches, L2-warm input"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import ops
from torch.profiler import profile, ProfilerActivity
dev = "cuda"
d, h, w = 16, 64, 64
g3 = torch.randn(1, 3, d, h, w, device=dev)
wh = torch.randn(64, 3, 3, 3, 3, device=dev) * 0.1
b64 = torch.zeros(64, device=dev)
stats = torch.zeros(128, device=dev)
for _ in range(3):
    ops.conv_raw(g3, wh, b64, 1, False, True)
    ops.conv_raw(g3, wh, b64, 1, False, True, stats=stats)
x = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
for _ in range(2):
    ops.wgrad_raw(g3, x, 1, (64, 3, 3, 3, 3), want_bias=True)
    ops.wgrad_raw(x, g3, 1, (3, 64, 3, 3, 3), want_bias=True)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        ops.conv_raw(g3, wh, b64, 1, False, True)
    for _ in range(5):
        ops.conv_raw(g3, wh, b64, 1, False, True, stats=stats)
    for _ in range(3):
        ops.wgrad_raw(g3, x, 1, (64, 3, 3, 3, 3), want_bias=True)
    for _ in range(3):
        ops.wgrad_raw(x, g3, 1, (3, 64, 3, 3, 3), want_bias=True)
    torch.cuda.synchronize()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        print("%8.2f us  %s" % (e.time_range.end - e.time_range.start, e.name[:90]))
