"""graphed dependent sequences of the narrow-end kernels at 16 x 64 x 64 (development aid)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import lib, ops
dev = "cuda"
d, h, w = 16, 64, 64
x = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
g3 = torch.randn(1, 3, d, h, w, device=dev)
wt = torch.randn(3, 64, 3, 3, 3, device=dev) * 0.03
wh = torch.randn(64, 3, 3, 3, 3, device=dev) * 0.1
b3 = torch.zeros(3, device=dev)
b64 = torch.zeros(64, device=dev)


def timed(name, body, n):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        body()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        body()
    for _ in range(3):
        gr.replay()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("%-44s %8.1f us per sequence, %6.2f us per element" % (name, ts[10], ts[10] / n), flush=True)


timed("tail 64->3 x10", lambda: [ops.conv_raw(x, wt, b3, 1, False, False) for _ in range(10)], 10)
timed("head dgrad 64->3 x10", lambda: [ops.conv_raw(x, wh, None, 1, True, False) for _ in range(10)], 10)
timed("head 3->64 x10", lambda: [ops.conv_raw(g3, wh, b64, 1, False, True) for _ in range(10)], 10)
timed("tail dgrad 3->64 x10", lambda: [ops.conv_raw(g3, wt, None, 1, True, True) for _ in range(10)], 10)
timed("head wgrad x10", lambda: [ops.wgrad_raw(g3, x, 1, (64, 3, 3, 3, 3), want_bias=True) for _ in range(10)], 10)
timed("tail wgrad x10", lambda: [ops.wgrad_raw(x, g3, 1, (3, 64, 3, 3, 3), want_bias=True) for _ in range(10)], 10)
