"""Per-launch cost of kernels inside dependent sequences (CUDA graph, no host gaps): is there a price for alternating the 222 KB
tcgen05 convolution with small-footprint element-wise kernels (shared-memory carve-out changes, instruction cache)?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import lib, ops

dev = "cuda"
d, h, w = 16, 64, 64
x = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
g = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
wt = torch.randn(64, 64, 3, 3, 3, device=dev) * 0.03
bias = torch.zeros(64, device=dev)
packed = ops.pack_weights(wt, 64, 64, 27, False)
y = torch.empty_like(x)
z = torch.empty_like(x)
st = torch.cuda.current_stream().cuda_stream


def conv(src, dst):
    lib.call("hpvg_conv_forward", src.data_ptr(), 1, wt.data_ptr(), packed.data_ptr(), bias.data_ptr(), dst.data_ptr(), 1, 1, 64, 64, d, h, w, 3, 1, 0,
             1, 0.2, None, None, torch.cuda.current_stream().cuda_stream)


def lrelu(src, dst):
    lib.call("hpvg_lrelu_bwd", g.data_ptr(), src.data_ptr(), dst.data_ptr(), src.numel(), 0.2, 64, None, torch.cuda.current_stream().cuda_stream)


def wgrad():
    ops.wgrad_raw(x, g, 1, (64, 64, 3, 3, 3))


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
    return ts[10]


def seq_conv():
    for i in range(10):
        conv(x if i % 2 == 0 else y, y if i % 2 == 0 else x)


def seq_lrelu():
    for i in range(10):
        lrelu(x, z)


def seq_alt():
    for i in range(10):
        lrelu(x, z)
        conv(z, y)


def seq_wg():
    for i in range(5):
        wgrad()


def seq_conv_wg():
    for i in range(5):
        conv(x, y)
        wgrad()


a = timed("conv x10", seq_conv, 10)
c = timed("lrelu_bwd x10", seq_lrelu, 10)
b = timed("(lrelu_bwd, conv) x10", seq_alt, 10)
print("alternation overhead per pair: %.2f us" % ((b - a - c) / 10))
wg = timed("wgrad(+reduce) x5", seq_wg, 5)
cw = timed("(conv, wgrad) x5", seq_conv_wg, 5)
print("alternation overhead per (conv, wgrad) pair: %.2f us" % ((cw - a / 2 - wg) / 5))
