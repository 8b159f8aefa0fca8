"""Timing of ConvBlock3D forward as one launch (hpvg_conv_bn_lrelu_fused) against conv + bn_finalize_apply_lrelu (development aid).
   python experiments/check_fused_bn.py            # cold L2 (256 MB flush between launches), CUDA events
"""
import os
import sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import lib, ops

dev = "cuda"
flush = torch.empty(256 * 2**20 // 4, device=dev)


def timeit(fn, reps=20, cold=True):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        if cold:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


for vol in [(16, 64, 64), (6, 54, 54), (6, 46, 46), (4, 39, 39), (4, 32, 32), (13, 64, 64)]:
    d, h, w = vol
    x = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
    wt = torch.randn(64, 64, 3, 3, 3, device=dev) * 0.03
    bias = torch.randn(64, device=dev) * 0.1
    gamma, beta = torch.ones(64, device=dev), torch.zeros(64, device=dev)
    rm, rv, nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
    flops = 2.0 * d * h * w * 64 * 64 * 27
    res = {}
    for fused in (False, True):
        for need_bwd in (False, True):
            xx = x.clone().requires_grad_(need_bwd)

            def fn():
                with ops.fused_bn(fused), torch.set_grad_enabled(need_bwd), ops.zero_arena(dev, 1024):
                    return ops.conv_bn_lrelu(xx, wt, bias, gamma, beta, rm, rv, nbt, 1)
            for cold in (True, False):
                res[(fused, need_bwd, cold)] = timeit(fn, cold=cold)
    print("%-12s" % (vol,), " | ".join("%s %s %s: %6.1f us" % ("fused" if f else "2-launch", "train" if b else "infer", "cold" if c else "warm", res[(f, b, c)][0])
                                       for f in (False, True) for b in (False, True) for c in (True, False)),
          " (fused train cold: %.0f TFLOP/s)" % (flops / res[(True, True, True)][0] / 1e6), flush=True)
