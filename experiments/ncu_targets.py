"""The dominant kernels at BASELINE configs[1]'s finest shape (64 -> 64 at 16 x 64 x 64), each launched a few times with the L2
flushed in between: the target of the `ncu --set full` capture (profiles/r02*_ncu_*.txt)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import lib, ops

dev = "cuda"
d, h, w = 16, 64, 64
x = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
g = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
wt = torch.randn(64, 64, 3, 3, 3, device=dev) * 0.03
bias = torch.randn(64, device=dev) * 0.1
gamma, beta = torch.ones(64, device=dev, requires_grad=True), torch.zeros(64, device=dev, requires_grad=True)
rm, rv, nbt = torch.zeros(64, device=dev), torch.ones(64, device=dev), torch.zeros((), dtype=torch.int64, device=dev)
flush = torch.empty(256 * 2**20 // 4, device=dev)
x3 = torch.randn(1, 3, d, h, w, device=dev)                      # the 3-channel clip: head convolution and the narrow weight gradients
wh = torch.randn(64, 3, 3, 3, 3, device=dev) * 0.1
stats3 = torch.zeros(128, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for rep in range(reps + 2):
    flush.zero_()
    y = ops.conv_raw(x, wt, bias, 1, False, True, act_slope=0.2)                 # conv_tc_kernel<1,4,3,1,64,true,false>: the critic's layer
    flush.zero_()
    xg = x.clone().requires_grad_(True)
    wg = wt.clone().requires_grad_(True)
    out = ops.conv_bn_lrelu(xg, wg, bias, gamma, beta, rm, rv, nbt, 1)            # conv_tc_kernel<1,4,3,1,64,true,true>: fused ConvBlock3D
    flush.zero_()
    out.backward(g)          # bn_lrelu_bwd_fused_kernel, conv_tc_kernel (data gradient), wgrad_tc_kdstack_kernel + wgrad_reduce_kernel
    flush.zero_()
    ops.conv_raw(x3, wh, bias, 1, False, True, stats=stats3)                      # expand_tc_kernel<3,3>: the 3 -> 64 head (narrow_tc.cu)
    flush.zero_()
    ops.wgrad_raw(x3, g, 1, (64, 3, 3, 3, 3), want_bias=True)                     # narrow_wgrad_tc_kernel<3,3,false>: head weight gradient
    flush.zero_()
    ops.wgrad_raw(x, x3, 1, (3, 64, 3, 3, 3), want_bias=True)                     # narrow_wgrad_tc_kernel<3,3,true>: tail weight gradient
    torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
