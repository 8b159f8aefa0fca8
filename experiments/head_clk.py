"""per-CTA time stamps of expand_tc_kernel (3 -> 64 head at 16 x 64 x 64): where does a CTA's life go?"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import ops, lib
dev = "cuda"
d, h, w = 16, 64, 64
g3 = torch.randn(1, 3, d, h, w, device=dev)
wh = torch.randn(64, 3, 3, 3, 3, device=dev) * 0.1
b64 = torch.zeros(64, device=dev)
for _ in range(3):
    ops.conv_raw(g3, wh, b64, 1, False, True)
dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
lib.call("hpvg_debug_set_clock_buffer", dbg.data_ptr())
names = ["start", "halo0 in smem", "A0 built", "A1 built", "A2 built", "A3 built", "-", "mma0 issued", "mma1", "mma2", "mma3",
         "store0 issued", "store1", "store2", "store3", "end"]
for rep in range(3):
    dbg.zero_()
    ops.conv_raw(g3, wh, b64, 1, False, True)
    torch.cuda.synchronize()
    t = dbg.view(148, 16).cpu()
    t0 = t[:, 0].min()
    print("rep", rep, "kernel span (first start -> last end): %d ns; CTA start spread %d ns" % (int(t[:, 15].max() - t0), int(t[:, 0].max() - t0)))
    for b in (0, 50, 100, 147):
        row = t[b]
        print("  cta %3d: " % b + ", ".join("%s %d" % (names[i], int(row[i] - row[0])) for i in range(1, 16) if row[i] > 0))
lib.call("hpvg_debug_set_clock_buffer", None)
