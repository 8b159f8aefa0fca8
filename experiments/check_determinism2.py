"""Development aid: bit-reproducibility of the pieces of a critic forward (spectral weights, head conv, tcgen05 convs, tail)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200"))
sys.path.insert(0, ROOT)
import torch
from hpvg import ops
from modules import networks_3d
from oracle import port


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


shape = (1, 3, 5, 18, 20)
opt = port.Opt(nfc=64, latent_dim=8, num_layer=3)
d = networks_3d.WDiscriminator3D(opt)
port.det_fill(d.state_dict(), 11)
d.cuda()
real = port.det_tensor(shape, 3).cuda()
u0 = {k: b.clone() for k, b in d.named_buffers()}
blocks = [d.head] + list(d.body)
first = None
with torch.no_grad():
    for rep in range(10):
        for k, b in d.named_buffers():
            b.copy_(u0[k])
        cur = {}
        ws = ops.spectral_weights([b.conv for b in blocks])
        for i, w in enumerate(ws):
            cur['w%d' % i] = w.detach().clone()
        for k, b in d.named_buffers():
            cur['buf.' + k] = b.clone()
        wfix = first and [first['w%d' % i] for i in range(len(ws))] or [w.detach().clone() for w in ws]
        h = real
        for i, b in enumerate(blocks):
            h = ops.conv(h, wfix[i], b.conv.bias, 1, True, 0.2)
            cur['h%d' % i] = h.clone()
        out = ops.conv(h, d.tail.weight, d.tail.bias, 1, False)
        cur['out'] = out.clone()
        torch.cuda.synchronize()
        if first is None:
            first = cur
            continue
        bad = ["%s:%.2e" % (k, rel(cur[k].float(), first[k].float())) for k in cur if not torch.equal(cur[k], first[k])]
        print("rep", rep, "differs:", bad if bad else "none")
