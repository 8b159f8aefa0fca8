"""Development aid: repeat the critic forward / first-order backward / gradient-penalty backward on identical inputs and
report which results are not bit-reproducible (atomics order vs a race)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200"))
sys.path.insert(0, ROOT)
import torch
from hpvg import ops
from modules import networks_3d
from modules import utils as mutils
from oracle import port


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


shape = (1, 3, 5, 18, 20)
opt = port.Opt(nfc=64, latent_dim=8, num_layer=3)
d = networks_3d.WDiscriminator3D(opt)
port.det_fill(d.state_dict(), 11)
d.cuda()
real, fake = port.det_tensor(shape, 3).cuda(), port.det_tensor(shape, 4).cuda()
torch.rand = lambda *a, **k: torch.full((1, 1), 0.3)
u0 = {k: b.clone() for k, b in d.named_buffers()}
fuse = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
ops._FUSE_MASK[0] = fuse
first = None
for rep in range(8):
    with torch.no_grad():
        for k, b in d.named_buffers():
            b.copy_(u0[k])
    cur = {}
    d.zero_grad()
    x = real.clone().requires_grad_(True)
    out = d(x)
    cur['out'] = out.detach().clone()
    (-out.mean()).backward()
    cur['gx1'] = x.grad.clone()
    for k, p in d.named_parameters():
        cur['g1.' + k] = p.grad.clone()
    with torch.no_grad():
        for k, b in d.named_buffers():
            b.copy_(u0[k])
    d.zero_grad()
    gp = mutils.calc_gradient_penalty(d, real, fake, 0.1, 'cuda')
    cur['gp'] = gp.detach().clone()
    gp.backward()
    for k, p in d.named_parameters():
        if p.grad is not None:
            cur['g2.' + k] = p.grad.clone()
    torch.cuda.synchronize()
    if first is None:
        first = cur
        continue
    bad = ["%s:%.2e" % (k, rel(cur[k], first[k])) for k in cur if not torch.equal(cur[k], first[k])]
    print("fuse", fuse, "rep", rep, "differs:", bad if bad else "none")
