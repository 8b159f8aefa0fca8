"""Development aid: the critic's first-order + gradient-penalty gradients with the LeakyReLU derivative fused into the
data-gradient epilogue (ops._FUSE_MASK) and with the separate leaky_relu_backward launches, both against the fp32 CPU
oracle — tells bf16 rounding noise (both equally far from fp32) from a kernel bug (one of them further)."""
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


for nfc, shape in ((64, (1, 3, 5, 18, 20)), (64, (1, 3, 8, 32, 32))):
    opt = port.Opt(nfc=nfc, latent_dim=8, num_layer=3)
    d = networks_3d.WDiscriminator3D(opt)
    port.det_fill(d.state_dict(), 11)
    sd = {k: v.detach().clone() for k, v in d.state_dict().items()}
    d.cuda()
    real, fake = port.det_tensor(shape, 3), port.det_tensor(shape, 4)
    saved_rand = torch.rand
    torch.rand = lambda *a, **k: torch.full((1, 1), 0.3)
    # fp32 oracle
    for k, v in sd.items():
        if not k.endswith(('weight_u', 'weight_v')):
            v.requires_grad_(True)
    xo = real.clone().requires_grad_(True)
    lo = -port.discriminator(sd, opt, xo).mean() + port.discriminator(sd, opt, fake).mean() + port.gradient_penalty(sd, opt, real, fake, 0.1, alpha=0.3)
    lo.backward()
    u0 = {k: b.clone() for k, b in d.named_buffers()}
    res = {}
    for fuse in (False, True):
        with torch.no_grad():
            for k, b in d.named_buffers():
                b.copy_(u0[k])
        ops._FUSE_MASK[0] = fuse
        d.zero_grad()
        x = real.cuda().requires_grad_(True)
        loss = -d(x).mean() + d(fake.cuda()).mean() + mutils.calc_gradient_penalty(d, real.cuda(), fake.cuda(), 0.1, 'cuda')
        loss.backward()
        torch.cuda.synchronize()
        res[fuse] = (x.grad.clone(), {k: p.grad.clone() for k, p in d.named_parameters()}, loss.item())
    torch.rand = saved_rand
    print("shape", shape, "loss oracle %.6f plain %.6f fused %.6f" % (lo.item(), res[False][2], res[True][2]))
    print("  gx: plain-vs-oracle %.4f fused-vs-oracle %.4f fused-vs-plain %.4f" % (rel(res[False][0], xo.grad), rel(res[True][0], xo.grad), rel(res[True][0], res[False][0])))
    for k in res[False][1]:
        print("  %-28s plain-vs-oracle %.4f fused-vs-oracle %.4f fused-vs-plain %.4f" % (k, rel(res[False][1][k], sd[k].grad), rel(res[True][1][k], sd[k].grad), rel(res[True][1][k], res[False][1][k])))
