"""BASELINE configs[2] (GeneratorSG, SinGAN-3D baseline) at full size: forward ('rec' and 'rand') + backward sanity and timing"""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200")); sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
from hpvg.options import Options
from modules import networks_3d
from modules.utils import calc_gradient_penalty
dev = torch.device("cuda", 0)
o = Options(img_size=64, sampling_rates=[5, 3, 1], nfc=64, num_layer=5, batch_size=1)
torch.manual_seed(0)
G = networks_3d.GeneratorSG(o)
for _ in range(o.stop_scale): G.init_next_stage()
D = networks_3d.WDiscriminator3D(o)
G.to(dev); D.to(dev)
t0_, h0, w0 = o.level_size(0)
amps = [1.0] + [0.1] * o.stop_scale
z = torch.randn(1, 3, t0_, h0, w0, device=dev)
T, H, W = o.level_size(o.stop_scale)
real = torch.rand(1, 3, T, H, W, device=dev) * 2 - 1
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    fake = G(z, amps, mode='rand')
    rec = G(z, amps, mode='rec')
    errD = -D(real).mean() + D(fake.detach()).mean() + calc_gradient_penalty(D, real, fake, 0.1, dev)
    D.zero_grad(); errD.backward()
    loss = -D(fake).mean() + 10 * F.mse_loss(rec, real)
    G.zero_grad(); loss.backward()
    torch.cuda.synchronize()
    print("iteration %d: %.1f ms  fake %s  errD %.4f  loss %.4f  finite grads %s" % (
        it, (time.perf_counter() - t0) * 1e3, tuple(fake.shape), errD.item(), loss.item(),
        all(torch.isfinite(p.grad).all().item() for p in G.parameters() if p.grad is not None)))
