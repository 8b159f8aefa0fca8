import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from helpers import opt_from, state_from, rel_err
from hpvg import images
from modules import networks_3d, networks_2d
from modules.losses import kl_criterion
name = sys.argv[1] if len(sys.argv) > 1 else "hp3d_tiny"
fx = torch.load(os.path.join(ROOT, "tests/golden", name + ".pt"), weights_only=False)
opt = opt_from(fx)
nets = networks_2d if name.startswith("hp2d") else networks_3d
g = nets.GeneratorHPVAEGAN(opt)
for _ in range(fx['stages']): g.init_next_stage()
g.load_state_dict(state_from(fx)); g.cuda()
rec = fx['rec']
q = [rec['eps']]
images.draw_normal = lambda shape, dtype, device: q.pop(0).to(device=device, dtype=dtype)
gen, gen_vae, (mu, logvar) = g(fx['real_zero'].cuda(), fx['amps'], mode='rec')
print("gen", rel_err(gen, rec['gen']), "vae", rel_err(gen_vae, rec['gen_vae']), "mu", rel_err(mu, rec['mu']))
loss = 10.0 * (F.mse_loss(gen, fx['real'].cuda()) + F.mse_loss(gen_vae, fx['real_zero'].cuda())) + kl_criterion(mu, logvar)
print("loss", loss.item(), rec['loss'])
loss.backward()
for k, p in g.named_parameters():
    gg = rec['grads'].get(k)
    if gg is None or p.grad is None:
        print(k, "none", p.grad is None, gg is None); continue
    if isinstance(gg, dict):
        print("%-50s norm %.4e ref %.4e headrel %.3e" % (k, p.grad.norm().item(), gg['norm'], rel_err(p.grad.flatten()[:64], gg['head'])))
    else:
        print("%-50s rel %.3e  norm %.3e" % (k, rel_err(p.grad, gg), gg.norm().item()))
