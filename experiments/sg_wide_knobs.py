"""rec_loss trajectory of the train_sg_wide / train_gan_wide fixtures under the development knobs (diagnosis aid)"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "hp-vae-gan_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from helpers import state_d_from, state_from, train_opt_from
from hpvg import images, train, ops
from modules import networks_3d

name = sys.argv[1] if len(sys.argv) > 1 else "train_sg_wide"
fx = torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), map_location="cpu", weights_only=False)
opt = train_opt_from(fx)
baseline = 'generator' in fx
g = getattr(networks_3d, fx['generator'] if baseline else 'GeneratorHPVAEGAN')(opt)
for _ in range(fx['stages']):
    g.init_next_stage()
g.load_state_dict(state_from(fx), strict=True)
g.cuda()
d = networks_3d.WDiscriminator3D(opt)
d.load_state_dict(state_d_from(fx), strict=True)
d.cuda()
tr = train.BaselineTrainer(opt, g, d) if baseline else train.ScaleTrainer(opt, g, d)
queue, alphas = [], []
images.draw_normal = lambda shape, dtype, device: queue.pop(0).to(device=device, dtype=dtype)
torch.rand = lambda *a, **k: torch.full((1, 1), alphas.pop(0))
real = fx['real'].cuda()
second = (fx['z_init'] if baseline else fx['real_zero']).cuda()
for it in range(fx['iters']):
    dr = fx['draws'][it]
    if baseline:
        queue[:] = [dr['noise_init']] + [dr['noises'][l] for l in sorted(dr['noises'])]
    else:
        queue[:] = [dr['noise_init']] + ([dr['eps_amp']] if 'eps_amp' in dr else []) + [dr['eps']] + [dr['noises'][l] for l in sorted(dr['noises'])]
    alphas.append(dr['alpha'])
    out = tr.iteration(real, second)
    ref = fx['losses'][it]
    print(it, " ".join("%s %.5f/%.5f (%+.2f%%)" % (k, out[k].item(), ref[k], 100 * (out[k].item() - ref[k]) / abs(ref[k])) for k in ('rec_loss', 'gradient_penalty', 'errD_real', 'errD_fake')))
