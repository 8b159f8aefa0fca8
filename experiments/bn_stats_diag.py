"""running statistics after one iteration: multi-stream (logged updates) vs single stream vs the CPU oracle"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "hp-vae-gan_b200"), ROOT, os.path.join(ROOT, "tests")]
import torch
from helpers import state_d_from, state_from, train_opt_from
from hpvg import train
from modules import networks_3d
from oracle import train_ref

fx = torch.load(os.path.join(ROOT, "tests", "golden", "train_gan_wide.pt"), map_location="cpu", weights_only=False)
real, real_zero = fx['real'].cuda(), fx['real_zero'].cuda()


def draws(dr):
    return [dr['noise_init']] + ([dr['eps_amp']] if 'eps_amp' in dr else []) + [dr['eps']] + [dr['noises'][l] for l in sorted(dr['noises'])]


def run(overlap, concurrent, iters=2):
    opt = train_opt_from(fx)
    g = networks_3d.GeneratorHPVAEGAN(opt)
    for _ in range(fx['stages']):
        g.init_next_stage()
    g.load_state_dict(state_from(fx), strict=True); g.cuda()
    d = networks_3d.WDiscriminator3D(opt); d.load_state_dict(state_d_from(fx), strict=True); d.cuda()
    tr = train.ScaleTrainer(opt, g, d, overlap=overlap)
    tr.concurrent_passes = concurrent
    feed = train.NoiseFeed(real.device)
    with feed:
        for it in range(iters):
            feed.load(draws(fx['draws'][it]), fx['draws'][it]['alpha'])
            tr.iteration(real, real_zero)
    torch.cuda.synchronize()
    return {k: v.detach().float().cpu() for k, v in g.state_dict().items() if 'running' in k or 'num_batches' in k}


oc = train_opt_from(fx)
sd_g, sd_d = state_from(fx), state_d_from(fx)
oracle = train_ref.ScaleTrainer(oc, sd_g, sd_d)
for it in range(2):
    dr = fx['draws'][it]
    oracle.iteration(fx['real'], fx['real_zero'], noise_init=dr['noise_init'], eps=dr['eps'], noises=dr['noises'], alpha=dr['alpha'], eps_amp=dr.get('eps_amp'))
ref = {k: v.detach().float() for k, v in sd_g.items() if 'running' in k or 'num_batches' in k}
for name, res in (("multi-stream logged", run(True, True)), ("overlap, serial passes", run(True, False)), ("single stream", run(False, False))):
    worst = max(((res[k] - ref[k]).abs().max().item() / (ref[k].abs().max().item() + 1e-9), k) for k in ref)
    bad = [(k, (res[k] - ref[k]).abs().max().item()) for k in ref if (res[k] - ref[k]).abs().max().item() > 2e-3 * ref[k].abs().max().item() + 1e-4]
    print("%-24s worst relative deviation from the oracle %.3e at %s; %d of %d buffers off: %s" % (name, worst[0], worst[1], len(bad), len(ref), bad[:4]))
