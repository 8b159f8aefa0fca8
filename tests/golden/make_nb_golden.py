"""Golden vectors for the Bernoulli-gated variants: GeneratorVAE_nb / Encode3DVAE_nb / Encode2DVAE_nb
(reference modules/networks_3d.py:110-138, :409-485, networks_2d.py:115-143, :272-348), recorded from the UNMODIFIED reference
modules on deterministic weights (oracle.port.det_fill) and inputs.  The random draws inside forward are avoided: the encoder is
run on its own (it draws nothing) and the generator is driven with noise_init_norm / noise_init_bern in 'rec' mode.
Run in the build container:  python tests/golden/make_nb_golden.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, "/root/reference"]

from modules import networks_2d, networks_3d  # noqa: E402  (the reference's)
from oracle import port  # noqa: E402


def case(dims):
    nets = networks_3d if dims == 3 else networks_2d
    opt = port.Opt(nfc=8, latent_dim=8, num_layer=2, vae_levels=1, img_size=20, min_size=12, sampling_rates=[4, 2, 1])
    g = nets.GeneratorVAE_nb(opt)
    g.init_next_stage()
    g.init_next_stage()
    port.det_fill(g.state_dict(), 21 + dims)
    state = [(k, tuple(v.shape)) for k, v in g.state_dict().items()]
    s0, t0 = port.scale_size(0, opt), port.time_depth(0, opt)
    sp = (t0, s0, s0) if dims == 3 else (s0, s0)
    video = port.det_tensor((2, 3) + sp, 41)
    mu, logvar, bern = g.encode(video)
    z_norm = port.det_tensor((2, 8) + (1,) * dims, 42)
    z_bern = port.det_tensor((2, 1) + sp, 43).abs()
    out, vae_out = g(None, [1.0, 0.1, 0.1], noise_init_norm=z_norm, noise_init_bern=z_bern, mode='rec')
    gout = port.det_tensor(tuple(out.shape), 44)
    loss = (out * gout).sum() + (mu * port.det_tensor(tuple(mu.shape), 45)).sum() + (logvar * port.det_tensor(tuple(logvar.shape), 46)).sum() \
        + (bern * port.det_tensor(tuple(bern.shape), 47)).sum()
    g.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in g.named_parameters() if p.grad is not None}
    return dict(opt=dict(opt.__dict__), fill_seed=21 + dims, state=state, video=video, z_norm=z_norm, z_bern=z_bern, gout=gout,
                out=out.detach(), vae_out=vae_out.detach(), mu=mu.detach(), logvar=logvar.detach(), bern=bern.detach(), grads=grads,
                amps=[1.0, 0.1, 0.1], seeds=dict(mu=45, logvar=46, bern=47))


def dbase_gp_case(name, opt, shape):
    """WDiscriminatorBaselines (modules/networks_3d.py:184-210) through the WGAN-GP double backward (modules/utils.py:4-19):
    the critic loss of train_video_baselines.py:131-149 on fixed inputs and a fixed alpha"""
    from modules.utils import calc_gradient_penalty
    d = networks_3d.WDiscriminatorBaselines(opt)
    port.det_fill(d.state_dict(), seed=9)
    real = port.det_tensor(shape, 33)
    fake = port.det_tensor(shape, 34, scale=0.8)
    alpha = 0.37
    stock = torch.rand
    torch.rand = lambda *a, **k: torch.full((1, 1), alpha)
    try:
        d.zero_grad()
        err_real = -d(real).mean()
        err_fake = d(fake).mean()
        gp = calc_gradient_penalty(d, real, fake, 0.1, 'cpu')
        (err_real + err_fake + gp).backward()
    finally:
        torch.rand = stock
    fx = {'opt': dict(opt.__dict__), 'fill_seed': 9, 'real': real, 'fake': fake, 'alpha': alpha, 'lambda': 0.1, 'gp': gp.item(),
          'state': [(k, tuple(v.shape)) for k, v in d.state_dict().items()],
          'grads': {k: p.grad.detach().clone() for k, p in d.named_parameters()}}
    path = os.path.join(HERE, name + '.pt')
    torch.save(fx, path)
    print("wrote", path, os.path.getsize(path), "bytes; gp", fx['gp'])


if __name__ == "__main__":
    dbase_gp_case('dbase3d_gp_tiny', port.Opt(nfc=8, latent_dim=8, num_layer=2, vae_levels=2, img_size=20, min_size=12, sampling_rates=[4, 2, 1]),
                  (1, 3, 4, 12, 11))
    for dims in (3, 2):
        fx = case(dims)
        path = os.path.join(HERE, "nb%dd_tiny.pt" % dims)
        torch.save(fx, path)
        print("wrote", path, os.path.getsize(path), "bytes;", len(fx['state']), "state tensors,", len(fx['grads']), "gradients")
