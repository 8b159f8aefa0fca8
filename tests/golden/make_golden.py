"""Generate the golden fixtures in this directory from the UNMODIFIED reference (run in the build container only):

    python tests/golden/make_golden.py            # needs /root/reference, writes tests/golden/*.pt
    python tests/golden/make_golden.py train      # the training-loop fixtures
    python tests/golden/make_golden.py baselines  # GeneratorCSG / WDiscriminatorBaselines fixtures
    python tests/golden/make_golden.py train_baselines   # train_video_baselines.py loop fixtures (BASELINE configs[2])

Each fixture holds deterministic closed-form weights' *recipe* (oracle.port.det_fill seed), the inputs, and what the
reference's own modules (modules/networks_3d.py, networks_2d.py, losses.py, utils.py) computed from them on CPU in
float64-free plain fp32: outputs, losses, parameter gradients (full for the tiny nets, per-tensor summaries for the
64-channel nets) and the buffers the forward pass mutates.  The reference itself cannot travel to the GPU box.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, '/root/reference')
sys.path.insert(1, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from modules import networks_2d, networks_3d  # noqa: E402  (the reference's)
from modules.losses import kl_criterion  # noqa: E402
from modules.utils import calc_gradient_penalty  # noqa: E402

from oracle import port  # noqa: E402

assert networks_3d.__file__.startswith('/root/reference'), networks_3d.__file__
torch.set_num_threads(8)


def summarize(t):
    t = t.detach().flatten().double()
    return {'norm': t.norm().item(), 'sum': t.sum().item(), 'head': t[:64].float().clone()}


def grads_of(module, full):
    out = {}
    for k, p in module.named_parameters():
        if p.grad is None:
            continue
        out[k] = p.grad.detach().clone() if full else summarize(p.grad)
    return out


def buffers_of(module):
    return {k: b.detach().clone() for k, b in module.named_buffers()
            if k.endswith(('running_mean', 'running_var', 'num_batches_tracked', 'weight_u', 'weight_v'))}


def tiny_opt(**kw):
    base = dict(nfc=8, latent_dim=8, num_layer=2, enc_blocks=2, vae_levels=2, img_size=20, min_size=12,
                sampling_rates=[4, 2, 1])
    base.update(kw)
    return port.Opt(**base)


def make_generator(nets, opt, stages, seed):
    g = nets.GeneratorHPVAEGAN(opt)
    for _ in range(stages):
        g.init_next_stage()
    port.det_fill(g.state_dict(), seed)
    return g


def gen_case(name, nets, opt, stages, three_d, full_grads, t0=None):
    """rec pass with loss + grads, then a rand pass, on the same (mutated) module state"""
    g = make_generator(nets, opt, stages, seed=3)
    s0 = port.scale_size(0, opt)
    sN = port.scale_size(stages, opt)
    if three_d:
        t0 = t0 or port.time_depth(0, opt)
        tN = port.time_depth(stages, opt) if stages > 0 else t0
        real_zero = port.det_tensor((1, 3, t0, s0, s0), 11)
        real = port.det_tensor((1, 3, tN, sN, sN), 12)
        z_shape = (1, opt.latent_dim, t0, s0, s0)
    else:
        real_zero = port.det_tensor((1, 3, s0, s0), 11)
        real = port.det_tensor((1, 3, sN, sN), 12)
        z_shape = (1, opt.latent_dim, s0, s0)
    amps = [1.0] + [0.1 + 0.01 * i for i in range(stages)]
    fx = {'opt': dict(opt.__dict__), 'stages': stages, 'fill_seed': 3, 'real_zero': real_zero, 'real': real, 'amps': amps,
          'state': [(k, tuple(v.shape)) for k, v in g.state_dict().items()],
          'sizes': [(port.scale_size(i, opt), port.time_depth(i, opt)) for i in range(stages + 1)]}

    torch.manual_seed(5)
    eps = torch.zeros(z_shape).normal_()
    torch.manual_seed(5)
    gen, gen_vae, (mu, logvar) = g(real_zero, amps, mode='rec')
    loss = 10.0 * (F.mse_loss(gen, real) + F.mse_loss(gen_vae, real_zero)) + kl_criterion(mu, logvar)
    g.zero_grad()
    loss.backward()
    fx['rec'] = {'eps': eps, 'gen': gen.detach().clone(), 'gen_vae': gen_vae.detach().clone(), 'mu': mu.detach().clone(),
                 'logvar': logvar.detach().clone(), 'loss': loss.item(), 'kl': kl_criterion(mu, logvar).item(),
                 'grads': grads_of(g, full_grads), 'buffers': buffers_of(g)}

    # rand pass: noises are drawn inside forward in level order; reproduce them from the same seed
    z = port.det_tensor(z_shape, 21)
    torch.manual_seed(9)
    noises = {}
    x_shape = None
    with torch.no_grad():
        probe_sd = {k: v.clone() for k, v in g.state_dict().items()}
        cur = port.generator(probe_sd, opt, None, amps, noise_init=z, mode='rec')[0]   # shapes only
    # shapes of x_up per level
    import utils as ref_utils
    x = torch.zeros(1, 3, *( [t0, s0, s0] if three_d else [s0, s0]))
    for idx in range(stages):
        x = ref_utils.upscale(x, idx + 1, opt) if three_d else ref_utils.upscale_2d(x, idx + 1, opt)
        if (not three_d) or opt.vae_levels <= idx + 1:
            noises[idx + 1] = torch.zeros_like(x).normal_(0, 1)
    torch.manual_seed(9)
    with torch.no_grad():
        fake, fake_vae = g(z, amps, noise_init=z, mode='rand')
    fx['rand'] = {'z': z, 'noises': noises, 'fake': fake.clone(), 'fake_vae': fake_vae.clone(), 'buffers': buffers_of(g)}
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'loss', fx['rec']['loss'], 'gen', tuple(gen.shape), 'fake', tuple(fake.shape))


def disc_case(name, nets, opt, three_d, shape, full_grads):
    d = (nets.WDiscriminator3D if three_d else nets.WDiscriminator2D)(opt)
    port.det_fill(d.state_dict(), seed=7)
    real = port.det_tensor(shape, 31)
    fake = port.det_tensor(shape, 32, scale=0.8)
    fx = {'opt': dict(opt.__dict__), 'fill_seed': 7, 'real': real, 'fake': fake, 'lambda': 0.1,
          'state': [(k, tuple(v.shape)) for k, v in d.state_dict().items()]}
    d.zero_grad()
    out_real = d(real)
    err_real = -out_real.mean()
    out_fake = d(fake)
    err_fake = out_fake.mean()
    torch.manual_seed(13)
    alpha = float(torch.rand(1, 1))
    torch.manual_seed(13)
    gp = calc_gradient_penalty(d, real, fake, 0.1, 'cpu')
    (err_real + err_fake + gp).backward()
    fx.update({'alpha': alpha, 'out_real': out_real.detach().clone(), 'out_fake': out_fake.detach().clone(), 'gp': gp.item(),
               'grads': grads_of(d, full_grads), 'buffers': buffers_of(d)})
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'gp', fx['gp'], 'out', tuple(out_real.shape))


def sg_case(name, opt, stages):
    g = networks_3d.GeneratorSG(opt)
    for _ in range(stages):
        g.init_next_stage()
    port.det_fill(g.state_dict(), seed=4)
    s0, t0 = port.scale_size(0, opt), port.time_depth(0, opt)
    z = port.det_tensor((1, 3, t0, s0, s0), 41)
    amps = [1.0] + [0.1] * stages
    fx = {'opt': dict(opt.__dict__), 'stages': stages, 'fill_seed': 4, 'z': z, 'amps': amps,
          'state': [(k, tuple(v.shape)) for k, v in g.state_dict().items()]}
    target = None
    out = g(z, amps, mode='rec')
    target = port.det_tensor(tuple(out.shape), 42)
    loss = F.mse_loss(out, target)
    g.zero_grad()
    loss.backward()
    fx['rec'] = {'out': out.detach().clone(), 'target': target, 'loss': loss.item(), 'grads': grads_of(g, True),
                 'buffers': buffers_of(g)}
    # rand: reproduce noise shapes
    import utils as ref_utils
    m = opt.num_layer + 2
    torch.manual_seed(17)
    noises = {}
    x = torch.zeros(1, 3, t0, s0, s0)
    for idx in range(1, stages + 1):
        x = ref_utils.upscale(x, idx, opt)
        noises[idx] = torch.zeros(1, 3, *[s + 2 * m for s in x.shape[-3:]]).normal_(0, 1)
    torch.manual_seed(17)
    with torch.no_grad():
        fake = g(z, amps, mode='rand')
    fx['rand'] = {'noises': noises, 'fake': fake.clone()}
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'loss', fx['rec']['loss'], 'out', tuple(out.shape))


def csg_case(name, opt, stages, full_grads):
    """GeneratorCSG (the default generator of train_video_baselines.py): reconstruction pass + gradients, and a 'rand' pass
    with the noise the reference draws under a fixed seed"""
    g = networks_3d.GeneratorCSG(opt)
    for _ in range(stages):
        g.init_next_stage()
    port.det_fill(g.state_dict(), seed=5)
    s0, t0 = port.scale_size(0, opt), port.time_depth(0, opt)
    z = port.det_tensor((1, 3, t0, s0, s0), 43)
    amps = [1.0] + [0.1] * stages
    fx = {'opt': dict(opt.__dict__), 'stages': stages, 'fill_seed': 5, 'z': z, 'amps': amps,
          'state': [(k, tuple(v.shape)) for k, v in g.state_dict().items()]}
    out = g(z, amps, mode='rec')
    target = port.det_tensor(tuple(out.shape), 44)
    loss = F.mse_loss(out, target)
    g.zero_grad()
    loss.backward()
    fx['rec'] = {'out': out.detach().clone(), 'target': target, 'loss': loss.item(), 'grads': grads_of(g, full_grads),
                 'buffers': buffers_of(g)}
    import utils as ref_utils
    m = opt.num_layer
    torch.manual_seed(19)
    noises = {}
    x = torch.zeros(1, 3, t0, s0, s0)
    for idx in range(1, stages + 1):
        x = ref_utils.upscale(x, idx, opt)
        noises[idx] = torch.zeros(1, opt.nfc, *[s + 2 * m for s in x.shape[-3:]]).normal_(0, 1)
    torch.manual_seed(19)
    with torch.no_grad():
        fake = g(z, amps, mode='rand')
    fx['rand'] = {'noises': noises, 'fake': fake.clone()}
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'loss', fx['rec']['loss'], 'out', tuple(out.shape))


def dbase_case(name, opt, shape, full_grads):
    """WDiscriminatorBaselines: critic outputs and the first-order gradients of -D(real).mean() + D(fake).mean()"""
    d = networks_3d.WDiscriminatorBaselines(opt)
    port.det_fill(d.state_dict(), seed=9)
    real = port.det_tensor(shape, 33)
    fake = port.det_tensor(shape, 34, scale=0.8)
    fx = {'opt': dict(opt.__dict__), 'fill_seed': 9, 'real': real, 'fake': fake,
          'state': [(k, tuple(v.shape)) for k, v in d.state_dict().items()]}
    d.zero_grad()
    out_real = d(real)
    out_fake = d(fake)
    (-out_real.mean() + out_fake.mean()).backward()
    fx.update({'out_real': out_real.detach().clone(), 'out_fake': out_fake.detach().clone(), 'grads': grads_of(d, full_grads),
               'buffers': buffers_of(d)})
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'out', tuple(out_real.shape))


class DrawQueue(object):
    """feeds pre-generated draws to the reference's own normal_() / torch.rand() call sites, in call order, and records
    nothing else: the reference modules stay unmodified, only the source of randomness is replaced"""

    def __init__(self):
        self.normals, self.rands = [], []
        self._normal_, self._rand = torch.Tensor.normal_, torch.rand

    def __enter__(self):
        q = self

        def normal_(t, *a, **k):
            src = q.normals.pop(0)
            assert tuple(src.shape) == tuple(t.shape), (tuple(src.shape), tuple(t.shape))
            return t.copy_(src)

        def rand(*a, **k):
            return torch.full((1, 1), q.rands.pop(0))
        torch.Tensor.normal_, torch.rand = normal_, rand
        return self

    def __exit__(self, *a):
        torch.Tensor.normal_, torch.rand = self._normal_, self._rand
        assert not self.normals and not self.rands, "unconsumed draws: the loop does not draw in the assumed order"


def train_case(name, opt, scale_idx, iters, seed=0, three_d=True):
    """K iterations of the reference's per-scale loop (train_video.py:111-202, restated here only as the DRIVER: every
    network, loss and penalty call goes to the reference's own modules) at pyramid level `scale_idx`."""
    import torch.optim as optim
    from oracle import train_ref
    for k, v in train_ref.TRAIN_DEFAULTS.items():
        setattr(opt, k, v)
    opt.scale_idx = scale_idx
    opt.device = 'cpu'
    nets = networks_3d if three_d else networks_2d
    g = make_generator(nets, opt, scale_idx, seed=3)
    gan = opt.vae_levels < scale_idx + 1
    s0, t0 = port.scale_size(0, opt), port.time_depth(0, opt)
    sN, tN = port.scale_size(scale_idx, opt), port.time_depth(scale_idx, opt)
    if three_d:
        real_zero = port.det_tensor((1, 3, t0, s0, s0), 51)
        real = port.det_tensor((1, 3, tN, sN, sN), 52) if scale_idx > 0 else real_zero
        opt.Z_init_size = [1, opt.latent_dim, t0, s0, s0]
    else:       # train_image.py: 4-D tensors, Z_init_size without the time axis (train_image.py:106-108)
        real_zero = port.det_tensor((1, 3, s0, s0), 51)
        real = port.det_tensor((1, 3, sN, sN), 52) if scale_idx > 0 else real_zero
        opt.Z_init_size = [1, opt.latent_dim, s0, s0]
    opt.Noise_Amps = [1.0] + [0.1] * (scale_idx - 1) if scale_idx > 0 else []
    amps_before = list(opt.Noise_Amps)
    fx = {'opt': {k: v for k, v in opt.__dict__.items() if k != 'Noise_Amps'}, 'stages': scale_idx, 'fill_seed': 3, 'fill_seed_d': 7,
          'real': real, 'real_zero': real_zero, 'amps_before': amps_before, 'iters': iters,
          'state': [(k, tuple(v.shape)) for k, v in g.state_dict().items()]}
    d = None
    if gan:
        d = networks_3d.WDiscriminator3D(opt) if three_d else networks_2d.WDiscriminator2D(opt)
        port.det_fill(d.state_dict(), seed=7)
        fx['state_d'] = [(k, tuple(v.shape)) for k, v in d.state_dict().items()]
        optimizerD = optim.Adam(d.parameters(), lr=opt.lr_d, betas=(opt.beta1, 0.999))
    fx['three_d'] = three_d
    # parameter groups exactly as train_video.py:57-88 builds them for the default flags (train_all False)
    if gan:
        depth = min(opt.train_depth, len(g.body) - opt.vae_levels + 1)
        groups = [{"params": b.parameters(), "lr": opt.lr_g * (opt.lr_scale ** (len(g.body[-depth:]) - 1 - i))}
                  for i, b in enumerate(g.body[-depth:])]
    else:
        lr = opt.lr_g * (opt.lr_scale ** scale_idx)
        groups = [{"params": g.encode.parameters(), "lr": lr}, {"params": g.decoder.parameters(), "lr": lr}]
        groups += [{"params": b.parameters(), "lr": opt.lr_g * (opt.lr_scale ** (len(g.body[-opt.train_depth:]) - 1 - i))}
                   for i, b in enumerate(g.body[-opt.train_depth:])]
    optimizerG = optim.Adam(groups, lr=opt.lr_g, betas=(opt.beta1, 0.999))

    # shapes of the per-level noise of a 'rand' pass
    import utils as ref_utils
    noise_shapes = {}
    x = torch.zeros(1, 3, t0, s0, s0) if three_d else torch.zeros(1, 3, s0, s0)
    for idx in range(scale_idx):
        x = ref_utils.upscale(x, idx + 1, opt) if three_d else ref_utils.upscale_2d(x, idx + 1, opt)
        if opt.vae_levels <= idx + 1 or not three_d:        # the 2-D generator adds noise at every level (networks_2d.py:261-263)
            noise_shapes[idx + 1] = tuple(x.shape)
    gen = torch.Generator().manual_seed(1000 + seed)

    def randn(shape):
        return torch.randn(shape, generator=gen)

    draws, losses = [], []
    for it in range(iters):
        dr = {'noise_init': randn(tuple(opt.Z_init_size))}
        if it == 0 and scale_idx > 0:
            dr['eps_amp'] = randn(tuple(opt.Z_init_size))
        dr['eps'] = randn(tuple(opt.Z_init_size))
        if gan:
            dr['noises'] = {lvl: randn(shape) for lvl, shape in sorted(noise_shapes.items())}
            dr['alpha'] = float(torch.rand(1, generator=gen))
        draws.append(dr)
        with DrawQueue() as q:
            q.normals = [dr['noise_init']] + ([dr['eps_amp']] if 'eps_amp' in dr else []) + [dr['eps']]
            if gan:
                q.normals += [dr['noises'][lvl] for lvl in sorted(dr['noises'])]
                q.rands = [dr['alpha']]
            noise_init = ref_utils.generate_noise(size=opt.Z_init_size, device='cpu')
            if it == 0:
                if scale_idx == 0:
                    opt.noise_amp = 1
                    opt.Noise_Amps.append(1)
                else:
                    with torch.no_grad():
                        opt.Noise_Amps.append(0)
                        z_rec, _, _ = g(real_zero, opt.Noise_Amps, mode="rec")
                        opt.noise_amp = opt.noise_amp_init * torch.sqrt(F.mse_loss(real, z_rec)).item() / opt.batch_size
                        opt.Noise_Amps[-1] = opt.noise_amp
            rec = {}
            generated, generated_vae, (mu, logvar) = g(real_zero, opt.Noise_Amps, mode="rec")
            if not gan:
                rec_vae_loss = F.mse_loss(generated, real) + F.mse_loss(generated_vae, real_zero)
                kl_loss = kl_criterion(mu, logvar)
                total_loss = opt.rec_weight * rec_vae_loss + opt.kl_weight * kl_loss
                rec.update(rec_vae_loss=rec_vae_loss.item(), kl_loss=kl_loss.item())
            else:
                d.zero_grad()
                errD_real = -d(real).mean()
                fake, _ = g(noise_init, opt.Noise_Amps, noise_init=noise_init, mode="rand")
                errD_fake = d(fake.detach()).mean()
                gp = calc_gradient_penalty(d, real, fake, opt.lambda_grad, 'cpu')
                (errD_real + errD_fake + gp).backward()
                optimizerD.step()
                rec_loss = F.mse_loss(generated, real)
                errG = -d(fake).mean() * opt.disc_loss_weight
                total_loss = opt.rec_weight * rec_loss + errG
                rec.update(rec_loss=rec_loss.item(), errG=errG.item(), errD_real=errD_real.item(), errD_fake=errD_fake.item(),
                           gradient_penalty=gp.item())
            g.zero_grad()
            total_loss.backward()
            torch.nn.utils.clip_grad_norm_(g.parameters(), opt.grad_clip)
            optimizerG.step()
            rec['total_loss'] = total_loss.item()
        losses.append(rec)
    fx.update({'draws': draws, 'losses': losses, 'noise_amps_after': list(opt.Noise_Amps),
               'final_tail_weight': g.state_dict()[('body.%d.tail.weight' % (scale_idx - 1)) if scale_idx > 0 else 'decoder.tail.weight'].clone()})
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'losses[0]', losses[0], 'losses[-1]', losses[-1], 'amps', opt.Noise_Amps)


def train_baselines_case(name, gen_name, opt, scale_idx, iters, seed=0):
    """K iterations of train_video_baselines.py:100-173 (restated here only as the DRIVER: every network and penalty call goes
    to the reference's own modules) at pyramid level `scale_idx`, default critic WDiscriminator3D, Dsteps = Gsteps = 1."""
    import torch.optim as optim
    import utils as ref_utils
    from oracle import train_ref
    for k, v in train_ref.BASELINE_DEFAULTS.items():
        setattr(opt, k, v)
    opt.scale_idx = scale_idx
    opt.device = 'cpu'
    g = getattr(networks_3d, gen_name)(opt)
    for _ in range(scale_idx):
        g.init_next_stage()
    port.det_fill(g.state_dict(), seed=6)
    d = networks_3d.WDiscriminator3D(opt)
    port.det_fill(d.state_dict(), seed=8)
    s0, t0 = port.scale_size(0, opt), port.time_depth(0, opt)
    sN, tN = port.scale_size(scale_idx, opt), port.time_depth(scale_idx, opt)
    z_init = port.det_tensor((1, 3, t0, s0, s0), 53)
    real = port.det_tensor((1, 3, tN, sN, sN), 54)
    opt.Noise_Amps = [1.0] + [0.1] * (scale_idx - 1) if scale_idx > 0 else []
    fx = {'opt': {k: v for k, v in opt.__dict__.items() if k != 'Noise_Amps'}, 'generator': gen_name, 'stages': scale_idx,
          'fill_seed': 6, 'fill_seed_d': 8, 'real': real, 'z_init': z_init, 'amps_before': list(opt.Noise_Amps), 'iters': iters,
          'state': [(k, tuple(v.shape)) for k, v in g.state_dict().items()],
          'state_d': [(k, tuple(v.shape)) for k, v in d.state_dict().items()]}
    optimizerD = optim.Adam(d.parameters(), lr=opt.lr_d, betas=(opt.beta1, 0.999))
    for block in g.body[:-opt.train_depth]:
        for param in block.parameters():
            param.requires_grad = False
    parameter_list = [{"params": block.parameters(), "lr": opt.lr_g * (opt.lr_scale ** (len(g.body[-opt.train_depth:]) - 1 - idx))}
                      for idx, block in enumerate(g.body[-opt.train_depth:])]
    if hasattr(g, 'head'):
        if opt.scale_idx - opt.train_depth < 0:
            parameter_list += [{"params": g.head.parameters(), "lr": opt.lr_g * (opt.lr_scale ** opt.scale_idx)}]
    if hasattr(g, 'tail'):
        parameter_list += [{"params": g.tail.parameters(), "lr": opt.lr_g}]
    optimizerG = optim.Adam(parameter_list, lr=opt.lr_g, betas=(opt.beta1, 0.999))
    # shapes of the per-stage noise of a 'rand' pass
    margin = opt.num_layer + 2 if gen_name == 'GeneratorSG' else opt.num_layer
    chans = 3 if gen_name == 'GeneratorSG' else opt.nfc
    noise_shapes = {}
    x = torch.zeros(1, 3, t0, s0, s0)
    for idx in range(1, scale_idx + 1):
        x = ref_utils.upscale(x, idx, opt)
        noise_shapes[idx] = (1, chans) + tuple(s + 2 * margin for s in x.shape[-3:])
    gen = torch.Generator().manual_seed(2000 + seed)
    draws, losses = [], []
    mse = torch.nn.MSELoss()
    for it in range(iters):
        dr = {'noise_init': torch.randn(tuple(z_init.shape), generator=gen),
              'noises': {lvl: torch.randn(shape, generator=gen) for lvl, shape in sorted(noise_shapes.items())},
              'alpha': float(torch.rand(1, generator=gen))}
        draws.append(dr)
        with DrawQueue() as q:
            q.normals = [dr['noise_init']] + [dr['noises'][lvl] for lvl in sorted(dr['noises'])]
            q.rands = [dr['alpha']]
            noise_init = ref_utils.generate_noise(ref=z_init)
            if it == 0:
                if scale_idx == 0:
                    opt.noise_amp = 1
                    opt.Noise_Amps.append(opt.noise_amp)
                else:
                    opt.Noise_Amps.append(0)
                    z_reconstruction = g(z_init, opt.Noise_Amps, mode="rec")
                    RMSE = torch.sqrt(F.mse_loss(real, z_reconstruction))
                    opt.noise_amp = opt.noise_amp_init * RMSE.item() / opt.batch_size
                    opt.Noise_Amps[-1] = opt.noise_amp
            d.zero_grad()
            errD_real = -d(real).mean()
            fake = g(noise_init, opt.Noise_Amps, mode="rand")
            errD_fake = d(fake.detach()).mean()
            gradient_penalty = calc_gradient_penalty(d, real, fake, opt.lambda_grad, 'cpu')
            (errD_real + errD_fake + gradient_penalty).backward()
            optimizerD.step()
            errG = -d(fake).mean() * opt.disc_loss_weight
            generated = g(z_init, opt.Noise_Amps, mode="rec")
            rec_loss = opt.alpha * mse(generated, real)
            g.zero_grad()
            (errG + rec_loss).backward()
            optimizerG.step()
        losses.append(dict(rec_loss=rec_loss.item(), errG=errG.item(), errD_real=errD_real.item(), errD_fake=errD_fake.item(),
                           gradient_penalty=gradient_penalty.item()))
    last_key = [k for k in g.state_dict() if k.startswith('body.%d.' % scale_idx) and k.endswith('conv.weight')][-1]
    fx.update({'draws': draws, 'losses': losses, 'noise_amps_after': list(opt.Noise_Amps), 'final_key': last_key,
               'final_weight': g.state_dict()[last_key].clone()})
    torch.save(fx, os.path.join(HERE, name + '.pt'))
    print(name, 'losses[0]', losses[0], 'losses[-1]', losses[-1], 'amps', opt.Noise_Amps)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'train_baselines':
        train_baselines_case('train_sg_tiny', 'GeneratorSG', tiny_opt(num_layer=2), scale_idx=1, iters=8)
        train_baselines_case('train_csg_tiny', 'GeneratorCSG', tiny_opt(num_layer=2), scale_idx=1, iters=8)
        train_baselines_case('train_sg_wide', 'GeneratorSG', port.Opt(nfc=64, num_layer=3, img_size=20, min_size=12, sampling_rates=[4, 2, 1]),
                             scale_idx=1, iters=4)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'baselines':
        csg_case('csg3d_tiny', tiny_opt(num_layer=2), stages=2, full_grads=True)
        csg_case('csg3d_wide', port.Opt(nfc=64, num_layer=3, img_size=20, min_size=12, sampling_rates=[4, 2, 1]), stages=1, full_grads=False)
        dbase_case('dbase3d_tiny', tiny_opt(num_layer=2), (1, 3, 4, 12, 11), True)
        dbase_case('dbase3d_wide', port.Opt(nfc=64, num_layer=3), (1, 3, 4, 14, 12), False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'train2d':
        # BASELINE configs[0]: train_image.py, 2-D HP-VAE-GAN (the same loop on networks_2d), a VAE level and a GAN level
        train_case('train_vae2d_tiny', tiny_opt(), scale_idx=1, iters=8, three_d=False)
        train_case('train_gan2d_tiny', tiny_opt(), scale_idx=2, iters=8, three_d=False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'train':
        train_case('train_vae_tiny', tiny_opt(), scale_idx=1, iters=12)
        train_case('train_gan_tiny', tiny_opt(), scale_idx=2, iters=8)
        train_case('train_gan_wide', port.Opt(nfc=64, latent_dim=128, num_layer=5, vae_levels=1, img_size=24, min_size=16,
                                              sampling_rates=[4, 2, 1]), scale_idx=1, iters=4)
        sys.exit(0)
    # tiny networks: every tensor and gradient stored (CUDA-core kernels on the GPU side)
    gen_case('hp3d_tiny', networks_3d, tiny_opt(), stages=3, three_d=True, full_grads=True)
    gen_case('hp3d_tiny_vae', networks_3d, tiny_opt(vae_levels=3), stages=2, three_d=True, full_grads=True)
    gen_case('hp2d_tiny', networks_2d, tiny_opt(), stages=3, three_d=False, full_grads=True)
    disc_case('d3d_tiny', networks_3d, tiny_opt(), True, (1, 3, 3, 12, 11), True)
    disc_case('d2d_tiny', networks_2d, tiny_opt(), False, (2, 3, 13, 12), True)
    sg_case('sg3d_tiny', tiny_opt(num_layer=2), stages=2)
    # 64-channel networks on small volumes: the tcgen05 kernels on the GPU side (gradient summaries only)
    wide_opt = port.Opt(nfc=64, latent_dim=128, num_layer=5, vae_levels=1, img_size=24, min_size=16, sampling_rates=[4, 2, 1])
    gen_case('hp3d_wide', networks_3d, wide_opt, stages=1, three_d=True, full_grads=False)
    disc_case('d3d_wide', networks_3d, wide_opt, True, (1, 3, 5, 20, 18), False)
    disc_case('d2d_wide', networks_2d, wide_opt, False, (1, 3, 24, 20), False)
