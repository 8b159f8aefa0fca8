"""CPU restatement of one pyramid scale of the reference's training loop (TEST INFRASTRUCTURE — see oracle/__init__.py).

Follows train_video.py:44-88 (optimizers and parameter groups) and :111-202 (one iteration) on top of oracle/port.py's
functional networks.  Pinned by tests/golden/train_*.pt, which record the losses of the UNMODIFIED reference modules
driven by the same loop (tests/golden/make_golden.py::train_case) — see tests/test_oracle.py::test_training_*.
bench.py times this as the CPU baseline / `--impl reference` arm.
"""
import torch
import torch.nn.functional as F

from . import port

TRAIN_DEFAULTS = dict(lr_g=5e-4, lr_d=5e-4, beta1=0.5, lambda_grad=0.1, rec_weight=10.0, kl_weight=1.0, disc_loss_weight=1.0,
                      lr_scale=0.2, train_depth=1, grad_clip=5.0, noise_amp_init=0.1, batch_size=1, train_all=False,
                      const_amp=False)

_NOT_PARAM = ('running_mean', 'running_var', 'num_batches_tracked', 'weight_u', 'weight_v')


def params_of(sd, prefix=''):
    """the tensors nn.Module.parameters() would yield for the sub-module at `prefix`, in state_dict order"""
    return [v for k, v in sd.items() if k.startswith(prefix) and not k.endswith(_NOT_PARAM)]


def make_leaf(sd):
    for k, v in sd.items():
        if not k.endswith(_NOT_PARAM):
            v.requires_grad_(True)
    return sd


def generator_param_groups(opt, sd_g):
    """train_video.py:57-86"""
    nbody = port.num_body(sd_g)

    def vae_groups():
        lr = opt.lr_g * (opt.lr_scale ** opt.scale_idx)
        return [{"params": params_of(sd_g, 'encode.'), "lr": lr}, {"params": params_of(sd_g, 'decoder.'), "lr": lr}]

    def body_groups(idxs):
        idxs = list(idxs)
        return [{"params": params_of(sd_g, 'body.%d.' % b), "lr": opt.lr_g * (opt.lr_scale ** (len(idxs) - 1 - i))}
                for i, b in enumerate(idxs)]

    all_idx = list(range(nbody))
    groups = []
    if not opt.train_all:
        if opt.vae_levels < opt.scale_idx + 1:
            depth = min(opt.train_depth, nbody - opt.vae_levels + 1)
            groups += body_groups(all_idx[-depth:])
        else:
            groups += vae_groups()
            groups += body_groups(all_idx[-opt.train_depth:])
    elif nbody < opt.train_depth:
        groups += vae_groups()
        groups += body_groups(all_idx)
    else:
        groups += body_groups(all_idx[-opt.train_depth:])
    return groups


class ScaleTrainer(object):
    def __init__(self, opt, sd_g, sd_d=None):
        for k, v in TRAIN_DEFAULTS.items():
            if not hasattr(opt, k):
                setattr(opt, k, v)
        self.opt, self.sd_g, self.sd_d = opt, make_leaf(sd_g), (make_leaf(sd_d) if sd_d is not None else None)
        self.gan = opt.vae_levels < opt.scale_idx + 1
        self.optimizerG = torch.optim.Adam(generator_param_groups(opt, sd_g), lr=opt.lr_g, betas=(opt.beta1, 0.999))
        self.optimizerD = torch.optim.Adam(params_of(sd_d), lr=opt.lr_d, betas=(opt.beta1, 0.999)) if self.gan else None
        self.iterations = 0

    @staticmethod
    def _zero(sd):
        for v in sd.values():
            v.grad = None

    def iteration(self, real, real_zero, noise_init=None, eps=None, noises=None, alpha=None, eps_amp=None):
        """train_video.py:126-202.  The random draws may be supplied (tests); otherwise they are taken from torch's
        generator in the reference's order."""
        opt, sd_g, sd_d = self.opt, self.sd_g, self.sd_d
        if noise_init is None:
            noise_init = torch.zeros(*opt.Z_init_size, device=real.device).normal_(0, 1)   # :126
        if self.iterations == 0 and len(opt.Noise_Amps) < opt.scale_idx + 1:               # :131-145
            if opt.const_amp:
                opt.Noise_Amps.append(1)
            elif opt.scale_idx == 0:
                opt.noise_amp = 1
                opt.Noise_Amps.append(1)
            else:
                opt.Noise_Amps.append(0)
                with torch.no_grad():
                    z_rec = port.generator(sd_g, opt, real_zero, opt.Noise_Amps, mode='rec', eps=eps_amp)[0]
                    opt.noise_amp = opt.noise_amp_init * torch.sqrt(F.mse_loss(real, z_rec)).item() / opt.batch_size
                opt.Noise_Amps[-1] = opt.noise_amp
        out = {}
        generated, generated_vae, (mu, logvar) = port.generator(sd_g, opt, real_zero, opt.Noise_Amps, mode='rec', eps=eps)
        if not self.gan:
            rec_vae_loss = F.mse_loss(generated, real) + F.mse_loss(generated_vae, real_zero)
            kl_loss = port.kl_criterion(mu, logvar)
            total_loss = opt.rec_weight * rec_vae_loss + opt.kl_weight * kl_loss
            out.update(rec_vae_loss=rec_vae_loss.detach(), kl_loss=kl_loss.detach())
        else:
            self._zero(sd_d)
            errD_real = -port.discriminator(sd_d, opt, real).mean()
            fake, _ = port.generator(sd_g, opt, None, opt.Noise_Amps, noise_init=noise_init, mode='rand', noises=noises)
            errD_fake = port.discriminator(sd_d, opt, fake.detach()).mean()
            gradient_penalty = port.gradient_penalty(sd_d, opt, real, fake, opt.lambda_grad, alpha=alpha)
            (errD_real + errD_fake + gradient_penalty).backward()
            self.optimizerD.step()
            rec_loss = F.mse_loss(generated, real)
            errG = -port.discriminator(sd_d, opt, fake).mean() * opt.disc_loss_weight
            total_loss = opt.rec_weight * rec_loss + errG
            out.update(rec_loss=rec_loss.detach(), errG=errG.detach(), errD_real=errD_real.detach(), errD_fake=errD_fake.detach(),
                       gradient_penalty=gradient_penalty.detach())
        self._zero(sd_g)
        total_loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in params_of(sd_g) if p.grad is not None], opt.grad_clip)
        self.optimizerG.step()
        out['total_loss'] = total_loss.detach()
        self.iterations += 1
        return out


BASELINE_DEFAULTS = dict(lr_g=5e-4, lr_d=5e-4, beta1=0.5, lambda_grad=0.1, alpha=10.0, disc_loss_weight=1.0, lr_scale=0.2,
                         train_depth=1, noise_amp_init=0.1, batch_size=1, Gsteps=1, Dsteps=1)


def _num_stages(sd_g):
    k = 0
    while any(key.startswith('body.%d.' % k) for key in sd_g):
        k += 1
    return k


def baseline_param_groups(opt, sd_g):
    """train_video_baselines.py:53-70: the last train_depth stages, the head while scale_idx < train_depth, the tail always
    (GeneratorSG has neither head nor tail attributes: its stages carry their own)"""
    n = _num_stages(sd_g)
    idxs = list(range(n))[-opt.train_depth:]
    groups = [{"params": params_of(sd_g, 'body.%d.' % b), "lr": opt.lr_g * (opt.lr_scale ** (len(idxs) - 1 - i))} for i, b in enumerate(idxs)]
    if any(k.startswith('head.') for k in sd_g) and opt.scale_idx - opt.train_depth < 0:
        groups.append({"params": params_of(sd_g, 'head.'), "lr": opt.lr_g * (opt.lr_scale ** opt.scale_idx)})
    if any(k.startswith('tail.') for k in sd_g):
        groups.append({"params": params_of(sd_g, 'tail.'), "lr": opt.lr_g})
    return groups


class BaselineTrainer(object):
    """One pyramid scale of train_video_baselines.py (:44-70 optimizers, :100-173 iteration) for GeneratorSG / GeneratorCSG
    with the WDiscriminator3D critic (the script's default), Dsteps = Gsteps = 1."""

    def __init__(self, opt, sd_g, sd_d, generator='GeneratorSG'):
        for k, v in BASELINE_DEFAULTS.items():
            if not hasattr(opt, k):
                setattr(opt, k, v)
        self.opt, self.sd_g, self.sd_d = opt, make_leaf(sd_g), make_leaf(sd_d)
        self.gen = port.generator_sg if generator == 'GeneratorSG' else port.generator_csg
        n = _num_stages(sd_g)
        for b in range(n - opt.train_depth):                         # :55-57
            for p in params_of(sd_g, 'body.%d.' % b):
                p.requires_grad_(False)
        self.optimizerD = torch.optim.Adam(params_of(sd_d), lr=opt.lr_d, betas=(opt.beta1, 0.999))
        self.optimizerG = torch.optim.Adam(baseline_param_groups(opt, sd_g), lr=opt.lr_g, betas=(opt.beta1, 0.999))
        self.iterations = 0

    def iteration(self, real, z_init, noise_init=None, noises=None, alpha=None):
        opt, sd_g, sd_d = self.opt, self.sd_g, self.sd_d
        if noise_init is None:
            noise_init = torch.zeros_like(z_init).normal_(0, 1)                              # :107
        if self.iterations == 0:                                                             # :112-122
            if opt.scale_idx == 0:
                opt.noise_amp = 1
                opt.Noise_Amps.append(opt.noise_amp)
            else:
                opt.Noise_Amps.append(0)
                z_rec = self.gen(sd_g, opt, z_init, opt.Noise_Amps, mode='rec')
                opt.noise_amp = opt.noise_amp_init * torch.sqrt(F.mse_loss(real, z_rec)).item() / opt.batch_size
                opt.Noise_Amps[-1] = opt.noise_amp
        ScaleTrainer._zero(sd_d)                                                             # :131
        errD_real = -port.discriminator(sd_d, opt, real).mean()
        fake = self.gen(sd_g, opt, noise_init, opt.Noise_Amps, mode='rand', noises=noises)
        errD_fake = port.discriminator(sd_d, opt, fake.detach()).mean()
        gradient_penalty = port.gradient_penalty(sd_d, opt, real, fake, opt.lambda_grad, alpha=alpha)
        (errD_real + errD_fake + gradient_penalty).backward()
        self.optimizerD.step()
        errG = -port.discriminator(sd_d, opt, fake).mean() * opt.disc_loss_weight           # :158-160
        generated = self.gen(sd_g, opt, z_init, opt.Noise_Amps, mode='rec')
        rec_loss = opt.alpha * F.mse_loss(generated, real)
        ScaleTrainer._zero(sd_g)
        (errG + rec_loss).backward()
        self.optimizerG.step()
        self.iterations += 1
        return dict(rec_loss=rec_loss.detach(), errG=errG.detach(), errD_real=errD_real.detach(), errD_fake=errD_fake.detach(),
                    gradient_penalty=gradient_penalty.detach())


@torch.no_grad()
def generate(sd_g, opt, n_samples, batch=1):
    """train_video.py:226-235: fresh z per draw, G(z, amps, noise_init=z, mode='rand')"""
    frames, fake = 0, None
    size = list(opt.Z_init_size)
    for i in range(0, n_samples, batch):
        size[0] = min(batch, n_samples - i)
        z = torch.zeros(*size).normal_(0, 1)
        fake, _ = port.generator(sd_g, opt, None, opt.Noise_Amps, noise_init=z, mode='rand')
        frames += fake.shape[0] * (fake.shape[2] if fake.dim() == 5 else 1)
    return frames, fake
