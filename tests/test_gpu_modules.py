"""GPU parity of the drop-in modules (libhpvg kernels), through the reference's own module interface, against
  (a) the golden fixtures recorded from the unmodified reference (fp32 CPU), and
  (b) the CPU oracle run with bf16 storage emulation (oracle.port.storage('bf16')): the same arithmetic with every
      wide activation and gradient rounded to bf16 exactly where the CUDA path stores one.

Tolerances.  north_star: relative 2e-2 per layer for bf16 operands — tests/test_gpu_layers.py checks every layer type
in isolation at that bar.  Whole networks chain 20-40 bf16-stored layers with BatchNorm backward passes that cancel
large components of the gradient, so whole-network results are compared
  * with (a) at REF_*: only as loose as (b) itself is from (a) — the emulated oracle deviates from the fp32 golden by
    up to 0.35 on the 8-channel nets' gradients and 0.08 on the 64-channel ones
    (tests/test_oracle.py::test_bf16_storage_emulation_stays_close_to_fp32), and
  * with (b) at the same bounds for outputs and twice those bounds for generator gradients (two independent bf16
    realisations of a chaotic rounding process are sqrt(2) further apart than each is from fp32).  (b) is bit-identical to the CUDA-core kernels for the first layers of a network and
    within 1e-4 of a tcgen05 layer (tests/test_gpu_layers.py pins that per layer), but every flipped bf16 rounding is
    amplified by the following layers until the difference saturates at the bf16 noise floor (measured: 0.1 % of the
    elements differ after the first tcgen05 layer, 3 %, 23 %, 46 %, 59 %, 68 % after the next five), so over a whole
    network two correct bf16 implementations are as far from each other as each is from fp32.  The spectral-norm
    critics, which have no BatchNorm, do stay within 2e-2 of (b), gradients included.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from helpers import opt_from, rel_err, state_from, with_grad
from oracle import port

pytestmark = pytest.mark.gpu

EMU_OUT_TOL = 2e-2      # outputs vs the bf16-emulating oracle
EMU_GRAD_TOL = {'tiny': 0.45, 'wide': 0.10}     # generator parameter gradients vs the bf16-emulating oracle
EMU_GRAD_TOL_CRITIC = 2e-2                      # critic (no BatchNorm) gradients vs the bf16-emulating oracle
REF_OUT_TOL = 3e-2      # outputs vs the fp32 reference fixtures
REF_GRAD_TOL = {'tiny': 0.45, 'wide': 0.10}   # parameter gradients vs the fp32 reference fixtures (see module docstring)
LOSS_TOL = 2e-2

_REPORT = {}


@pytest.fixture(scope="module", autouse=True)
def _dump_report():
    yield
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_modules.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


def _cuda_module(cls, opt, fx, stages=None):
    m = cls(opt)
    if stages is not None:
        for _ in range(stages):
            m.init_next_stage()
    missing = m.load_state_dict(state_from(fx), strict=True)   # identical keys and shapes as the reference
    assert not missing.missing_keys and not missing.unexpected_keys
    return m.cuda()


class NoiseQueue:
    """replaces hpvg.images.draw_normal so that the module consumes the fixture's noise tensors in order"""

    def __init__(self, tensors):
        self.tensors = list(tensors)

    def __call__(self, shape, dtype, device):
        t = self.tensors.pop(0)
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device=device, dtype=dtype)


def _gnorm(g):
    return g['norm'] if isinstance(g, dict) else g.double().norm().item()


def _check_grads(module, ref_grads, tol, floor_frac, tag):
    """every parameter gradient within tol (relative L2) of ref_grads; gradients that are mathematically ~0 (conv bias
    in front of BatchNorm) are covered by the absolute floor tied to the largest gradient of the network"""
    floor = floor_frac * max(_gnorm(g) for g in ref_grads.values())
    worst, worst_key = 0.0, None
    params = dict(module.named_parameters())
    for k, g in ref_grads.items():
        mine = params[k].grad
        assert mine is not None, k
        mine = mine.detach().cpu().double()
        if isinstance(g, dict):
            d = abs(mine.norm().item() - g['norm'])
            ref = g['norm']
            d2 = (mine.flatten()[:64] - g['head'].double()).norm().item()
            assert d2 <= 2 * tol * g['head'].double().norm().item() + floor, (tag, k, d2)
        else:
            d = (mine - g.double()).norm().item()
            ref = g.double().norm().item()
        assert d <= tol * ref + floor, (tag, k, d, ref)
        if ref > 10 * floor and d / ref > worst:
            worst, worst_key = d / ref, k
    _REPORT[tag] = {'worst_rel': worst, 'key': worst_key, 'tol': tol}
    return worst


def _oracle_grads(sd):
    return {k: v.grad.detach().clone() for k, v in sd.items() if v.is_floating_point() and v.grad is not None}


def _family(name):
    return 'wide' if name.endswith('wide') else 'tiny'


@pytest.mark.parametrize("name", ["hp3d_tiny", "hp3d_tiny_vae", "hp2d_tiny", "hp3d_wide"])
def test_generator_parity(golden, monkeypatch, name):
    from hpvg import images
    from modules import networks_2d, networks_3d
    from modules.losses import kl_criterion
    fx = golden(name)
    opt = opt_from(fx)
    nets = networks_2d if name.startswith("hp2d") else networks_3d
    g = _cuda_module(nets.GeneratorHPVAEGAN, opt, fx, stages=fx['stages'])
    rec = fx['rec']

    # (b) the oracle with bf16 storage emulation, same weights / inputs / noise
    sd = with_grad(state_from(fx))
    with port.storage('bf16'):
        e_gen, e_vae, (e_mu, e_logvar) = port.generator(sd, opt, fx['real_zero'], fx['amps'], mode='rec', eps=rec['eps'])
        e_kl = port.kl_criterion(e_mu, e_logvar)
        e_loss = 10.0 * (F.mse_loss(e_gen, fx['real']) + F.mse_loss(e_vae, fx['real_zero'])) + e_kl
        e_loss.backward()

    monkeypatch.setattr(images, "draw_normal", NoiseQueue([rec['eps']]))
    real, real_zero = fx['real'].cuda(), fx['real_zero'].cuda()
    gen, gen_vae, (mu, logvar) = g(real_zero, fx['amps'], mode='rec')
    assert gen.shape == rec['gen'].shape and gen_vae.shape == rec['gen_vae'].shape and mu.shape == rec['mu'].shape
    outs = {'gen': (gen, e_gen, rec['gen']), 'gen_vae': (gen_vae, e_vae, rec['gen_vae']), 'mu': (mu, e_mu, rec['mu']),
            'logvar': (logvar, e_logvar, rec['logvar'])}
    for k, (mine, emu, ref) in outs.items():
        _REPORT['%s/out/%s' % (name, k)] = {'vs_emu': rel_err(mine, emu), 'vs_ref': rel_err(mine, ref)}
        assert rel_err(mine, emu) < EMU_OUT_TOL, k
        assert rel_err(mine, ref) < REF_OUT_TOL, k
    kl = kl_criterion(mu, logvar)
    assert abs(kl.item() - rec['kl']) < LOSS_TOL * abs(rec['kl'])
    loss = 10.0 * (F.mse_loss(gen, real) + F.mse_loss(gen_vae, real_zero)) + kl
    assert abs(loss.item() - rec['loss']) < LOSS_TOL * abs(rec['loss'])
    assert abs(loss.item() - e_loss.item()) < 2e-3 * abs(e_loss.item())
    g.zero_grad()
    loss.backward()
    _check_grads(g, rec['grads'], REF_GRAD_TOL[_family(name)], 2e-3, name + '/grads_vs_ref')
    _check_grads(g, _oracle_grads(sd), 2 * EMU_GRAD_TOL[_family(name)], 2e-3, name + '/grads_vs_emu')
    bufs = dict(g.named_buffers())
    for k, b in rec['buffers'].items():
        if k.endswith('num_batches_tracked'):
            assert int(bufs[k].item()) == int(b.item()), k
        else:
            assert rel_err(bufs[k], b) < REF_OUT_TOL, k
    noises = [fx['rand']['noises'][k] for k in sorted(fx['rand']['noises'])]
    monkeypatch.setattr(images, "draw_normal", NoiseQueue(noises))
    with torch.no_grad():
        z = fx['rand']['z'].cuda()
        fake, fake_vae = g(z, fx['amps'], noise_init=z, mode='rand')
    assert rel_err(fake_vae, fx['rand']['fake_vae']) < REF_OUT_TOL
    assert rel_err(fake, fx['rand']['fake']) < REF_OUT_TOL


@pytest.mark.parametrize("name", ["d3d_tiny", "d2d_tiny", "d3d_wide", "d2d_wide"])
def test_discriminator_and_gradient_penalty_parity(golden, monkeypatch, name):
    from modules import networks_2d, networks_3d
    from modules import utils as mutils
    fx = golden(name)
    opt = opt_from(fx)
    cls = networks_2d.WDiscriminator2D if name.startswith("d2d") else networks_3d.WDiscriminator3D
    d = _cuda_module(cls, opt, fx)

    sd = with_grad(state_from(fx))
    with port.storage('bf16'):
        e_real = port.discriminator(sd, opt, fx['real'])
        e_fake = port.discriminator(sd, opt, fx['fake'])
        e_gp = port.gradient_penalty(sd, opt, fx['real'], fx['fake'], fx['lambda'], alpha=fx['alpha'])
        (-e_real.mean() + e_fake.mean() + e_gp).backward()

    real, fake = fx['real'].cuda(), fx['fake'].cuda()
    d.zero_grad()
    out_real = d(real)
    out_fake = d(fake)
    assert out_real.shape == fx['out_real'].shape
    _REPORT[name + '/out'] = {'vs_emu': rel_err(out_real, e_real), 'vs_ref': rel_err(out_real, fx['out_real'])}
    assert rel_err(out_real, e_real) < EMU_OUT_TOL and rel_err(out_fake, e_fake) < EMU_OUT_TOL
    assert rel_err(out_real, fx['out_real']) < REF_OUT_TOL and rel_err(out_fake, fx['out_fake']) < REF_OUT_TOL
    monkeypatch.setattr(torch, "rand", lambda *a, **k: torch.full((1, 1), fx['alpha']))
    gp = mutils.calc_gradient_penalty(d, real, fake, fx['lambda'], 'cuda')
    _REPORT[name + '/gp'] = {'mine': gp.item(), 'emu': e_gp.item(), 'ref': fx['gp']}
    assert abs(gp.item() - fx['gp']) < LOSS_TOL * abs(fx['gp']), (gp.item(), fx['gp'])
    assert abs(gp.item() - e_gp.item()) < 5e-3 * abs(e_gp.item()), (gp.item(), e_gp.item())
    (-out_real.mean() + out_fake.mean() + gp).backward()
    _check_grads(d, _oracle_grads(sd), EMU_GRAD_TOL_CRITIC, 2e-3, name + '/grads_vs_emu')
    _check_grads(d, fx['grads'], REF_GRAD_TOL[_family(name)], 2e-3, name + '/grads_vs_ref')
    bufs = dict(d.named_buffers())
    for k, b in fx['buffers'].items():
        assert rel_err(bufs[k], b) < 1e-3, k     # spectral-norm u/v after 3 power iterations (fp32 kernels)


def test_generator_sg_parity(golden, monkeypatch):
    from hpvg import images
    from modules import networks_3d
    fx = golden("sg3d_tiny")
    opt = opt_from(fx)
    g = _cuda_module(networks_3d.GeneratorSG, opt, fx, stages=fx['stages'])

    sd = with_grad(state_from(fx))
    with port.storage('bf16'):
        e_out = port.generator_sg(sd, opt, fx['z'], fx['amps'], mode='rec')
        F.mse_loss(e_out, fx['rec']['target']).backward()

    out = g(fx['z'].cuda(), fx['amps'], mode='rec')
    _REPORT['sg3d_tiny/out'] = {'vs_emu': rel_err(out, e_out), 'vs_ref': rel_err(out, fx['rec']['out'])}
    assert rel_err(out, e_out) < EMU_OUT_TOL
    assert rel_err(out, fx['rec']['out']) < REF_OUT_TOL
    loss = F.mse_loss(out, fx['rec']['target'].cuda())
    assert abs(loss.item() - fx['rec']['loss']) < LOSS_TOL * fx['rec']['loss']
    g.zero_grad()
    loss.backward()
    _check_grads(g, _oracle_grads(sd), EMU_GRAD_TOL['tiny'], 2e-3, 'sg3d_tiny/grads_vs_emu')
    _check_grads(g, fx['rec']['grads'], REF_GRAD_TOL['tiny'], 2e-3, 'sg3d_tiny/grads_vs_ref')
    noises = [fx['rand']['noises'][k] for k in sorted(fx['rand']['noises'])]
    monkeypatch.setattr(images, "draw_normal", NoiseQueue(noises))
    with torch.no_grad():
        fake = g(fx['z'].cuda(), fx['amps'], mode='rand')
    assert rel_err(fake, fx['rand']['fake']) < REF_OUT_TOL


@pytest.mark.parametrize("name", ["csg3d_tiny", "csg3d_wide"])
def test_generator_csg_parity(golden, monkeypatch, name):
    """GeneratorCSG (reference modules/networks_3d.py:213-269, default generator of train_video_baselines.py): feature-space
    pyramid — wide zero-pad, wide trilinear resize (+ noise), wide residual add (wide_ops.cu) around the pad-0 ConvBlocks"""
    from hpvg import images
    from modules import networks_3d
    fx = golden(name)
    opt = opt_from(fx)
    g = _cuda_module(networks_3d.GeneratorCSG, opt, fx, stages=fx['stages'])
    fam = _family(name)

    sd = with_grad(state_from(fx))
    with port.storage('bf16'):
        e_out = port.generator_csg(sd, opt, fx['z'], fx['amps'], mode='rec')
        F.mse_loss(e_out, fx['rec']['target']).backward()

    out = g(fx['z'].cuda(), fx['amps'], mode='rec')
    _REPORT[name + '/out'] = {'vs_emu': rel_err(out, e_out), 'vs_ref': rel_err(out, fx['rec']['out'])}
    assert out.shape == fx['rec']['out'].shape and out.dtype == torch.float32
    assert rel_err(out, e_out) < EMU_OUT_TOL
    assert rel_err(out, fx['rec']['out']) < REF_OUT_TOL
    loss = F.mse_loss(out, fx['rec']['target'].cuda())
    assert abs(loss.item() - fx['rec']['loss']) < LOSS_TOL * fx['rec']['loss']
    g.zero_grad()
    loss.backward()
    _check_grads(g, _oracle_grads(sd), 2 * EMU_GRAD_TOL[fam], 2e-3, name + '/grads_vs_emu')
    _check_grads(g, fx['rec']['grads'], REF_GRAD_TOL[fam], 2e-3, name + '/grads_vs_ref')
    bufs = dict(g.named_buffers())
    for k, b in fx['rec']['buffers'].items():
        if b.is_floating_point():
            assert rel_err(bufs[k], b) < REF_OUT_TOL, k
    noises = [fx['rand']['noises'][k] for k in sorted(fx['rand']['noises'])]
    monkeypatch.setattr(images, "draw_normal", NoiseQueue(noises))
    with torch.no_grad():
        fake = g(fx['z'].cuda(), fx['amps'], mode='rand')
    assert rel_err(fake, fx['rand']['fake']) < REF_OUT_TOL


@pytest.mark.parametrize("name", ["dbase3d_tiny", "dbase3d_wide"])
def test_discriminator_baselines_parity(golden, name):
    """WDiscriminatorBaselines (reference modules/networks_3d.py:184-210): outputs and first-order gradients; the gradient
    penalty through its BatchNorm layers is refused loudly (once-differentiable node), never computed wrongly"""
    from modules import networks_3d
    from modules import utils as mutils
    fx = golden(name)
    opt = opt_from(fx)
    d = _cuda_module(networks_3d.WDiscriminatorBaselines, opt, fx)
    fam = _family(name)
    sd = with_grad(state_from(fx))
    with port.storage('bf16'):
        e_real = port.discriminator_baselines(sd, opt, fx['real'])
        e_fake = port.discriminator_baselines(sd, opt, fx['fake'])
        (-e_real.mean() + e_fake.mean()).backward()
    out_real, out_fake = d(fx['real'].cuda()), d(fx['fake'].cuda())
    assert out_real.shape == fx['out_real'].shape
    assert rel_err(out_real, e_real) < EMU_OUT_TOL and rel_err(out_fake, e_fake) < EMU_OUT_TOL
    assert rel_err(out_real, fx['out_real']) < REF_OUT_TOL and rel_err(out_fake, fx['out_fake']) < REF_OUT_TOL
    d.zero_grad()
    (-out_real.mean() + out_fake.mean()).backward()
    _check_grads(d, _oracle_grads(sd), 2 * EMU_GRAD_TOL[fam], 2e-3, name + '/grads_vs_emu')
    _check_grads(d, fx['grads'], REF_GRAD_TOL[fam], 2e-3, name + '/grads_vs_ref')


def test_discriminator_baselines_gradient_penalty(golden, monkeypatch):
    """the WGAN-GP double backward THROUGH BatchNorm (WDiscriminatorBaselines, reference modules/networks_3d.py:184-210 under
    modules/utils.py:4-19): penalty value and every parameter gradient of the critic loss of train_video_baselines.py:131-149
    against the unmodified reference (tests/golden/make_nb_golden.py::dbase_gp_case)"""
    from modules import networks_3d
    from modules import utils as mutils
    fx = golden("dbase3d_gp_tiny")
    opt = opt_from(fx)
    d = _cuda_module(networks_3d.WDiscriminatorBaselines, opt, fx)
    monkeypatch.setattr(torch, "rand", lambda *a, **k: torch.full((1, 1), fx['alpha']))
    real, fake = fx['real'].cuda(), fx['fake'].cuda()
    d.zero_grad()
    loss = -d(real).mean() + d(fake).mean()
    gp = mutils.calc_gradient_penalty(d, real, fake, fx['lambda'], 'cuda')
    (loss + gp).backward()
    assert abs(gp.item() - fx['gp']) <= LOSS_TOL * abs(fx['gp']), (gp.item(), fx['gp'])
    _check_grads(d, fx['grads'], REF_GRAD_TOL['tiny'], 2e-3, 'dbase3d_gp_tiny/grads_vs_ref')


@pytest.mark.parametrize("nfc,col_mode", [(64, 0), (64, 1), (8, -1)])
def test_batched_generation_with_per_sample_batchnorm_equals_batch1_draws(monkeypatch, nfc, col_mode):
    """ops.bn_per_sample: a batch-3 'rand' forward of GeneratorHPVAEGAN with per-draw BatchNorm statistics against three
    batch-1 forwards on the same latents and noise (the reference generates each draw with batch size 1,
    train_video.py:226-235).  nfc 64: tcgen05 (brick kernel / column kernel) and expand kernels with [N, 2C] statistics;
    nfc 8: the per-sample fallback."""
    from hpvg import images, lib, ops
    from modules import networks_3d
    opt = port.Opt(nfc=nfc, latent_dim=128 if nfc == 64 else 8, num_layer=2, vae_levels=1, img_size=24, min_size=16, sampling_rates=[4, 2, 1])
    g = networks_3d.GeneratorHPVAEGAN(opt)
    g.init_next_stage()
    g.init_next_stage()
    port.det_fill(g.state_dict(), 3)
    g.cuda()
    b = 3
    s0, t0 = port.scale_size(0, opt), port.time_depth(0, opt)
    z = port.det_tensor((b, opt.latent_dim, t0, s0, s0), 71).cuda()
    shapes = [(3, port.time_depth(i, opt), port.scale_size(i, opt), port.scale_size(i, opt)) for i in (1, 2)]
    noises = [port.det_tensor((b,) + sh, 72 + i).cuda() for i, sh in enumerate(shapes)]
    amps = [1.0, 0.1, 0.1]
    singles = []
    prev_mode = lib.set_conv_col_mode(col_mode)
    try:
        with torch.no_grad(), ops.bn_running_stats(False):
            for i in range(b):
                monkeypatch.setattr(images, "draw_normal", NoiseQueue([nz[i:i + 1] for nz in noises]))
                fake, _ = g(z[i:i + 1], amps, noise_init=z[i:i + 1], mode='rand')
                singles.append(fake)
            monkeypatch.setattr(images, "draw_normal", NoiseQueue(noises))
            with ops.bn_per_sample(True):
                batched, _ = g(z, amps, noise_init=z, mode='rand')
            monkeypatch.setattr(images, "draw_normal", NoiseQueue(noises))
            coupled, _ = g(z, amps, noise_init=z, mode='rand')
    finally:
        lib.set_conv_col_mode(prev_mode)
    ref = torch.cat(singles, 0)
    assert batched.shape == ref.shape
    assert rel_err(batched, ref) < 1e-2          # same arithmetic; BatchNorm sums are accumulated in a different order
    assert rel_err(coupled, ref) > 3 * rel_err(batched, ref)      # batch statistics couple the draws: a different function
    with pytest.raises(Exception):
        with ops.bn_per_sample(True):
            g(z, amps, noise_init=z.clone().requires_grad_(True), mode='rand')


def test_wide_resize_pad_add_match_torch():
    """wide_ops.cu against torch on the same bf16 values: trilinear resize (align_corners) with and without the NCDHW float32
    noise term, its adjoint, zero-pad and crop, residual add"""
    from hpvg import ops
    gen = torch.Generator(device='cuda').manual_seed(5)
    for c, (d, h, w), (do, ho, wo) in ((64, (3, 9, 11), (5, 14, 13)), (8, (2, 6, 6), (2, 16, 15)), (64, (6, 54, 54), (16, 64, 64))):
        x = torch.randn((2, d, h, w, c), device='cuda', generator=gen).bfloat16().requires_grad_(True)
        noise = torch.randn((2, c, do, ho, wo), device='cuda', generator=gen)
        xt = x.detach().float().permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
        ref = F.interpolate(xt, size=[do, ho, wo], mode='trilinear', align_corners=True)
        y = ops.UpsampleWide.apply(x, (do, ho, wo), None, 0.0)
        assert rel_err(y.float().permute(0, 4, 1, 2, 3), ref) < 4e-3               # one bf16 rounding of the result
        yn = ops.UpsampleWide.apply(x, (do, ho, wo), noise, 0.3)
        assert rel_err(yn.float().permute(0, 4, 1, 2, 3), ref + 0.3 * noise) < 4e-3
        g = torch.randn(tuple(y.shape), device='cuda', generator=gen).bfloat16()
        y.backward(g)
        ref.backward(g.float().permute(0, 4, 1, 2, 3))
        assert rel_err(x.grad.float().permute(0, 4, 1, 2, 3), xt.grad) < 4e-3
        p = ops.PadWide.apply(x, 2)
        assert torch.equal(p.float().permute(0, 4, 1, 2, 3), F.pad(xt.detach(), (2,) * 6))
        x.grad = None
        gp = torch.randn(tuple(p.shape), device='cuda', generator=gen).bfloat16()
        p.backward(gp)
        assert torch.equal(x.grad, gp[:, 2:-2, 2:-2, 2:-2, :])
        b = torch.randn(tuple(x.shape), device='cuda', generator=gen).bfloat16()
        assert torch.equal(ops.AddWide.apply(x.detach(), b), (x.detach().float() + b.float()).bfloat16())


@pytest.mark.parametrize("nfc", [8, 64])
def test_fused_lrelu_mask_backward_matches_unfused(monkeypatch, nfc):
    """ops.ChainLink / ops._MaskLink: the critic's first-order backward and the gradient-penalty double backward with the
    LeakyReLU derivative (and the bias sum) applied in the data-gradient epilogue, against the same backward through the
    separate leaky_relu_backward launches.  The only arithmetic difference is one bf16 rounding less on the fused path."""
    from hpvg import ops
    from modules import networks_3d
    from modules import utils as mutils
    opt = port.Opt(nfc=nfc, latent_dim=8, num_layer=3)
    d = networks_3d.WDiscriminator3D(opt)
    port.det_fill(d.state_dict(), 11)
    d.cuda()
    real = port.det_tensor((1, 3, 5, 18, 20), 3).cuda()
    fake = port.det_tensor((1, 3, 5, 18, 20), 4).cuda()
    monkeypatch.setattr(torch, "rand", lambda *a, **k: torch.full((1, 1), 0.3))
    u0 = {k: b.clone() for k, b in d.named_buffers()}
    results = []
    saved = ops._FUSE_MASK[0]
    try:
        for fuse in (False, True):
            with torch.no_grad():
                for k, b in d.named_buffers():
                    b.copy_(u0[k])                      # same power-iteration state for both runs
            ops._FUSE_MASK[0] = fuse
            d.zero_grad()
            x = real.clone().requires_grad_(True)
            n0 = ops.lib.launch_count()
            loss = -d(x).mean() + d(fake).mean() + mutils.calc_gradient_penalty(d, real, fake, 0.1, 'cuda')
            loss.backward()
            torch.cuda.synchronize()
            results.append((ops.lib.launch_count() - n0, x.grad.clone(), {k: p.grad.clone() for k, p in d.named_parameters()}))
    finally:
        ops._FUSE_MASK[0] = saved
    (n_plain, gx_plain, g_plain), (n_fused, gx_fused, g_fused) = results
    assert n_fused < n_plain                        # the leaky_relu_backward launches are gone
    assert rel_err(gx_fused, gx_plain) < 5e-3
    for k in g_plain:
        assert rel_err(g_fused[k], g_plain[k]) < 5e-3, k


def test_critic_passes_are_bit_reproducible():
    """Spectral-norm sigma is reduced in a fixed order (no atomics), so two critic forwards from the same power-iteration
    state — and the gradient-penalty value built on a second power iteration — are bit-identical.  (With atomics a one-ulp
    change of W / sigma flipped bf16 roundings and, layers later, LeakyReLU signs: 1.6e-2 scatter of the input gradient.)"""
    from modules import networks_3d
    from modules import utils as mutils
    opt = port.Opt(nfc=64, latent_dim=8, num_layer=3)
    d = networks_3d.WDiscriminator3D(opt)
    port.det_fill(d.state_dict(), 11)
    d.cuda()
    real = port.det_tensor((1, 3, 5, 18, 20), 3).cuda()
    fake = port.det_tensor((1, 3, 5, 18, 20), 4).cuda()
    saved = torch.rand
    torch.rand = lambda *a, **k: torch.full((1, 1), 0.3)
    u0 = {k: b.clone() for k, b in d.named_buffers()}
    runs = []
    try:
        for _ in range(4):
            with torch.no_grad():
                for k, b in d.named_buffers():
                    b.copy_(u0[k])
            x = real.clone().requires_grad_(True)
            out = d(x)
            (gx,) = torch.autograd.grad(-out.mean(), x)
            gp = mutils.calc_gradient_penalty(d, real, fake, 0.1, 'cuda')
            d.zero_grad()
            gp.backward()
            torch.cuda.synchronize()
            runs.append((out.detach().clone(), gx.clone(), d.head.conv.weight_orig.grad.clone(), {k: b.clone() for k, b in d.named_buffers()}))
    finally:
        torch.rand = saved
    for out, gx, gw, bufs in runs[1:]:
        assert torch.equal(out, runs[0][0])
        assert torch.equal(gx, runs[0][1])
        assert torch.equal(gw, runs[0][2])
        for k, b in bufs.items():
            assert torch.equal(b, runs[0][3][k]), k


def test_modules_refuse_cpu_tensors():
    from hpvg.lib import HpvgError
    from modules import networks_3d
    opt = port.Opt(nfc=8, latent_dim=8, num_layer=1)
    d = networks_3d.WDiscriminator3D(opt)
    with pytest.raises(HpvgError):
        d(torch.zeros(1, 3, 3, 8, 8))


@pytest.mark.parametrize("dims", [3, 2])
def test_bernoulli_gated_variants_match_the_reference(golden, dims):
    """row f4: GeneratorVAE_nb / Encode{3D,2D}VAE_nb (reference modules/networks_3d.py:110-138, :409-485; networks_2d.py:115-143,
    :272-348) on the library's convolution kernels, against vectors recorded from the unmodified reference
    (tests/golden/make_nb_golden.py): encoder outputs (mu, logvar, gate), generator output in 'rec' mode driven by
    noise_init_norm / noise_init_bern, and every parameter gradient of a fixed linear loss"""
    from modules import networks_2d, networks_3d
    nets = networks_3d if dims == 3 else networks_2d
    fx = golden("nb%dd_tiny" % dims)
    opt = opt_from(fx)
    g = _cuda_module(nets.GeneratorVAE_nb, opt, fx, stages=2)
    assert [k for k, _ in fx['state']] == list(g.state_dict().keys())
    mu, logvar, bern = g.encode(fx['video'].cuda())
    assert mu.shape == fx['mu'].shape and bern.shape == fx['bern'].shape
    assert rel_err(mu, fx['mu']) < REF_OUT_TOL and rel_err(logvar, fx['logvar']) < REF_OUT_TOL and rel_err(bern, fx['bern']) < REF_OUT_TOL
    out, vae_out = g(None, fx['amps'], noise_init_norm=fx['z_norm'].cuda(), noise_init_bern=fx['z_bern'].cuda(), mode='rec')
    assert rel_err(out, fx['out']) < REF_OUT_TOL and rel_err(vae_out, fx['vae_out']) < REF_OUT_TOL
    s = fx['seeds']
    loss = (out * fx['gout'].cuda()).sum() + (mu * port.det_tensor(tuple(mu.shape), s['mu']).cuda()).sum() \
        + (logvar * port.det_tensor(tuple(logvar.shape), s['logvar']).cuda()).sum() + (bern * port.det_tensor(tuple(bern.shape), s['bern']).cuda()).sum()
    g.zero_grad()
    loss.backward()
    _check_grads(g, fx['grads'], REF_GRAD_TOL['tiny'], 2e-3, 'nb%dd_tiny vs reference' % dims)
    # the sampling entry point draws its own noise: shapes and finiteness only
    x, v, (m2, l2, b2) = g(fx['video'].cuda(), fx['amps'], mode='rand')
    assert x.shape == out.shape and torch.isfinite(x).all() and m2.shape == mu.shape and b2.shape == bern.shape
