"""End-of-scale criterion of north_star: after a fixed number of training iterations on identical weights, data and
random draws, the reconstruction loss of the CUDA path must agree with the reference's within 1 %.

The golden fixtures (tests/golden/train_*.pt) hold the per-iteration losses of the UNMODIFIED reference modules
stepping the loop of train_video.py:111-202 (Adam on generator and critic, WGAN-GP, gradient clipping).
"""
import pytest
import torch

from helpers import state_d_from, state_from, train_opt_from

pytestmark = pytest.mark.gpu

REC_TOL = 1e-2        # north_star: end-of-scale reconstruction loss within 1 %


class DrawQueue:
    def __init__(self):
        self.tensors = []

    def __call__(self, shape, dtype, device):
        t = self.tensors.pop(0)
        assert tuple(t.shape) == tuple(shape), (tuple(t.shape), tuple(shape))
        return t.to(device=device, dtype=dtype)


@pytest.mark.parametrize("name", ["train_vae_tiny", "train_gan_tiny", "train_gan_wide", "train_vae2d_tiny", "train_gan2d_tiny"])
def test_reconstruction_loss_after_k_iterations(golden, monkeypatch, name):
    """3-D fixtures: train_video.py's loop; *2d*: train_image.py's (BASELINE configs[0]: the same loop on networks_2d)"""
    from hpvg import images, train
    from modules import networks_2d, networks_3d
    fx = golden(name)
    opt = train_opt_from(fx)
    three_d = fx.get('three_d', True)
    nets = networks_3d if three_d else networks_2d
    g = nets.GeneratorHPVAEGAN(opt)
    for _ in range(fx['stages']):
        g.init_next_stage()
    g.load_state_dict(state_from(fx), strict=True)
    g.cuda()
    d = None
    if 'state_d' in fx:
        d = networks_3d.WDiscriminator3D(opt) if three_d else networks_2d.WDiscriminator2D(opt)
        d.load_state_dict(state_d_from(fx), strict=True)
        d.cuda()
    tr = train.ScaleTrainer(opt, g, d)
    q = DrawQueue()
    monkeypatch.setattr(images, "draw_normal", q)
    alphas = []
    monkeypatch.setattr(torch, "rand", lambda *a, **k: torch.full((1, 1), alphas.pop(0)))
    real, real_zero = fx['real'].cuda(), fx['real_zero'].cuda()
    rec_key = 'rec_loss' if d is not None else 'rec_vae_loss'
    history = []
    for it in range(fx['iters']):
        dr = fx['draws'][it]
        q.tensors = [dr['noise_init']] + ([dr['eps_amp']] if 'eps_amp' in dr else []) + [dr['eps']]
        if d is not None:
            q.tensors += [dr['noises'][lvl] for lvl in sorted(dr['noises'])]
            alphas.append(dr['alpha'])
        out = tr.iteration(real, real_zero)
        assert not q.tensors and not alphas, "the CUDA path draws in a different order than the reference"
        history.append({k: v.item() for k, v in out.items()})
    ref = fx['losses']
    for it in range(fx['iters']):
        assert abs(history[it][rec_key] - ref[it][rec_key]) <= REC_TOL * abs(ref[it][rec_key]), (it, history[it], ref[it])
    assert abs(opt.Noise_Amps[-1] - fx['noise_amps_after'][-1]) <= REC_TOL * abs(fx['noise_amps_after'][-1])
    if d is not None:
        # the critic's losses are differences of near-equal means: compare on the scale of the penalty term
        scale = max(abs(ref[-1]['gradient_penalty']), 1e-3)
        assert abs(history[-1]['gradient_penalty'] - ref[-1]['gradient_penalty']) <= 0.1 * scale
    else:
        assert abs(history[-1]['kl_loss'] - ref[-1]['kl_loss']) <= 5e-2 * abs(ref[-1]['kl_loss'])


@pytest.mark.parametrize("name", ["train_sg_tiny", "train_csg_tiny", "train_sg_wide"])
def test_baselines_reconstruction_loss_after_k_iterations(golden, monkeypatch, name):
    """BASELINE configs[2]: the train_video_baselines.py loop (GeneratorSG / GeneratorCSG against WDiscriminator3D,
    train-depth 1) on the drop-in modules; reconstruction loss within 1 % of the reference after every iteration"""
    from hpvg import images, train
    from modules import networks_3d
    fx = golden(name)
    opt = train_opt_from(fx)
    g = getattr(networks_3d, fx['generator'])(opt)
    for _ in range(fx['stages']):
        g.init_next_stage()
    g.load_state_dict(state_from(fx), strict=True)
    g.cuda()
    d = networks_3d.WDiscriminator3D(opt)
    d.load_state_dict(state_d_from(fx), strict=True)
    d.cuda()
    tr = train.BaselineTrainer(opt, g, d)
    q = DrawQueue()
    monkeypatch.setattr(images, "draw_normal", q)
    alphas = []
    monkeypatch.setattr(torch, "rand", lambda *a, **k: torch.full((1, 1), alphas.pop(0)))
    real, z_init = fx['real'].cuda(), fx['z_init'].cuda()
    history = []
    for it in range(fx['iters']):
        dr = fx['draws'][it]
        q.tensors = [dr['noise_init']] + [dr['noises'][lvl] for lvl in sorted(dr['noises'])]
        alphas.append(dr['alpha'])
        out = tr.iteration(real, z_init)
        assert not q.tensors and not alphas, "the CUDA path draws in a different order than the reference"
        history.append({k: v.item() for k, v in out.items()})
    ref = fx['losses']
    for it in range(fx['iters']):
        assert abs(history[it]['rec_loss'] - ref[it]['rec_loss']) <= REC_TOL * abs(ref[it]['rec_loss']), (it, history[it], ref[it])
    assert abs(opt.Noise_Amps[-1] - fx['noise_amps_after'][-1]) <= REC_TOL * abs(fx['noise_amps_after'][-1])
    scale = max(abs(ref[-1]['gradient_penalty']), 1e-3)
    assert abs(history[-1]['gradient_penalty'] - ref[-1]['gradient_penalty']) <= 0.1 * scale


def test_recorded_iteration_replays_with_the_library_optimizer(golden):
    """ScaleTrainer.capture / replay: the whole GAN-level iteration, clipping and both Adam steps included,
    replays from one CUDA graph; every replay is a training step (device-side step count, moving weights, finite losses that
    stay on the scale of the eager iterations)"""
    from hpvg import optim, train
    from modules import networks_3d
    fx = golden("train_gan_tiny")
    opt = train_opt_from(fx)
    g = networks_3d.GeneratorHPVAEGAN(opt)
    for _ in range(fx['stages']):
        g.init_next_stage()
    g.load_state_dict(state_from(fx), strict=True)
    g.cuda()
    d = networks_3d.WDiscriminator3D(opt)
    d.load_state_dict(state_d_from(fx), strict=True)
    d.cuda()
    torch.manual_seed(0)
    tr = train.ScaleTrainer(opt, g, d, capturable=True)
    assert isinstance(tr.optimizerG, optim.Adam) and isinstance(tr.optimizerD, optim.Adam)
    real, real_zero = fx['real'].cuda(), fx['real_zero'].cuda()
    tr.capture(real, real_zero, warmup=2)
    tail = [p for p in g.body[-1].parameters()][0]
    before = tail.detach().clone()
    rec = []
    for _ in range(3):
        out = tr.replay()
        rec.append(out['rec_loss'].item())
        assert all(torch.isfinite(v).all() for v in out.values())
    assert not torch.equal(before, tail.detach())
    assert float(tr.optimizerG.state[tail]['step']) == 5.0 and float(tr.optimizerD.state[next(d.parameters())]['step']) == 5.0
    ref = fx['losses'][min(4, fx['iters'] - 1)]['rec_loss']
    assert 0.5 * ref < rec[-1] < 2.0 * ref, (rec, ref)
