"""The CPU oracle (oracle/port.py) against the golden fixtures recorded from the unmodified reference."""
import pytest
import torch
import torch.nn.functional as F

from helpers import opt_from, rel_err, state_from, with_grad
from oracle import port

TOL = 2e-5


def _check_grads(sd, golden_grads, tol):
    """per-parameter gradient check; gradients that are mathematically zero (conv bias in front of BatchNorm) are
    compared on an absolute scale tied to the largest gradient of the network"""
    def gnorm(g):
        return g['norm'] if isinstance(g, dict) else g.double().norm().item()
    floor = 1e-6 * (1.0 + max(gnorm(g) for g in golden_grads.values()))
    checked = 0
    for k, g in golden_grads.items():
        mine = sd[k].grad
        assert mine is not None, k
        if isinstance(g, dict):
            assert abs(mine.double().norm().item() - g['norm']) <= tol * g['norm'] + floor, k
            d = (mine.flatten()[:64].double() - g['head'].double()).norm().item()
            assert d <= 50 * tol * g['head'].double().norm().item() + floor, k
        else:
            d = (mine.double() - g.double()).norm().item()
            assert d <= tol * g.double().norm().item() + floor, (k, d)
        checked += 1
    assert checked > 0


@pytest.mark.parametrize("name", ["hp3d_tiny", "hp3d_tiny_vae", "hp2d_tiny", "hp3d_wide"])
def test_generator_matches_reference(golden, name):
    fx = golden(name)
    opt = opt_from(fx)
    sd = with_grad(state_from(fx))
    rec = fx['rec']
    gen, gen_vae, (mu, logvar) = port.generator(sd, opt, fx['real_zero'], fx['amps'], mode='rec', eps=rec['eps'])
    assert rel_err(gen, rec['gen']) < TOL
    assert rel_err(gen_vae, rec['gen_vae']) < TOL
    assert rel_err(mu, rec['mu']) < TOL and rel_err(logvar, rec['logvar']) < TOL
    kl = port.kl_criterion(mu, logvar)
    assert abs(kl.item() - rec['kl']) < TOL * max(1.0, abs(rec['kl']))
    loss = 10.0 * (F.mse_loss(gen, fx['real']) + F.mse_loss(gen_vae, fx['real_zero'])) + kl
    assert abs(loss.item() - rec['loss']) < TOL * abs(rec['loss'])
    loss.backward()
    _check_grads(sd, rec['grads'], 5e-4)
    for k, b in rec['buffers'].items():
        assert rel_err(sd[k].float(), b.float()) < TOL, k
    with torch.no_grad():
        fake, fake_vae = port.generator(sd, opt, None, fx['amps'], noise_init=fx['rand']['z'], mode='rand', noises=fx['rand']['noises'])
    assert rel_err(fake, fx['rand']['fake']) < TOL
    assert rel_err(fake_vae, fx['rand']['fake_vae']) < TOL
    for k, b in fx['rand']['buffers'].items():
        assert rel_err(sd[k].float(), b.float()) < TOL, k


@pytest.mark.parametrize("name", ["d3d_tiny", "d2d_tiny", "d3d_wide", "d2d_wide"])
def test_discriminator_and_gradient_penalty_match_reference(golden, name):
    fx = golden(name)
    opt = opt_from(fx)
    sd = with_grad(state_from(fx))
    out_real = port.discriminator(sd, opt, fx['real'])
    out_fake = port.discriminator(sd, opt, fx['fake'])
    assert rel_err(out_real, fx['out_real']) < TOL and rel_err(out_fake, fx['out_fake']) < TOL
    gp = port.gradient_penalty(sd, opt, fx['real'], fx['fake'], fx['lambda'], alpha=fx['alpha'])
    assert abs(gp.item() - fx['gp']) < 1e-4 * abs(fx['gp'])
    (-out_real.mean() + out_fake.mean() + gp).backward()
    _check_grads(sd, fx['grads'], 2e-3)
    for k, b in fx['buffers'].items():
        assert rel_err(sd[k], b) < TOL, k


def test_generator_sg_matches_reference(golden):
    fx = golden("sg3d_tiny")
    opt = opt_from(fx)
    sd = with_grad(state_from(fx))
    out = port.generator_sg(sd, opt, fx['z'], fx['amps'], mode='rec')
    assert rel_err(out, fx['rec']['out']) < TOL
    loss = F.mse_loss(out, fx['rec']['target'])
    assert abs(loss.item() - fx['rec']['loss']) < TOL
    loss.backward()
    _check_grads(sd, fx['rec']['grads'], 5e-4)
    with torch.no_grad():
        fake = port.generator_sg(sd, opt, fx['z'], fx['amps'], mode='rand', noises=fx['rand']['noises'])
    assert rel_err(fake, fx['rand']['fake']) < TOL


@pytest.mark.parametrize("name", ["csg3d_tiny", "csg3d_wide"])
def test_generator_csg_matches_reference(golden, name):
    """GeneratorCSG (networks_3d.py:213-269), the default generator of train_video_baselines.py"""
    fx = golden(name)
    opt = opt_from(fx)
    sd = with_grad(state_from(fx))
    out = port.generator_csg(sd, opt, fx['z'], fx['amps'], mode='rec')
    assert rel_err(out, fx['rec']['out']) < TOL
    loss = F.mse_loss(out, fx['rec']['target'])
    assert abs(loss.item() - fx['rec']['loss']) < TOL
    loss.backward()
    _check_grads(sd, fx['rec']['grads'], 5e-4)
    for k, b in fx['rec']['buffers'].items():
        assert rel_err(sd[k].float(), b.float()) < 1e-5, k
    with torch.no_grad():
        fake = port.generator_csg(sd, opt, fx['z'], fx['amps'], mode='rand', noises=fx['rand']['noises'])
    assert rel_err(fake, fx['rand']['fake']) < TOL


@pytest.mark.parametrize("name", ["dbase3d_tiny", "dbase3d_wide"])
def test_discriminator_baselines_matches_reference(golden, name):
    """WDiscriminatorBaselines (networks_3d.py:184-210): outputs and first-order gradients"""
    fx = golden(name)
    opt = opt_from(fx)
    sd = with_grad(state_from(fx))
    out_real = port.discriminator_baselines(sd, opt, fx['real'])
    out_fake = port.discriminator_baselines(sd, opt, fx['fake'])
    assert rel_err(out_real, fx['out_real']) < TOL and rel_err(out_fake, fx['out_fake']) < TOL
    (-out_real.mean() + out_fake.mean()).backward()
    _check_grads(sd, fx['grads'], 5e-4)
    for k, b in fx['buffers'].items():
        assert rel_err(sd[k].float(), b.float()) < 1e-5, k


def test_scale_schedule(golden):
    fx = golden("hp3d_tiny")
    opt = opt_from(fx)
    assert [(port.scale_size(i, opt), port.time_depth(i, opt)) for i in range(fx['stages'] + 1)] == fx['sizes']
    # the schedules quoted in SURVEY.md App. A
    o = port.Opt(img_size=64, sampling_rates=[5, 3, 1])
    assert [(port.scale_size(i, o), port.time_depth(i, o)) for i in range(5)] == [(32, 4), (39, 4), (46, 6), (54, 6), (64, 16)]
    o = port.Opt(img_size=64)
    assert [(port.scale_size(i, o), port.time_depth(i, o)) for i in range(5)] == [(32, 4), (39, 4), (46, 5), (54, 7), (64, 13)]


@pytest.mark.parametrize("name,bound", [("hp3d_tiny", 0.45), ("hp2d_tiny", 0.45), ("hp3d_wide", 0.10)])
def test_bf16_storage_emulation_stays_close_to_fp32(golden, name, bound):
    """oracle.port.storage('bf16') — the precision model of the CUDA path — against the fp32 reference fixtures: outputs
    within 3e-2, parameter gradients within the bound the GPU tests (tests/test_gpu_modules.py REF_GRAD_TOL) allow."""
    fx = golden(name)
    opt = opt_from(fx)
    sd = with_grad(state_from(fx))
    rec = fx['rec']
    with port.storage('bf16'):
        gen, gen_vae, (mu, logvar) = port.generator(sd, opt, fx['real_zero'], fx['amps'], mode='rec', eps=rec['eps'])
        loss = 10.0 * (F.mse_loss(gen, fx['real']) + F.mse_loss(gen_vae, fx['real_zero'])) + port.kl_criterion(mu, logvar)
        loss.backward()
    assert rel_err(gen, rec['gen']) < 3e-2 and rel_err(gen_vae, rec['gen_vae']) < 3e-2
    assert abs(loss.item() - rec['loss']) < 2e-3 * abs(rec['loss'])
    big = max((g['norm'] if isinstance(g, dict) else g.double().norm().item()) for g in rec['grads'].values())
    for k, g in rec['grads'].items():
        mine = sd[k].grad.double()
        if isinstance(g, dict):
            assert abs(mine.norm().item() - g['norm']) <= bound * g['norm'] + 2e-3 * big, k
        else:
            assert (mine - g.double()).norm().item() <= bound * g.double().norm().item() + 2e-3 * big, k


@pytest.mark.parametrize("name", ["train_vae_tiny", "train_gan_tiny", "train_gan_wide", "train_vae2d_tiny", "train_gan2d_tiny"])
def test_training_loop_matches_reference(golden, name):
    """oracle/train_ref.py (optimizer groups + iteration body of train_video.py:44-202) against the losses recorded from
    the unmodified reference modules stepping the same loop on the same draws"""
    from helpers import state_d_from, train_opt_from
    from oracle import train_ref
    fx = golden(name)
    opt = train_opt_from(fx)
    sd_g = state_from(fx)
    sd_d = state_d_from(fx) if 'state_d' in fx else None
    tr = train_ref.ScaleTrainer(opt, sd_g, sd_d)
    for it in range(fx['iters']):
        dr = fx['draws'][it]
        out = tr.iteration(fx['real'], fx['real_zero'], noise_init=dr['noise_init'], eps=dr['eps'], noises=dr.get('noises'),
                           alpha=dr.get('alpha'), eps_amp=dr.get('eps_amp'))
        for k, v in fx['losses'][it].items():
            assert abs(out[k].item() - v) <= 2e-3 * abs(v) + 2e-5, (it, k, out[k].item(), v)
    assert len(opt.Noise_Amps) == len(fx['noise_amps_after'])
    assert abs(opt.Noise_Amps[-1] - fx['noise_amps_after'][-1]) < 1e-4 * abs(fx['noise_amps_after'][-1])
    key = ('body.%d.tail.weight' % (fx['stages'] - 1)) if fx['stages'] > 0 else 'decoder.tail.weight'
    assert rel_err(sd_g[key], fx['final_tail_weight']) < 1e-3


@pytest.mark.parametrize("name", ["train_sg_tiny", "train_csg_tiny", "train_sg_wide"])
def test_baselines_training_loop_matches_reference(golden, name):
    """oracle/train_ref.py::BaselineTrainer (train_video_baselines.py:44-70, :100-173; BASELINE configs[2]: GeneratorSG,
    train-depth 1) against the losses recorded from the unmodified reference modules stepping the same loop"""
    from helpers import state_d_from, train_opt_from
    from oracle import train_ref
    fx = golden(name)
    opt = train_opt_from(fx)
    sd_g, sd_d = state_from(fx), state_d_from(fx)
    tr = train_ref.BaselineTrainer(opt, sd_g, sd_d, generator=fx['generator'])
    for it in range(fx['iters']):
        dr = fx['draws'][it]
        out = tr.iteration(fx['real'], fx['z_init'], noise_init=dr['noise_init'], noises=dr['noises'], alpha=dr['alpha'])
        for k, v in fx['losses'][it].items():
            assert abs(out[k].item() - v) <= 2e-3 * abs(v) + 2e-5, (it, k, out[k].item(), v)
    assert abs(opt.Noise_Amps[-1] - fx['noise_amps_after'][-1]) < 1e-4 * abs(fx['noise_amps_after'][-1])
    assert rel_err(sd_g[fx['final_key']], fx['final_weight']) < 1e-3


def test_numpy_primitives_match_torch():
    """oracle/np_ops.py (independent float64 numpy definitions) against the torch CPU operators oracle/port.py calls"""
    import numpy as np
    from oracle import np_ops
    x = port.det_tensor((2, 3, 4, 6, 5), 61).double()
    w = port.det_tensor((5, 3, 3, 3, 3), 62, scale=0.3).double()
    b = port.det_tensor((5,), 63).double()
    for pad in (0, 1, 2):
        assert np.allclose(np_ops.conv_nd(x.numpy(), w.numpy(), b.numpy(), pad), F.conv3d(x, w, b, padding=pad).numpy(), atol=1e-10)
    x2, w2 = x[:, :, 0], w[:, :, 0]
    assert np.allclose(np_ops.conv_nd(x2.numpy(), w2.numpy(), None, 1), F.conv2d(x2, w2, None, padding=1).numpy(), atol=1e-10)
    y = F.conv3d(x, w, b, padding=1)
    gamma, beta = port.det_tensor((5,), 64).double() + 1.5, port.det_tensor((5,), 65).double()
    sd = {'weight': gamma, 'bias': beta, 'running_mean': torch.zeros(5).double(), 'running_var': torch.ones(5).double(),
          'num_batches_tracked': torch.zeros((), dtype=torch.int64)}
    out_t = port.batch_norm_train(sd, '', y)
    out_n, mean, var_unbiased = np_ops.batch_norm_train(y.numpy(), gamma.numpy(), beta.numpy())
    assert np.allclose(out_n, out_t.numpy(), atol=1e-9)
    assert np.allclose(sd['running_mean'].numpy(), 0.1 * mean, atol=1e-12)
    assert np.allclose(sd['running_var'].numpy(), 0.9 + 0.1 * var_unbiased, atol=1e-12)
    assert np.allclose(np_ops.leaky_relu(y.numpy()), F.leaky_relu(y, 0.2).numpy())
    for size in ((4, 8, 7), (9, 3, 11), (1, 6, 5)):
        assert np.allclose(np_ops.resize_linear(x.numpy(), size), port.resize(x, size).numpy(), atol=1e-12)
    assert np.allclose(np_ops.resize_linear(x2.numpy(), (9, 4)), port.resize(x2, (9, 4)).numpy(), atol=1e-12)
    u = port.det_tensor((5,), 66).double()
    u = u / u.norm()
    sdw = {'weight_orig': w.clone(), 'weight_u': u.clone(), 'weight_v': torch.zeros(81).double()}
    w_sn = port.spectral_weight(sdw, '')
    w_np, u_np, v_np = np_ops.spectral_norm_step(w.numpy(), u.numpy())
    assert np.allclose(w_np, w_sn.numpy(), atol=1e-10) and np.allclose(u_np, sdw['weight_u'].numpy(), atol=1e-10)
    assert np.allclose(v_np, sdw['weight_v'].numpy(), atol=1e-10)
    mu, lv = port.det_tensor((2, 4, 3, 3), 67).double(), port.det_tensor((2, 4, 3, 3), 68).double()
    assert abs(np_ops.kl(mu.numpy(), lv.numpy()) - port.kl_criterion(mu, lv).item()) < 1e-12
    g = port.det_tensor((2, 3, 2, 4, 4), 69).double()
    assert abs(np_ops.gp_penalty(g.numpy(), 0.1) - (((g.norm(2, dim=1) - 1) ** 2).mean() * 0.1).item()) < 1e-12


def test_data_path_restatement():
    """oracle/data_ref.py: the reference's per-iteration clip construction (datasets/video.py:44-82) and the uint8 frame
    conversion of utils/saver.py:16-18 on hand-checkable values"""
    import numpy as np
    from oracle import data_ref
    frames = np.zeros((9, 2, 3, 3), dtype=np.uint8)
    for f in range(9):
        frames[f] = f * 25
    frames[:, 0, 0, 1] = 255
    clip = data_ref.clip_from_frames(frames, 1, 6, 3)          # frames 1, 4, 7
    assert tuple(clip.shape) == (3, 3, 2, 3) and clip.dtype == torch.float32
    assert clip[1, :, 0, 0].tolist() == [1.0, 1.0, 1.0]
    expect = [(np.float32(v) / np.float32(255) - np.float32(0.5)) / np.float32(0.5) for v in (25, 100, 175)]
    assert clip[0, :, 1, 2].tolist() == [float(e) for e in expect]
    flipped = data_ref.clip_from_frames(frames, 1, 6, 3, hflip=True)
    assert torch.equal(flipped, clip.flip(-1))
    video = np.array([-1.0, -0.5, 0.0, 0.5, 1.0, 0.999], dtype=np.float32).reshape(1, 1, 1, 6).repeat(3, 0)
    assert data_ref.frames_to_uint8(video)[0, 0, :, 0].tolist() == [0, 63, 127, 191, 255, 254]


def test_numpy_clip_and_adam_match_torch():
    """oracle/np_ops.py clip_grad_norm / adam_step (the definition hpvg.optim.Adam's kernels are held to on the GPU, SURVEY.md
    §8f-1) against torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the CPU, in float64, over several steps"""
    import numpy as np
    from oracle import np_ops
    gen = torch.Generator().manual_seed(0)
    shapes = [(4, 3, 3, 3, 3), (4,), (7, 5)]
    params = [torch.nn.Parameter(torch.randn(s, generator=gen, dtype=torch.float64)) for s in shapes]
    opt = torch.optim.Adam([{"params": params[:1], "lr": 1e-4}, {"params": params[1:], "lr": 5e-4}], lr=5e-4, betas=(0.5, 0.999))
    lrs = [1e-4, 5e-4, 5e-4]
    p_np = [p.detach().numpy().copy() for p in params]
    m_np = [np.zeros_like(a) for a in p_np]
    v_np = [np.zeros_like(a) for a in p_np]
    for step in range(5):
        grads = [torch.randn(s, generator=gen, dtype=torch.float64) * 10.0 ** (step - 2) for s in shapes]
        for p, g in zip(params, grads):
            p.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_(params, 5.0)
        scaled, total_np = np_ops.clip_grad_norm([g.numpy() for g in grads], 5.0)
        assert abs(total_np - total.item()) <= 1e-12 * total.item()
        for p, s_ in zip(params, scaled):
            assert np.allclose(p.grad.numpy(), s_, rtol=1e-12, atol=0)
        opt.step()
        for i in range(len(params)):
            p_np[i], m_np[i], v_np[i] = np_ops.adam_step(p_np[i], scaled[i], m_np[i], v_np[i], step, lrs[i], 0.5)
            assert np.allclose(params[i].detach().numpy(), p_np[i], rtol=1e-10, atol=1e-14)
            assert np.allclose(opt.state[params[i]]['exp_avg'].numpy(), m_np[i], rtol=1e-10, atol=1e-300)
            assert np.allclose(opt.state[params[i]]['exp_avg_sq'].numpy(), v_np[i], rtol=1e-10, atol=1e-300)


def test_data_formats_match_the_reference_dataset_and_writer(golden):
    """rows f2 / f3: oracle/data_ref.py against what the UNMODIFIED reference produced (tests/golden/make_data_golden.py):
    SingleVideoDataset.__getitem__ (datasets/video.py:44-66) at every pyramid level, two indices, with and without the
    horizontal flip, and the uint8 frames utils/saver.py::write_video hands to the encoder — bit for bit"""
    import numpy as np
    from oracle import data_ref
    fx = golden("data_video")
    rates, lcm = fx['sampling_rates'], fx['fps_lcm']
    assert any(c['hflip'] for c in fx['cases']) and not all(c['hflip'] for c in fx['cases'])
    for c in fx['cases']:
        frames = fx['levels'][c['scale']].numpy()
        real = data_ref.clip_from_frames(frames, c['idx'], lcm, rates[c['fps_index']], c['hflip'])
        assert torch.equal(real, c['real']), (c['scale'], c['idx'], c['hflip'])
        if c['scale'] > 0:
            zero = data_ref.clip_from_frames(fx['zero'].numpy(), c['idx'], lcm, rates[0], c['hflip'])
            assert torch.equal(zero, c['real_zero'])
    out = data_ref.frames_to_uint8(fx['video'].numpy())
    assert np.array_equal(out, fx['written'].numpy())
