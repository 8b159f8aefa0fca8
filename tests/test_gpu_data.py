"""Bit-exact parity of the on-device data path and output path (hpvg.data) with the oracle's restatement of the
reference's host code (datasets/video.py:44-92, utils/saver.py:8-19)."""
import numpy as np
import pytest
import torch

from oracle import data_ref, port

pytestmark = pytest.mark.gpu


def _frames(f, h, w, seed):
    return np.random.default_rng(seed).integers(0, 256, size=(f, h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("rates,level,hflip", [([4, 3, 2, 1], 0, False), ([4, 3, 2, 1], 3, True), ([5, 3, 1], 4, False),
                                                 ([5, 3, 1], 2, True)])
def test_clip_from_resident_frames_is_bit_exact(rates, level, hflip):
    from hpvg import data
    opt = port.Opt(img_size=64, sampling_rates=rates)
    opt.fps_index = int((level / opt.stop_scale_time) * (len(rates) - 1))      # utils.get_fps_td_by_index (utils/images.py:74-80)
    size, size0 = port.scale_size(level, opt), port.scale_size(0, opt)
    nframes = opt.fps_lcm + 7
    frames, zero = _frames(nframes, size, size + 3, 1), _frames(nframes, size0, size0 + 2, 2)
    rv = data.ResidentVideo(frames, zero, opt, 'cuda')
    assert len(rv) == nframes - opt.fps_lcm
    for idx in (0, 3, len(rv) - 1, len(rv) + 2):
        real, real_zero = rv.clip(idx, hflip)
        i = idx % (nframes - opt.fps_lcm)
        ref = data_ref.clip_from_frames(frames, i, opt.fps_lcm, rates[opt.fps_index], hflip)
        ref0 = data_ref.clip_from_frames(zero, i, opt.fps_lcm, rates[0], hflip)
        assert real.shape == ref.shape and real_zero.shape == ref0.shape
        assert torch.equal(real.cpu(), ref) and torch.equal(real_zero.cpu(), ref0)
        assert real.shape[1] == port.time_depth(level, opt)


def test_frames_to_uint8_is_bit_exact():
    from hpvg import data
    gen = torch.Generator().manual_seed(3)
    video = torch.tanh(torch.randn((3, 13, 37, 41), generator=gen) * 2.0)
    video[:, 0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, 0.999999, -0.999999, 0.5, -0.5, 1e-8])
    out = data.to_uint8_frames(video.cuda())
    ref = data_ref.frames_to_uint8(video.numpy())
    assert out.shape == ref.shape and out.dtype == torch.uint8
    assert np.array_equal(out.cpu().numpy(), ref)


def test_device_data_path_against_the_reference_dataset_and_writer(golden):
    """hpvg.data.ResidentVideo / to_uint8_frames against the vectors recorded from the UNMODIFIED reference dataset and writer
    (tests/golden/data_video.pt): bit for bit, every pyramid level, both flip states; the batched conversion too"""
    from hpvg import data
    fx = golden("data_video")
    opt = port.Opt(img_size=28, min_size=18, sampling_rates=fx['sampling_rates'])
    assert opt.fps_lcm == fx['fps_lcm']
    for c in fx['cases']:
        opt.fps_index = c['fps_index']
        rv = data.ResidentVideo(fx['levels'][c['scale']], fx['zero'], opt, 'cuda')
        real, real_zero = rv.clip(c['idx'], hflip=c['hflip'])
        assert torch.equal(real.cpu(), c['real']), (c['scale'], c['idx'], c['hflip'])
        if c['scale'] > 0:
            assert torch.equal(real_zero.cpu(), c['real_zero'])
    out = data.to_uint8_frames(fx['video'].cuda())
    assert torch.equal(out.cpu(), fx['written'])
    batch = torch.stack([fx['video'], -fx['video'], fx['video'] * 0.5]).cuda()
    outs = data.to_uint8_frames(batch)
    for k in range(3):
        assert torch.equal(outs[k], data.to_uint8_frames(batch[k]))


def test_resident_video_rejects_out_of_range_slices():
    from hpvg import data, lib
    opt = port.Opt(img_size=64, sampling_rates=[4, 3, 2, 1])
    opt.fps_index = 0
    rv = data.ResidentVideo(_frames(14, 8, 8, 4), _frames(14, 8, 8, 5), opt, 'cuda')
    with pytest.raises(lib.HpvgError):
        rv._clip(rv.frames, 5, 4, False)       # 5 + 3*4 = 17 > 13
