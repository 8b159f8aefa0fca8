"""Scale-pyramid helpers the generators call inside forward.

Restates the reference's utils/images.py for the functions on the hot path: the per-scale size / frame-count
schedule (get_scales_by_index :60-64, get_fps_by_index :67-71, get_fps_td_by_index :74-80), the resize wrappers
(interpolate :9-19, interpolate_3D :22-26, upscale :83-93, upscale_2d :96-105) and generate_noise (:39-57).
Resizing runs in libhpvg (hpvg_upsample_linear_*); noise is drawn with torch's generator in the reference's
order and shape so that equal seeds give equal draws.
"""
import math

import torch

from . import ops


def scale_size(index, scale_factor, stop_scale, img_size):
    """spatial size (width) of pyramid level `index`"""
    return math.ceil(math.pow(scale_factor, stop_scale - index) * img_size)


def frames_at(index, opt):
    """-> (fps, time_depth, fps_index) of pyramid level `index`"""
    fps_index = int((index / opt.stop_scale_time) * (len(opt.sampling_rates) - 1))
    every = opt.sampling_rates[fps_index]
    return opt.org_fps / every, opt.fps_lcm // every + 1, fps_index


def resize(x, size, noise=None, amp=0.0):
    """align_corners=True linear resize of a thin tensor [N,C,D,H,W] to `size` = (D,H,W), optionally + amp*noise"""
    return ops.UpsampleLinear.apply(x, tuple(int(s) for s in size), noise, float(amp))


def video_target_size(index, opt):
    if index <= 0:
        raise AssertionError("upscale is defined for levels > 0")
    s = scale_size(index, opt.scale_factor, opt.stop_scale, opt.img_size)
    _, td, _ = frames_at(index, opt)
    return [td, int(s * opt.ar), s]


def image_target_size(index, opt):
    if index <= 0:
        raise AssertionError("upscale is defined for levels > 0")
    s = scale_size(index, opt.scale_factor, opt.stop_scale, opt.img_size)
    return [int(s * opt.ar), s]


def upscale(video, index, opt, noise=None, amp=0.0):
    """5-D video [N,C,T,H,W] -> level `index` size (trilinear, time axis included)"""
    return resize(video, video_target_size(index, opt), noise, amp)


def upscale_2d(image, index, opt, noise=None, amp=0.0):
    """4-D image [N,C,H,W] -> level `index` size (bilinear)"""
    h, w = image_target_size(index, opt)
    n5 = None if noise is None else noise.unsqueeze(2)
    return resize(image.unsqueeze(2), (1, h, w), n5, amp).squeeze(2)


def interpolate_3D(video, size):
    if video.dim() != 5:
        raise AssertionError("input must be 5D")
    return resize(video, size)


def draw_normal(shape, dtype, device):
    """the single place where the networks draw N(0,1) numbers: zeros(shape).normal_(0, 1) on `device`, i.e. the same
    generator consumption as the reference's zeros_like(ref).normal_() calls (tests replace this hook to inject noise)"""
    return torch.zeros(tuple(shape), dtype=dtype, device=device).normal_(0, 1)


def generate_noise(ref=None, size=None, device=None):
    """N(0,1) noise shaped like `ref` (or `size`), reference utils/images.py:39-49"""
    if ref is not None:
        return draw_normal(ref.shape, ref.dtype, ref.device)
    if size is not None:
        return draw_normal(size, torch.float32, device)
    raise Exception("ref or size must be applied")
