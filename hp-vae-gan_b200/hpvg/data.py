"""Data formats on either side of the path, on the device.

The reference builds every training clip on the host, per iteration, in DataLoader workers: slice the frames of the current
pyramid level at the level's sampling rate, /255, optional horizontal flip, normalize to [-1, 1], permute to CTHW
(datasets/video.py:44-92), then copies it to the GPU (train_video.py:120-123).  ResidentVideo keeps the uint8 frames of the
level (and of level 0) in HBM and produces the same tensors with one kernel each — bit-exact.  The per-level cv2 resize
(datasets/generate_frames.py:44-46) happens once per scale on the host and stays there: pass its result in.
"""
import torch

from . import lib, ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


class ResidentVideo:
    def __init__(self, frames, zero_scale_frames, opt, device):
        """frames / zero_scale_frames: uint8 [F, H, W, 3] RGB arrays or tensors as returned by the reference's
        dataset._generate_frames(scale_idx) / dataset.zero_scale_frames"""
        self.opt = opt
        self.frames = torch.as_tensor(frames).to(device=device, dtype=torch.uint8).contiguous()
        self.zero = torch.as_tensor(zero_scale_frames).to(device=device, dtype=torch.uint8).contiguous()
        if self.frames.dim() != 4 or self.frames.shape[-1] != 3 or self.zero.dim() != 4 or self.zero.shape[-1] != 3:
            raise ValueError("frames must be uint8 [F, H, W, 3]")

    def __len__(self):
        """datasets/video.py:41-42"""
        return (self.zero.shape[0] - self.opt.fps_lcm) * getattr(self.opt, 'data_rep', 1)

    def _clip(self, frames, first, every, hflip):
        f, h, w, _ = frames.shape
        t = len(range(first, first + self.opt.fps_lcm + 1, every))
        out = torch.empty((3, t, h, w), dtype=torch.float32, device=frames.device)
        lib.call("hpvg_clip_from_frames", frames.data_ptr(), out.data_ptr(), f, first, every, t, h, w, int(bool(hflip)), _stream())
        return out

    def clip(self, idx, hflip=False):
        """-> (real [3,T,H,W], real_zero [3,T0,H0,W0]) for dataset index idx, as SingleVideoDataset.__getitem__ returns them
        (at scale 0 both are the same tensor's content)"""
        idx = idx % (self.zero.shape[0] - self.opt.fps_lcm)
        every = self.opt.sampling_rates[self.opt.fps_index]
        real = self._clip(self.frames, idx, every, hflip)
        real_zero = self._clip(self.zero, idx, self.opt.sampling_rates[0], hflip)
        return real, real_zero


def to_uint8_frames(video, out=None):
    """[3, T, H, W] float32 in [-1, 1] -> uint8 [T, H, W, 3] as utils/saver.py::write_video converts frames before encoding;
    a batch [N, 3, T, H, W] -> [N, T, H, W, 3] in one launch.  `out`: optional preallocated result."""
    ops._require_cuda(video)
    video = video.contiguous()
    batched = video.dim() == 5
    if video.dim() not in (4, 5) or video.shape[-4] != 3 or video.dtype != torch.float32:
        raise ValueError("expected a float32 [3, T, H, W] video or a [N, 3, T, H, W] batch")
    n = video.shape[0] if batched else 1
    t, h, w = video.shape[-3:]
    shape = (n, t, h, w, 3) if batched else (t, h, w, 3)
    if out is None:
        out = torch.empty(shape, dtype=torch.uint8, device=video.device)
    elif tuple(out.shape) != shape or out.dtype != torch.uint8 or not out.is_contiguous():
        raise ValueError("out must be a contiguous uint8 tensor of shape %s" % (shape,))
    lib.call("hpvg_frames_to_uint8_batched", video.data_ptr(), out.data_ptr(), n, t, h, w, _stream())
    return out
