"""Minimal stand-in for `kornia` 0.2 (absent from this image): the three functions the reference's datasets call
(datasets/video.py:50-78, datasets/image.py)."""
import numpy as np
import torch


def image_to_tensor(image):
    """HWC -> CHW, BHWC -> BCHW (kornia 0.2 semantics for numpy input)"""
    t = torch.from_numpy(np.ascontiguousarray(image))
    if t.dim() == 2:
        return t.unsqueeze(0)
    if t.dim() == 3:
        return t.permute(2, 0, 1)
    if t.dim() == 4:
        return t.permute(0, 3, 1, 2)
    raise ValueError("image_to_tensor: unsupported shape %s" % (tuple(t.shape),))


def hflip(t):
    return t.flip(-1)


def normalize(t, mean, std):
    return (t - mean) / std
