"""CPU restatement of the data formats either side of the path (TEST INFRASTRUCTURE — see oracle/__init__.py).

clip_from_frames follows SingleVideoDataset.__getitem__ / _get_transformed_frames (datasets/video.py:44-82); kornia
(pinned kornia==0.2.0 in env.sh:5, absent here) is restated from its documented semantics: image_to_tensor = HWC -> CHW
(batched: BHWC -> BCHW) without scaling, hflip = flip of the last axis, normalize(x, mean, std) = (x - mean) / std.
frames_to_uint8 follows utils/saver.py::write_video (:16-18).  Pinned by tests/golden/data_video.pt, recorded from the unmodified
reference's SingleVideoDataset.__getitem__ and write_video by tests/golden/make_data_golden.py (kornia through the shim of
tests/integration/shims): tests/test_oracle.py::test_data_formats_match_the_reference_dataset_and_writer holds both functions to
those vectors bit for bit.
"""
import numpy as np
import torch


def clip_from_frames(frames, idx, fps_lcm, every, hflip=False):
    """frames: uint8 ndarray [F, H, W, 3] -> float32 tensor [3, T, H, W]"""
    sel = frames[idx:idx + fps_lcm + 1:every]                      # datasets/video.py:52
    t = torch.from_numpy(np.ascontiguousarray(sel)).permute(0, 3, 1, 2)   # K.image_to_tensor: T,H,W,C -> T,C,H,W
    t = t / 255                                                      # :54
    if hflip:
        t = t.flip(-1)                                               # K.hflip (:75)
    t = (t - 0.5) / 0.5                                              # K.normalize(x, 0.5, 0.5) (:78)
    return t.permute(1, 0, 2, 3).contiguous()                        # CTHW (:81)


def frames_to_uint8(video):
    """video: float32 ndarray [3, T, H, W] -> uint8 ndarray [T, H, W, 3] (utils/saver.py:16-18)"""
    out = []
    for i in range(video.shape[1]):
        frame = (video[:, i, :, :] + 1) * 127.5
        frame = frame.transpose(1, 2, 0)
        out.append(np.uint8(frame))
    return np.stack(out)
