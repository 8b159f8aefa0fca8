"""Golden vectors for the data formats either side of the path, recorded from the UNMODIFIED reference:

  * datasets/video.py::SingleVideoDataset.__getitem__ (:44-66) on a synthetic MJPG clip, at several pyramid levels, dataset
    indices and with / without the horizontal flip — the decoded uint8 frames the dataset holds (self.frames,
    self.zero_scale_frames) go into the fixture next to the tensors it returns, so the fixture does not depend on cv2's decoder;
  * utils/saver.py::write_video (:8-19): the uint8 frames it hands to cv2.VideoWriter for a float video in [-1, 1].

kornia (pinned 0.2.0 in env.sh:5) is absent from this image; tests/integration/shims/kornia provides the three functions the
dataset calls.  Run in the build container (needs /root/reference):  python tests/golden/make_data_golden.py
"""
import os
import random
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(ROOT, "tests", "integration", "shims"), "/root/reference", os.path.join(ROOT, "tests", "integration")]

import cv2  # noqa: E402
import synth  # noqa: E402
import utils  # noqa: E402  (the reference's package)
from datasets import SingleVideoDataset  # noqa: E402
from utils import saver  # noqa: E402


def main():
    tmp = tempfile.mkdtemp()
    video = synth.write_video(os.path.join(tmp, "syn.avi"), frames=14, size=36)
    opt = types.SimpleNamespace(video_path=video, sampling_rates=[4, 3, 2, 1], img_size=28, min_size=18, max_size=256, scale_factor=0.75,
                                start_frame=0, max_frames=1000, hflip=True, data_rep=1, scale_factor_init=0.75)
    utils.adjust_scales2image(opt.img_size, opt)
    opt.scale_idx = 0
    ds = SingleVideoDataset(opt)
    opt.stop_scale_time = opt.stop_scale
    cases, levels = [], {}
    for scale in range(opt.stop_scale + 1):
        opt.scale_idx = scale
        _, _, opt.fps_index = utils.get_fps_td_by_index(scale, opt)
        ds.generate_frames(scale)
        levels[scale] = torch.from_numpy(ds.frames.copy())
        for idx in (0, 1):
            for seed in (0, 1):
                random.seed(seed)                  # __getitem__ flips when random.random() < 0.5
                state = random.getstate()
                flipped = random.random() < 0.5
                random.setstate(state)
                item = ds[idx]
                real, real_zero = (item if isinstance(item, list) else (item, item))
                cases.append(dict(scale=scale, idx=idx, hflip=bool(flipped), fps_index=int(opt.fps_index), real=real.clone(),
                                  real_zero=real_zero.clone()))
    # write_video: capture what reaches cv2.VideoWriter.write
    written = []

    class FakeWriter(object):
        def __init__(self, *a, **k):
            pass

        def write(self, frame):
            written.append(np.array(frame, copy=True))

        def release(self):
            pass

    real_writer = cv2.VideoWriter
    cv2.VideoWriter = FakeWriter
    try:
        gen = torch.Generator().manual_seed(7)
        vid = torch.tanh(torch.randn((3, 9, 21, 19), generator=gen) * 2.0)
        vid[:, 0, 0, :8] = torch.tensor([-1.0, 1.0, 0.0, 0.999999, -0.999999, 0.5, -0.5, 1e-8])
        saver.write_video(vid.numpy(), os.path.join(tmp, "out.avi"), types.SimpleNamespace(fps=24.0))
    finally:
        cv2.VideoWriter = real_writer
    fx = dict(cases=cases, levels=levels, zero=torch.from_numpy(ds.zero_scale_frames.copy()), fps_lcm=int(opt.fps_lcm), sampling_rates=list(opt.sampling_rates), video=vid, written=torch.from_numpy(np.stack(written)),
              note="recorded from the unmodified reference (datasets/video.py, utils/saver.py) with the kornia shim of tests/integration/shims")
    out = os.path.join(HERE, "data_video.pt")
    torch.save(fx, out)
    print("wrote", out, os.path.getsize(out), "bytes,", len(cases), "dataset cases, flips:", sum(c['hflip'] for c in cases))


if __name__ == "__main__":
    main()
