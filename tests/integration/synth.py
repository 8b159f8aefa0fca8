"""Synthetic inputs for the script-level runs: a smooth moving pattern written as an MJPG .avi / a PNG."""
import numpy as np


def _pattern(t, h, w):
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    r = 127.5 + 127.5 * np.sin(0.21 * xx + 0.4 * t) * np.cos(0.17 * yy - 0.3 * t)
    g = 127.5 + 127.5 * np.sin(0.11 * (xx + yy) + 0.25 * t)
    b = 127.5 + 127.5 * np.cos(0.19 * xx - 0.13 * yy + 0.5 * t)
    return np.stack([b, g, r], -1).clip(0, 255).astype(np.uint8)      # BGR for cv2


def write_video(path, frames=16, size=64):
    import cv2
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc('M', 'J', 'P', 'G'), 24.0, (size, size))
    if not vw.isOpened():
        raise RuntimeError("cv2.VideoWriter could not open %s" % path)
    for t in range(frames):
        vw.write(_pattern(t, size, size))
    vw.release()
    return path


def write_image(path, size=128):
    import cv2
    cv2.imwrite(path, _pattern(3, size, size))
    return path
