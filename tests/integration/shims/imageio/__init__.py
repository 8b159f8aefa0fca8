"""Stand-in for `imageio` (absent): imread through PIL, RGB uint8 ndarray (datasets/image.py)."""
import numpy as np
from PIL import Image


def imread(path, *a, **k):
    return np.asarray(Image.open(path).convert("RGB"))
