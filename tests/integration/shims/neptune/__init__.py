"""Stand-in for `neptune` (absent): train_image.py wraps neptune.init in try/except and falls back to TensorBoard."""


def init(*a, **k):
    raise RuntimeError("neptune is not available in this environment")
