"""Stand-in for `matplotlib` (absent): train_image.py only calls matplotlib.use('Agg')."""


def use(*a, **k):
    pass
