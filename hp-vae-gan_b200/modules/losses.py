"""Drop-in for the reference's modules/losses.py."""
import math

import torch

from hpvg import ops as _ops

__all__ = ['kl_criterion', 'kl_bern_criterion']


def kl_criterion(mu, logvar):
    """KL(N(mu, exp(logvar)) || N(0, 1)) averaged over every element (reference modules/losses.py:7-9): one fused
    reduction kernel forward, one elementwise kernel backward."""
    return _ops.KlCriterion.apply(mu, logvar)


def kl_bern_criterion(x):
    """Bernoulli KL against p = 0.5 (reference :12-14) — used only by the *_nb variants, which are out of scope; plain torch."""
    tiny = 1e-20
    pos = x * (torch.log(x + tiny) - math.log(0.5))
    neg = (1 - x) * (torch.log(1 - x + tiny) - math.log(0.5))
    return (pos + neg).mean()
