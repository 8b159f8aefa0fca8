"""Drop-in for the reference's modules/networks_2d.py (train_image.py:21, :101): the 2-D twins of the video networks.
Same kernels as the 3-D family with D == 1 and 3x3 weights; bilinear resize; noise at every refinement level in
'rand' mode (reference modules/networks_2d.py:261-263).
"""
import torch

from hpvg import blocks as _blocks
from hpvg import ops as _ops
from hpvg.blocks import weights_init  # noqa: F401

_family = _blocks.make_family(2)


def _export(key, name):
    cls = _family[key]
    cls.__name__ = cls.__qualname__ = name
    cls.__module__ = __name__
    return cls


ConvBlock2D = _export('ConvBlock', 'ConvBlock2D')                      # reference :53-61
ConvBlock2DSN = _export('ConvBlockSN', 'ConvBlock2DSN')                # reference :64-75
FeatureExtractor = _export('FeatureExtractor', 'FeatureExtractor')     # reference :78-90
Encode2DVAE = _export('EncodeVAE', 'Encode2DVAE')                      # reference :93-112
WDiscriminator2D = _export('WDiscriminator', 'WDiscriminator2D')       # reference :168-185
GeneratorHPVAEGAN = _export('GeneratorHPVAEGAN', 'GeneratorHPVAEGAN')  # reference :188-269


def reparameterize(mu, logvar, training):
    """reference :36-42"""
    if not training:
        return torch.zeros_like(mu).normal_()
    eps = torch.zeros_like(logvar).normal_()
    five = _blocks.as5d
    z = _ops.Reparam.apply(_ops.ToWide.apply(five(mu).contiguous()), _ops.ToWide.apply(five(logvar).contiguous()), five(eps))
    return _ops.ToThin.apply(z).view_as(mu)


def reparameterize_bern(x, training):
    """reference :45-50"""
    return _family['reparameterize_bern'](x, training)


Encode2DVAE_nb = _export('EncodeVAE_nb', 'Encode2DVAE_nb')             # reference :115-143
GeneratorVAE_nb = _export('GeneratorVAE_nb', 'GeneratorVAE_nb')        # reference :272-348
