"""Drop-in for the reference's modules/networks_3d.py: same public names, constructor arguments, forward signatures,
return values and state_dict keys, so `getattr(networks_3d, opt.generator)(opt)` (train_video.py:396-397) and
`getattr(networks_3d, opt.discriminator)(opt)` (train_video.py:45) keep working — but every convolution, BatchNorm,
LeakyReLU, resize, tanh and the VAE head run in libhpvg.so (hand-written sm_100a kernels) via hpvg.blocks/hpvg.ops.
"""
import torch

from hpvg import blocks as _blocks
from hpvg import ops as _ops
from hpvg.blocks import weights_init  # noqa: F401  (reference modules/networks_3d.py:9-15)

_family = _blocks.make_family(3)


def _export(key, name):
    cls = _family[key]
    cls.__name__ = cls.__qualname__ = name
    cls.__module__ = __name__
    return cls


ConvBlock3D = _export('ConvBlock', 'ConvBlock3D')                      # reference :48-56
ConvBlock3DSN = _export('ConvBlockSN', 'ConvBlock3DSN')                # reference :59-70
FeatureExtractor = _export('FeatureExtractor', 'FeatureExtractor')     # reference :73-85
Encode3DVAE = _export('EncodeVAE', 'Encode3DVAE')                      # reference :88-107
WDiscriminator3D = _export('WDiscriminator', 'WDiscriminator3D')       # reference :163-181
GeneratorHPVAEGAN = _export('GeneratorHPVAEGAN', 'GeneratorHPVAEGAN')  # reference :325-406
GeneratorSG = _export('GeneratorSG', 'GeneratorSG')                    # reference :272-322
GeneratorCSG = _export('GeneratorCSG', 'GeneratorCSG')                 # reference :213-269
WDiscriminatorBaselines = _export('WDiscriminatorBaselines', 'WDiscriminatorBaselines')   # reference :184-210


def get_activation(act):
    """reference :18-26 — only the LeakyReLU(0.2) entry is used by the networks on the path"""
    table = {
        "relu": lambda: torch.nn.ReLU(inplace=True),
        "lrelu": lambda: torch.nn.LeakyReLU(0.2, inplace=True),
        "elu": lambda: torch.nn.ELU(alpha=1.0, inplace=True),
        "prelu": lambda: torch.nn.PReLU(num_parameters=1, init=0.25),
        "selu": lambda: torch.nn.SELU(inplace=True),
    }
    return table[act]()


def reparameterize(mu, logvar, training):
    """z = eps * exp(logvar / 2) + mu with eps ~ N(0, 1) drawn like the reference (:29-35); thin tensors in and out."""
    if not training:
        return torch.zeros_like(mu).normal_()
    eps = torch.zeros_like(logvar).normal_()
    five = _blocks.as5d
    z = _ops.Reparam.apply(_ops.ToWide.apply(five(mu).contiguous()), _ops.ToWide.apply(five(logvar).contiguous()), five(eps))
    return _ops.ToThin.apply(z).view_as(mu)


def reparameterize_bern(x, training):
    """reference :38-43"""
    return _family['reparameterize_bern'](x, training)


Encode3DVAE_nb = _export('EncodeVAE_nb', 'Encode3DVAE_nb')             # reference :110-138
Encode3DVAE1x1 = _export('EncodeVAE1x1', 'Encode3DVAE1x1')             # reference :141-160 (plain torch: 1x1x1 convolutions)
GeneratorVAE_nb = _export('GeneratorVAE_nb', 'GeneratorVAE_nb')        # reference :409-485
