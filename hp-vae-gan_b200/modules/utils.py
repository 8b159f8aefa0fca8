"""Drop-in for the reference's modules/utils.py.  The train scripts do `from modules.utils import *` and rely on it
re-exporting `torch` (train_video.py:16), so this module imports torch at top level and defines no __all__."""
import torch

from hpvg import ops as _ops


def calc_gradient_penalty(netD, real_data, fake_data, LAMBDA, device):
    """WGAN-GP penalty (reference modules/utils.py:4-19): one scalar alpha from the CPU generator, interpolates as a
    detached leaf, critic gradient w.r.t. them with create_graph=True, channel-axis L2 norm per voxel.
    The first-order sweep asks only for d/d(interpolates); the double backward is composed of the same
    fprop/dgrad/wgrad kernels by hpvg.ops."""
    alpha = _ops.gp_alpha(real_data.device)         # same CPU-generator draw as the reference (:5), as a device float
    interpolates = _ops.lerp(real_data, fake_data, alpha).requires_grad_(True)
    with _ops.twice_differentiable():      # critics with BatchNorm (WDiscriminatorBaselines): see hpvg.ops.twice_differentiable
        disc_interpolates = netD(interpolates)
    ones = torch.ones(disc_interpolates.size(), device=disc_interpolates.device)
    with _ops.input_grad_only():
        gradients = torch.autograd.grad(outputs=disc_interpolates, inputs=interpolates, grad_outputs=ones,
                                        create_graph=True, retain_graph=True, only_inputs=True)[0]
    g5 = gradients.unsqueeze(2) if gradients.dim() == 4 else gradients
    return _ops.GpPenalty.apply(g5, float(LAMBDA))
