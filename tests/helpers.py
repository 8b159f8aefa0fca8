"""Shared helpers of the parity tests: build a state_dict from a fixture recipe and run the oracle port on it."""
import torch

from oracle import port


def opt_from(fx):
    o = port.Opt()
    o.__dict__.update(fx['opt'])
    return o


def state_from(fx, dtype=torch.float32):
    """state_dict with the fixture's keys/shapes filled by the deterministic recipe"""
    sd = {}
    for k, shape in fx['state']:
        if k.endswith('num_batches_tracked'):
            sd[k] = torch.zeros(shape, dtype=torch.int64)
        else:
            sd[k] = torch.zeros(shape, dtype=torch.float32)
    port.det_fill(sd, fx['fill_seed'])
    if dtype != torch.float32:
        sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    return sd


def with_grad(sd):
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
            v.requires_grad_(True)
    return sd


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def param_key(k):
    """oracle/state_dict key -> named_parameters key of the reference (identical)"""
    return k


def state_d_from(fx):
    sd = {k: torch.zeros(shape, dtype=torch.float32) for k, shape in fx['state_d']}
    port.det_fill(sd, fx['fill_seed_d'])
    return sd


def train_opt_from(fx):
    o = opt_from(fx)
    o.Noise_Amps = list(fx['amps_before'])
    return o
