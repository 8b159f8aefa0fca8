"""Gradient clipping + Adam of one training iteration on the library's multi-tensor kernels.

Replaces, for CUDA parameters, the pair the reference ends every iteration with:
    torch.nn.utils.clip_grad_norm_(G_curr.parameters(), opt.grad_clip)      train_video.py:201, train_image.py:216
    optimizerG.step() / optimizerD.step()                                   train_video.py:183,202 (optim.Adam, :55,:88)
by `hpvg_grad_clip_coef` (fixed-order squared-norm reduction -> clip coefficient, one launch) and `hpvg_adam_step`
(gradient scaling, both moments and the parameter update of every tensor, one launch per 32 tensors).  The step count
lives on the device, so the pair records into a CUDA graph.  Same update rule, same `state_dict()` layout
(`step`, `exp_avg`, `exp_avg_sq` per parameter; param_groups with lr / betas / eps) as torch.optim.Adam.
There is no CPU path: CPU parameters raise HpvgError.
"""
import ctypes

import torch

from . import lib
from .lib import HpvgError


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas) with the reference's remaining defaults (eps 1e-8, weight_decay 0, amsgrad False)"""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        if not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 or eps < 0.0 or lr < 0.0:
            raise ValueError("invalid Adam hyper-parameters: lr=%r betas=%r eps=%r" % (lr, betas, eps))
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=0, amsgrad=False, maximize=False))
        self._dev_state = None      # HPVG_OPT_STATE_FLOATS floats: step, clip coefficient, total norm, tickets
        self._partials = None
        self._owned_ids = None
        self.last_launches = 0

    def state_dict(self):
        """torch.optim.Adam's layout.  Every parameter gets its OWN host copy of the step count (what a default
        torch.optim.Adam holds): the live state shares one device counter between all parameters, and an optimizer that
        loaded aliased tensors would advance it once per parameter."""
        sd = super().state_dict()
        sd['state'] = {k: {name: (val.detach().cpu().clone() if name == 'step' and torch.is_tensor(val) else val) for name, val in st.items()}
                       for k, st in sd['state'].items()}
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._dev_state = None      # rebuilt from the loaded per-parameter step on the next step()
        self._owned_ids = None

    def _device_state(self, device):
        if self._dev_state is None or self._dev_state.device != device:
            self._dev_state = torch.zeros(lib.OPT_STATE_FLOATS, dtype=torch.float32, device=device)
            steps = [float(st['step']) for st in self.state.values() if st.get('step') is not None]
            if steps:
                if max(steps) != min(steps):
                    raise HpvgError("hpvg.optim.Adam keeps one step count per optimizer; the loaded state has %r" % sorted(set(steps)))
                self._dev_state[0] = steps[0]
            for st in self.state.values():
                if 'exp_avg' in st:
                    st['step'] = self._dev_state[0]      # 0-dim view: every parameter reports the shared device counter
        return self._dev_state

    @staticmethod
    def _check(t, what):
        if not t.is_cuda:
            raise HpvgError("hpvg.optim.Adam: %s is on %s; the optimizer kernels run on CUDA tensors only (no CPU fallback)" % (what, t.device))
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise HpvgError("hpvg.optim.Adam: %s must be contiguous float32 (got %s, contiguous=%s)" % (what, t.dtype, t.is_contiguous()))

    @torch.no_grad()
    def step(self, closure=None, clip_params=None, max_norm=None):
        """One Adam step.  clip_params / max_norm: clip_grad_norm_(clip_params, max_norm) is applied first (the norm runs
        over every tensor of clip_params that has a gradient, whether this optimizer owns it or not, and all of those
        gradients are scaled in place — train_video.py:201 clips G_curr.parameters(), of which the optimizer owns a part)."""
        if closure is not None:
            raise HpvgError("hpvg.optim.Adam.step does not take a closure (the reference never passes one)")
        owned = []      # (param, grad, exp_avg, exp_avg_sq, lr)
        hyper = None
        for group in self.param_groups:
            h = (float(group['betas'][0]), float(group['betas'][1]), float(group['eps']))
            if hyper is None:
                hyper = h
            elif h != hyper:
                raise HpvgError("hpvg.optim.Adam: betas / eps must be the same in every param group (only lr differs in the reference)")
            if group.get('weight_decay', 0) or group.get('amsgrad', False) or group.get('maximize', False):
                raise HpvgError("hpvg.optim.Adam: weight_decay / amsgrad / maximize are not on the reference path")
            for p in group['params']:
                if p.grad is None:
                    continue
                self._check(p, "a parameter")
                self._check(p.grad, "a gradient")
                st = self.state[p]
                if 'exp_avg' not in st:
                    st['exp_avg'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['exp_avg_sq'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st['step'] = None
                owned.append((p, p.grad, st['exp_avg'], st['exp_avg_sq'], float(group['lr'])))
        # ONE device-side step count per optimizer (torch keeps one per parameter): equivalent as long as the same parameters
        # receive gradients at every step, which holds for the reference's loops; anything else is refused, not approximated
        ids = tuple(id(p) for p, *_ in owned)
        if self._owned_ids is None:
            self._owned_ids = ids
        elif ids != self._owned_ids:
            raise HpvgError("hpvg.optim.Adam: the set of parameters with gradients changed between steps (%d -> %d tensors); "
                            "per-parameter step counts would diverge from the shared one" % (len(self._owned_ids), len(ids)))
        clip = max_norm is not None
        extra = []
        if clip:
            if clip_params is None:
                raise ValueError("max_norm without clip_params")
            mine = {id(p) for p, *_ in owned}
            for p in clip_params:
                if p.grad is not None and id(p) not in mine:
                    self._check(p.grad, "a gradient")
                    extra.append(p.grad)
        if not owned and not extra:
            return None
        device = (owned[0][0] if owned else extra[0]).device
        state = self._device_state(device)
        for p, *_ in owned:
            if self.state[p]['step'] is None:
                self.state[p]['step'] = state[0]
        st_ptr, stream = ctypes.c_void_p(state.data_ptr()), _stream()
        launches = 0
        MAXT = lib.OPT_MAX_TENSORS
        grads = [g for _, g, _, _, _ in owned] + extra
        if clip:
            slots = len(grads) * lib.OPT_BLOCKS
            if self._partials is None or self._partials.numel() < slots or self._partials.device != device:
                self._partials = torch.empty(slots, dtype=torch.float32, device=device)
            for c0 in range(0, len(grads), MAXT):
                chunk = grads[c0:c0 + MAXT]
                lib.call("hpvg_grad_clip_coef", len(chunk), lib.ptr_array(chunk), lib.longlong_array([g.numel() for g in chunk]),
                         ctypes.c_void_p(self._partials.data_ptr()), c0 * lib.OPT_BLOCKS, slots, int(c0 + MAXT >= len(grads)),
                         float(max_norm), st_ptr, stream)
                launches += 1
        # gradients of tensors the optimizer does not own are only scaled (exp_avg == NULL)
        rows = [(p, g, m, v, lr) for p, g, m, v, lr in owned] + ([(None, g, None, None, 0.0) for g in extra] if clip else [])

        def ptrs(ts):
            return (ctypes.c_void_p * len(ts))(*[None if t is None else t.data_ptr() for t in ts])

        for c0 in range(0, len(rows), MAXT):
            chunk = rows[c0:c0 + MAXT]
            lib.call("hpvg_adam_step", len(chunk), ptrs([r[0] for r in chunk]), ptrs([r[1] for r in chunk]), ptrs([r[2] for r in chunk]),
                     ptrs([r[3] for r in chunk]), lib.longlong_array([r[1].numel() for r in chunk]),
                     (ctypes.c_float * len(chunk))(*[r[4] for r in chunk]), hyper[0], hyper[1], hyper[2], int(clip),
                     int(c0 + MAXT >= len(rows)), st_ptr, stream)
            launches += 1
        # the kernels wrote through raw pointers: move the version counters as torch's in-place update would (the packed bf16
        # weight images of hpvg.ops are stamped with them)
        torch.autograd.graph.increment_version([r[0] for r in rows if r[0] is not None] + ([r[1] for r in rows] if clip else []))
        self.last_launches = launches
        return None

    def total_norm(self):
        """the gradient norm the last clipped step saw (device scalar; what clip_grad_norm_ returns)"""
        if self._dev_state is None:
            raise HpvgError("no step has run yet")
        return self._dev_state[2]


def use_library_optimizer(params):
    """True when the trainers should build hpvg.optim.Adam: CUDA parameters, unless HPVG_TORCH_ADAM=1 asks for torch's"""
    import os
    if os.environ.get("HPVG_TORCH_ADAM", "0") == "1":
        return False
    for p in params:
        return bool(p.is_cuda)
    return False
