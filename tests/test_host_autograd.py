"""Autograd plumbing of hpvg.ops.ConvFwd / WeightProxy on the CPU, with the two raw kernel wrappers replaced by torch's own
convolution (the kernels themselves are GPU-tested in tests/test_gpu_layers.py): the default node and the deferred
weight-gradient form (HPVG_CRITIC_WSIDE, hpvg.ops.deferred_weight) must both return torch's gradients, and a proxy that
nothing was deposited into must be a plain identity."""
import contextlib

import pytest
import torch
import torch.nn.functional as F


@pytest.fixture
def cpu_kernels(monkeypatch):
    from hpvg import ops
    calls = {"wgrad": 0}

    def conv_raw(x, w, bias, pad, transposed, out_wide, act_slope=None, stats=None, mask_src=None, mask_slope=None, **kw):
        assert act_slope is None and stats is None and mask_src is None
        if transposed:      # data gradient: `pad` is 2 - p of the forward conv
            return F.conv_transpose3d(x, w, None, padding=2 - pad)
        return F.conv3d(x, w, bias, padding=pad)

    def wgrad_raw(x, gy, pad, wshape, want_bias=False):
        import inspect
        calls["wgrad"] += 1
        caller = inspect.stack()[1]
        calls["deferred"] = calls.get("deferred", 0) + (caller.function == "backward")     # WeightProxy.backward vs ConvWgrad.forward
        return torch.nn.grad.conv3d_weight(x, wshape, gy, padding=pad), None

    monkeypatch.setattr(ops, "conv_raw", conv_raw)
    monkeypatch.setattr(ops, "wgrad_raw", wgrad_raw)
    monkeypatch.setattr(ops, "channel_sum", lambda t: t.sum((0, 2, 3, 4)))
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: object())
    monkeypatch.setattr(torch.Tensor, "record_stream", lambda self, s: None, raising=False)
    return calls


def _case(seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((1, 3, 4, 6, 5), generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn((4, 3, 3, 3, 3), generator=g, dtype=torch.float64, requires_grad=True)
    b = torch.randn((4,), generator=g, dtype=torch.float64, requires_grad=True)
    gy = torch.randn((1, 4, 4, 6, 5), generator=g, dtype=torch.float64)
    return x, w, b, gy


def _reference(x, w, b, gy):
    y = F.conv3d(x, w, b, padding=1)
    return torch.autograd.grad(y, (x, w, b), gy)


@pytest.mark.parametrize("deferred", [False, True])
def test_conv_node_gradients(cpu_kernels, monkeypatch, deferred):
    from hpvg import ops
    x, w, b, gy = _case(1)
    ref = _reference(x, w, b, gy)
    monkeypatch.setattr(ops, "_CRITIC_WSIDE", [deferred])
    monkeypatch.setattr(ops, "_WGRAD_STREAM", [ops._StreamRing([object()]) if deferred else None])
    w_use, token = ops.deferred_weight(w)
    assert (token is not None) == deferred
    y = ops.conv(x, w_use, b, 1, False, token=token)
    got = torch.autograd.grad(y, (x, w, b), gy)
    for a, r in zip(got, ref):
        torch.testing.assert_close(a, r, rtol=1e-10, atol=1e-12)
    assert cpu_kernels["wgrad"] == 1
    if deferred:
        assert token.slot is None          # consumed by WeightProxy.backward


def test_proxy_is_off_inside_the_gradient_penalty_pass_and_without_a_side_stream(cpu_kernels, monkeypatch):
    from hpvg import ops
    _, w, _, _ = _case(2)
    monkeypatch.setattr(ops, "_CRITIC_WSIDE", [True])
    monkeypatch.setattr(ops, "_WGRAD_STREAM", [None])
    assert ops.deferred_weight(w) == (w, None) or ops.deferred_weight(w)[1] is None
    monkeypatch.setattr(ops, "_WGRAD_STREAM", [ops._StreamRing([object()])])
    with ops.no_wgrad_proxy():
        assert ops.deferred_weight(w)[1] is None
    with torch.no_grad():
        assert ops.deferred_weight(w)[1] is None
    assert ops.deferred_weight(w.detach())[1] is None
    assert ops.deferred_weight(w)[1] is not None


def test_proxy_without_a_deposit_is_an_identity(cpu_kernels, monkeypatch):
    """a node that computed its weight gradient itself (create_graph sweep) sends a REAL gradient through the proxy"""
    from hpvg import ops
    x, w, b, gy = _case(3)
    monkeypatch.setattr(ops, "_CRITIC_WSIDE", [True])
    monkeypatch.setattr(ops, "_WGRAD_STREAM", [ops._StreamRing([object()])])
    w_use, token = ops.deferred_weight(w)
    y = ops.conv(x, w_use, b, 1, False, token=token)
    # create_graph=True: ConvFwd.backward is not `plain`, takes the differentiable ConvWgrad path and deposits nothing
    got = torch.autograd.grad(y, (w,), gy, create_graph=True)[0]
    torch.testing.assert_close(got, _reference(x, w, b, gy)[1], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("deferred", [False, True])
def test_gradient_penalty_like_double_backward(cpu_kernels, monkeypatch, deferred):
    """calc_gradient_penalty's pattern (modules/utils.py:14-18): a create_graph sweep for d/d(input) only, then a backward of a
    function of that gradient.  The weight is referenced by the ConvFwd node and by the sweep's ConvDgrad node: with deferred
    weight gradients each gets its own proxy, and the weight gradient is the sum of both contributions, as in torch."""
    from hpvg import ops
    x, w, b, gy = _case(4)
    ones = torch.ones_like(gy)

    def run(conv):
        for t in (x, w, b):
            t.grad = None
        y = conv()
        ctx = ops.input_grad_only() if conv is not torch_conv else contextlib.nullcontext()
        with ctx:
            gx = torch.autograd.grad(y, x, ones, create_graph=True)[0]
        ((gx ** 2).sum() + (y * gy).sum()).backward()
        return x.grad.clone(), w.grad.clone(), b.grad.clone()

    def torch_conv():
        return F.conv3d(x, w, b, padding=1)

    ref = run(torch_conv)
    monkeypatch.setattr(ops, "_CRITIC_WSIDE", [deferred])
    monkeypatch.setattr(ops, "_WGRAD_STREAM", [ops._StreamRing([object()]) if deferred else None])
    tokens = []

    def hpvg_conv():
        w_use, token = ops.deferred_weight(w)
        tokens.append(token)
        return ops.conv(x, w_use, b, 1, False, token=token)

    cpu_kernels["wgrad"] = cpu_kernels["deferred"] = 0
    got = run(hpvg_conv)
    for a, r in zip(got, ref):
        torch.testing.assert_close(a, r, rtol=1e-10, atol=1e-12)
    assert cpu_kernels["wgrad"] == 2                  # the ConvFwd node's and the ConvDgrad node's contribution
    assert cpu_kernels["deferred"] == (2 if deferred else 0)      # both computed by a WeightProxy, or both by ConvWgrad nodes
    assert (tokens[0] is not None) == deferred
