"""hpvg.optim.Adam (hpvg_grad_clip_coef + hpvg_adam_step) against what it replaces: torch.nn.utils.clip_grad_norm_ followed by
torch.optim.Adam(betas=(beta1, 0.999)).step() (train_video.py:201-202, :183), on identical parameters and gradients."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [(64, 64, 3, 3, 3), (64,), (64, 3, 3, 3, 3), (1, 64, 3, 3, 3), (1,), (7, 5, 3), (129,), (3, 64, 3, 3)]


def _make(n_tensors, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter((torch.randn(SHAPES[i % len(SHAPES)], generator=g) * 0.05).cuda()) for i in range(n_tensors)]


def _clone(params):
    return [torch.nn.Parameter(p.detach().clone()) for p in params]


@pytest.mark.parametrize("n_owned,n_extra,max_norm", [(21, 0, None), (40, 5, 5.0), (12, 30, 0.05)])
def test_adam_and_clipping_match_torch(n_owned, n_extra, max_norm):
    """> 32 tensors crosses the per-launch chunk; max_norm 0.05 makes the clip coefficient bite, 5.0 leaves it at 1;
    `extra` tensors are clipped but not owned by the optimizer (the frozen-but-differentiated blocks of the generator)"""
    from hpvg import optim
    ours, extra_o = _make(n_owned, 1), _make(n_extra, 2)
    theirs, extra_t = _clone(ours), _clone(extra_o)
    half = n_owned // 2

    def groups(ps):
        return [{"params": ps[:half], "lr": 5e-4 * 0.2}, {"params": ps[half:], "lr": 5e-4}]

    opt_o = optim.Adam(groups(ours), lr=5e-4, betas=(0.5, 0.999))
    opt_t = torch.optim.Adam(groups(theirs), lr=5e-4, betas=(0.5, 0.999))
    gen = torch.Generator().manual_seed(3)
    for step in range(6):
        scale = 10.0 ** (step - 3)        # gradient magnitudes over six decades
        for a, b in zip(ours + extra_o, theirs + extra_t):
            gr = (torch.randn(a.shape, generator=gen) * scale).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()
        if step == 2 and n_extra:         # a clipped tensor without a gradient is skipped by both
            extra_o[0].grad = extra_t[0].grad = None
        if max_norm is None:
            opt_o.step()
        else:
            opt_o.step(clip_params=ours + extra_o, max_norm=max_norm)
            total = torch.nn.utils.clip_grad_norm_(theirs + extra_t, max_norm)
            torch.testing.assert_close(opt_o.total_norm(), total, rtol=2e-6, atol=0)
        opt_t.step()
        for a, b in zip(ours + extra_o, theirs + extra_t):
            if a.grad is not None:
                torch.testing.assert_close(a.grad, b.grad, rtol=1e-5, atol=1e-30)
    for a, b in zip(ours, theirs):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=2e-7)
        # the two clip coefficients differ in the last bits (sum of squares here, norm of per-tensor norms in torch); moments
        # that cancel to near zero carry that as an absolute error on the scale of the tensor
        m_o, m_t = opt_o.state[a]['exp_avg'], opt_t.state[b]['exp_avg']
        v_o, v_t = opt_o.state[a]['exp_avg_sq'], opt_t.state[b]['exp_avg_sq']
        torch.testing.assert_close(m_o, m_t, rtol=1e-5, atol=2e-6 * m_t.abs().max().item())
        torch.testing.assert_close(v_o, v_t, rtol=1e-5, atol=2e-6 * v_t.abs().max().item())
    assert float(opt_o.state[ours[1]]['step']) == 6.0
    assert ours[1]._version >= 6          # the packed-weight caches of hpvg.ops watch the version counters


def test_state_dict_moves_between_the_library_optimizer_and_torch():
    """the reference checkpoints optimizer.state_dict() (train_video.py:250,257): same layout in both directions"""
    from hpvg import optim
    ours = _make(5, 4)
    theirs = _clone(ours)
    opt_o = optim.Adam(ours, lr=5e-4, betas=(0.5, 0.999))
    gen = torch.Generator().manual_seed(5)

    def grads():
        for a, b in zip(ours, theirs):
            gr = torch.randn(a.shape, generator=gen).cuda()
            a.grad, b.grad = gr.clone(), gr.clone()

    for _ in range(3):
        grads()
        opt_o.step()
    for a, b in zip(ours, theirs):
        b.data.copy_(a.data)
    opt_t = torch.optim.Adam(theirs, lr=5e-4, betas=(0.5, 0.999))
    # (deepcopy stands in for the torch.save / torch.load round trip: load_state_dict keeps tensors that already have the
    # parameter's dtype and device, and two live optimizers must not share moments)
    opt_t.load_state_dict(copy.deepcopy(opt_o.state_dict()))
    grads()
    opt_o.step()
    opt_t.step()
    for a, b in zip(ours, theirs):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=2e-7)
    # and back: a fresh library optimizer continues from torch's state
    opt_o2 = optim.Adam(ours, lr=5e-4, betas=(0.5, 0.999))
    opt_o2.load_state_dict(copy.deepcopy(opt_t.state_dict()))
    grads()
    opt_o2.step()
    opt_t.step()
    for a, b in zip(ours, theirs):
        torch.testing.assert_close(a, b, rtol=1e-5, atol=2e-7)
    assert float(opt_o2.state[ours[0]]['step']) == 5.0


def test_recorded_steps_advance_the_device_step_count():
    """the pair records into a CUDA graph: replays advance the step count and match eager steps on the same gradients"""
    from hpvg import optim
    ours = _make(6, 6)
    eager = _clone(ours)
    opt_g = optim.Adam(ours, lr=5e-4, betas=(0.5, 0.999))
    opt_e = optim.Adam(eager, lr=5e-4, betas=(0.5, 0.999))
    gen = torch.Generator().manual_seed(7)
    for a, b in zip(ours, eager):
        gr = torch.randn(a.shape, generator=gen).cuda()
        a.grad, b.grad = gr.clone(), gr.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt_g.step(clip_params=ours, max_norm=1.0)
    torch.cuda.current_stream().wait_stream(side)
    opt_e.step(clip_params=eager, max_norm=1.0)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt_g.step(clip_params=ours, max_norm=1.0)
    for _ in range(3):
        graph.replay()
        opt_e.step(clip_params=eager, max_norm=1.0)      # the gradients keep shrinking by the same coefficient in both
    torch.cuda.synchronize()
    assert float(opt_g.state[ours[0]]['step']) == 4.0
    for a, b in zip(ours, eager):
        torch.testing.assert_close(a, b, rtol=0, atol=0)
        torch.testing.assert_close(a.grad, b.grad, rtol=0, atol=0)


def test_changing_gradient_set_is_refused():
    """one step count per optimizer: a parameter that skips a step would need its own count (torch's behaviour); refused loudly"""
    from hpvg import optim
    from hpvg.lib import HpvgError
    ours = _make(3, 8)
    opt = optim.Adam(ours, lr=5e-4, betas=(0.5, 0.999))
    for p in ours:
        p.grad = torch.ones_like(p)
    opt.step()
    ours[1].grad = None
    with pytest.raises(HpvgError):
        opt.step()
