"""The tcgen05 kernels of the narrow network ends (csrc/narrow_tc.cu) against PyTorch's CPU convolution and its adjoints.

expand_tc_kernel          thin (float32 NCDHW, 1 or 3 channels) -> wide (bf16 NDHWC, 64 channels): the 3 -> 64 head convolutions
                          (reference modules/networks_3d.py:51,63 with in_channel = nc_im) and the data gradient of the
                          64 -> 3 / 64 -> 1 tails (:175,341,362)
narrow_wgrad_tc_kernel    both narrow weight gradients (aten::convolution_backward grad_weight of those layers) + the bias gradient

Cases the full-size and per-layer tests do not pin: ragged volumes (tiles that hang over every face), padding 0 / 1 / 2, 2-D
(KD = 1), one input channel, batches (the tile counter's carry into the sample index) with per-sample BatchNorm sums, volumes
with more tiles than SMs and with fewer.  Operands are bf16-representable, so both sides multiply the same values: what is left
is summation order and one bf16 rounding of a stored wide result.
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

WIDE_OUT_TOL = 3e-3
F32_OUT_TOL = 2e-4


def _vals(shape, seed, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=gen) * scale).bfloat16().float()


def _wide(t):
    return t.permute(0, 2, 3, 4, 1).contiguous().to(device='cuda', dtype=torch.bfloat16)


def _ncdhw(t):
    return t.float().permute(0, 4, 1, 2, 3).contiguous().cpu()


CASES = [
    # (N, Cthin, D, H, W, pad)
    (1, 3, 3, 20, 9, 1),       # ragged: every tile hangs over
    (2, 3, 5, 17, 19, 1),      # batch: the tile counter carries into the sample index
    (1, 3, 4, 33, 40, 0),      # pad 0: output smaller than input
    (1, 3, 2, 11, 13, 2),      # pad 2: output larger than input
    (1, 1, 6, 24, 31, 1),      # one thin channel (the critic's 64 -> 1 tail)
    (3, 3, 16, 64, 64, 1),     # 1 536 tiles on 148 SMs: ten tiles per CTA, three samples
]
IDS = ["3x20x9", "n2_5x17x19", "pad0_4x33x40", "pad2_2x11x13", "c1_6x24x31", "n3_16x64x64"]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_thin_to_wide_convolution_and_its_per_sample_sums(case):
    from hpvg import ops
    n, c, d, h, w, pad = case
    x = _vals((n, c, d, h, w), 1)
    wt = _vals((64, c, 3, 3, 3), 2, 0.1)
    bias = _vals((64,), 3, 0.1)
    y_ref = F.conv3d(x, wt, bias, padding=pad)
    stats = torch.zeros((n, 128), device='cuda')
    y = ops.conv_raw(x.cuda(), wt.cuda(), bias.cuda(), pad, False, True, stats=stats, stats_per_sample=True)
    y_act = ops.conv_raw(x.cuda(), wt.cuda(), bias.cuda(), pad, False, True, act_slope=0.2)
    assert rel_err(_ncdhw(y), y_ref) < WIDE_OUT_TOL
    assert rel_err(_ncdhw(y_act), F.leaky_relu(y_ref, 0.2)) < WIDE_OUT_TOL
    assert rel_err(stats[:, :64].cpu(), y_ref.sum((2, 3, 4))) < 3e-3
    assert rel_err(stats[:, 64:].cpu(), (y_ref * y_ref).sum((2, 3, 4))) < 2e-3
    # the data gradient of a wide -> thin layer with the same filter extents is the same kernel on the flipped filter
    wt_tail = _vals((c, 64, 3, 3, 3), 4, 0.1)
    g_thin = _vals((n, c, d + 2 * pad - 2, h + 2 * pad - 2, w + 2 * pad - 2), 5)
    gx_ref = torch.nn.grad.conv3d_input((n, 64, d, h, w), wt_tail, g_thin, padding=pad)
    gx = ops.conv_raw(g_thin.cuda(), wt_tail.cuda(), None, 2 - pad, True, True)      # the adjoint of a pad-p convolution pads 2 - p
    assert rel_err(_ncdhw(gx), gx_ref) < WIDE_OUT_TOL


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_narrow_weight_gradients(case):
    from hpvg import ops
    n, c, d, h, w, pad = case
    do, ho, wo = d + 2 * pad - 2, h + 2 * pad - 2, w + 2 * pad - 2
    # head: x thin, gy wide
    x = _vals((n, c, d, h, w), 11)
    gy = _vals((n, 64, do, ho, wo), 12)
    dw_ref = torch.nn.grad.conv3d_weight(x, (64, c, 3, 3, 3), gy, padding=pad)
    dw, db = ops.wgrad_raw(x.cuda(), _wide(gy), pad, (64, c, 3, 3, 3), want_bias=True)
    assert rel_err(dw.cpu(), dw_ref) < 1e-3
    assert rel_err(db.cpu(), gy.sum((0, 2, 3, 4))) < F32_OUT_TOL
    # tail: x wide, gy thin
    xw = _vals((n, 64, d, h, w), 13)
    gt = _vals((n, c, do, ho, wo), 14)
    dwt_ref = torch.nn.grad.conv3d_weight(xw, (c, 64, 3, 3, 3), gt, padding=pad)
    dwt, dbt = ops.wgrad_raw(_wide(xw), gt.cuda(), pad, (c, 64, 3, 3, 3), want_bias=True)
    assert rel_err(dwt.cpu(), dwt_ref) < 1e-3
    assert rel_err(dbt.cpu(), gt.sum((0, 2, 3, 4))) < F32_OUT_TOL


@pytest.mark.parametrize("c", [3, 1])
def test_two_dimensional_narrow_layers(c):
    """train_image.py's layers: KD = 1 (reference modules/networks_2d.py:56)"""
    from hpvg import ops
    n, h, w = 2, 45, 70
    x = _vals((n, c, h, w), 21)
    wt = _vals((64, c, 3, 3), 22, 0.1)
    bias = _vals((64,), 23, 0.1)
    gy = _vals((n, 64, h, w), 24)
    y_ref = F.conv2d(x, wt, bias, padding=1)
    dw_ref = torch.nn.grad.conv2d_weight(x, (64, c, 3, 3), gy, padding=1)
    x5, gy5 = x.unsqueeze(2), gy.unsqueeze(2)
    y = ops.conv_raw(x5.cuda(), wt.cuda(), bias.cuda(), 1, False, True)
    dw, db = ops.wgrad_raw(x5.cuda(), _wide(gy5), 1, (64, c, 3, 3), want_bias=True)
    assert rel_err(_ncdhw(y).squeeze(2), y_ref) < WIDE_OUT_TOL
    assert rel_err(dw.cpu(), dw_ref) < 1e-3
    assert rel_err(db.cpu(), gy.sum((0, 2, 3))) < F32_OUT_TOL
    xw = _vals((n, 64, h, w), 25)
    gt = _vals((n, c, h, w), 26)
    dwt_ref = torch.nn.grad.conv2d_weight(xw, (c, 64, 3, 3), gt, padding=1)
    dwt, _ = ops.wgrad_raw(_wide(xw.unsqueeze(2)), gt.unsqueeze(2).cuda(), 1, (c, 64, 3, 3), want_bias=True)
    assert rel_err(dwt.cpu(), dwt_ref) < 1e-3


def test_the_tcgen05_kernels_are_the_ones_that_run(monkeypatch):
    """the narrow layers must not fall back to the mma.sync kernels silently: count the launches of one call through ncu-free means —
    the library's own kernel-choice entry point"""
    from hpvg import lib
    l = lib.load()
    if not hasattr(l, "hpvg_narrow_kernel_choice"):
        pytest.skip("library without hpvg_narrow_kernel_choice")
    assert l.hpvg_narrow_kernel_choice(3, 3) == 1      # Cthin = 3, KD = 3 -> tcgen05
    assert l.hpvg_narrow_kernel_choice(1, 1) == 1
    assert l.hpvg_narrow_kernel_choice(4, 3) == 0      # four thin channels: CUDA-core kernels
