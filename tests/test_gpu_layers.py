"""Per-layer GPU parity: every layer type of the path in isolation (forward tensor, input gradient, parameter
gradients) against the CPU oracle's fp32 arithmetic on identical inputs, through the reference's module interface.

Tolerances (north_star): relative 2e-2 for layers with bf16 operands (every 64/128-channel activation is stored in
bf16), 1e-3 or tighter for the fp32 kernels (resize, tanh, KL, gradient penalty, spectral norm).
At BASELINE.json's full size (16 x 64 x 64, 64 channels) the oracle would take too long per case, so the kernels are
checked there through size-independent properties: the adjoint identities <conv(x), g> = <x, dgrad(g)> and
<wgrad(x, g), w'> = <conv_{w'}(x), g>, and BatchNorm's zero-mean / unit-variance output statistics.
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err
from oracle import port

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2
# Input gradients that pass through a LeakyReLU whose pre-activation is stored in bf16: the derivative (0.2 | 1) is read
# from the sign of the STORED value, so it differs from the fp32 reference wherever rounding moved a pre-activation across
# zero (about 0.3 % of the elements for a BatchNorm output with |mean| ~ std; each such element is off by 0.8 |g|, i.e.
# sqrt(0.003) * 0.8 = 4e-2 in relative L2; the weight gradient of the layer consumes the same masked tensor).  Forward
# tensors and the gradients of activation-free layers meet BF16_TOL; for gradients behind a LeakyReLU the bar against fp32 is
# LRELU_GRAD_TOL, and the bar against the oracle's bf16 storage emulation — which
# stores the same values and therefore reads the same signs — is EMU_BWD_TOL.
LRELU_GRAD_TOL = 5e-2
FUSED_GRAD_TOL = 3.5e-2     # the one-launch ConvBlock3D (fp32 statistics and sign): the bf16-operand floor, see test_convblock3d_layer
F32_TOL = 1e-3
EMU_FWD_TOL = 5e-4     # one layer vs the oracle with bf16 storage emulation: what is left is fp32 summation order
EMU_BWD_TOL = 3e-3


def _fill_module(m, seed):
    sd = m.state_dict()
    port.det_fill(sd, seed)
    return {k: v.detach().clone() for k, v in sd.items()}


def _leafs(sd):
    for k, v in sd.items():
        if v.is_floating_point() and not k.endswith(('running_mean', 'running_var', 'weight_u', 'weight_v')):
            v.requires_grad_(True)
    return sd


def _run_block(m, x_gpu):
    """a block as the networks chain it: 64/128-channel inputs arrive as bf16 NDHWC tensors from the previous block (the
    module's own forward(), which takes float32 NCDHW, is the reference-API entry used for 3-channel inputs)"""
    from hpvg import ops
    if x_gpu.shape[1] >= 64:
        x5 = x_gpu.unsqueeze(2) if x_gpu.dim() == 4 else x_gpu
        y = ops.ToThin.apply(m.run(ops.ToWide.apply(x5)))
        return y.squeeze(2) if x_gpu.dim() == 4 else y
    return m(x_gpu)


def _compare_grads(module, sd, tol=None, prefix=''):
    tol = LRELU_GRAD_TOL if tol is None else tol
    worst = 0.0
    big = max(v.grad.norm().item() for v in sd.values() if v.is_floating_point() and v.grad is not None)
    for k, p in module.named_parameters():
        ref = sd[prefix + k].grad
        assert p.grad is not None and ref is not None, k
        d = (p.grad.detach().cpu().double() - ref.double()).norm().item()
        # a conv bias in front of BatchNorm has a mathematically zero gradient: absolute floor
        assert d <= tol * ref.double().norm().item() + 2e-3 * big, (k, d, ref.norm().item())
        worst = max(worst, d / (ref.double().norm().item() + 2e-3 * big))
    return worst


@pytest.mark.parametrize("cin,cout,shape", [(64, 64, (1, 6, 20, 24)), (3, 64, (2, 5, 17, 19)), (128, 64, (1, 4, 16, 16)),
                                            (64, 64, (1, 3, 33, 9))])
def test_convblock3d_layer(cin, cout, shape):
    """ConvBlock3D = Conv3d + BatchNorm3d(train) + LeakyReLU (reference modules/networks_3d.py:48-56)"""
    from modules import networks_3d
    m = networks_3d.ConvBlock3D(cin, cout, 3, 1, 1)
    sd = _leafs(_fill_module(m, 5))
    m.cuda()
    n, d, h, w = shape
    x = port.det_tensor((n, cin, d, h, w), 1)
    g = port.det_tensor((n, cout, d, h, w), 2)
    if cin >= 64:      # what the preceding layer hands over is a bf16 tensor: give both sides the same values
        x = x.bfloat16().float()
    x_ref = x.clone().requires_grad_(True)
    y_ref = port.conv_block(sd, '', x_ref, 1)
    y_ref.backward(g)
    x_gpu = x.cuda().requires_grad_(True)
    y = _run_block(m, x_gpu)
    assert y.shape == y_ref.shape and y.dtype == torch.float32
    assert rel_err(y, y_ref) < BF16_TOL
    y.backward(g.cuda())
    # 64 -> 64 3-D blocks run as ONE launch that takes BatchNorm statistics and the LeakyReLU sign from the fp32 accumulators
    # (hpvg_conv_bn_lrelu_fused); the other blocks read the sign of a bf16-stored value.  What is left for the fused layer is the
    # floor of bf16 OPERANDS: rounding weights and inputs to bf16 moves every pre-activation by ~3e-3 sigma, which flips the
    # sign of a fraction f ~ 2.4e-3 of them, and each flip changes that gradient element by 0.8 |g|: sqrt(f) * 0.8 = 2-3e-2 in
    # relative L2 (measured 2.8e-2 here; 2.3e-3 when the operands are bf16-representable, tests/test_gpu_fullsize.py).
    tol = FUSED_GRAD_TOL if (cin == 64 and cout == 64) else LRELU_GRAD_TOL
    print("ConvBlock3D %d -> %d %s: input-gradient error vs fp32 %.2e (bar %.0e)" % (cin, cout, shape, rel_err(x_gpu.grad, x_ref.grad), tol))
    assert rel_err(x_gpu.grad, x_ref.grad) < tol
    _compare_grads(m, sd, tol)
    for k in ('norm.running_mean', 'norm.running_var'):
        assert rel_err(dict(m.named_buffers())[k], sd[k]) < 1e-3, k
    assert int(m.norm.num_batches_tracked.item()) == 1
    # the same layer against the oracle's bf16 storage emulation: tight
    sd2 = _leafs({k: v.detach().clone() for k, v in _fill_module(networks_3d.ConvBlock3D(cin, cout, 3, 1, 1), 5).items()})
    x_emu = x.clone().requires_grad_(True)
    with port.storage('bf16'):
        y_emu = port.conv_block(sd2, '', x_emu, 1)
        y_emu.backward(g.bfloat16().float())
    assert rel_err(y, y_emu) < EMU_FWD_TOL
    if cin >= 64:
        assert rel_err(x_gpu.grad, x_emu.grad.bfloat16().float()) < EMU_BWD_TOL
    assert rel_err(m.conv.weight.grad, sd2['conv.weight'].grad) < EMU_BWD_TOL
    assert rel_err(m.norm.weight.grad, sd2['norm.weight'].grad) < EMU_BWD_TOL


@pytest.mark.parametrize("cin,shape", [(64, (1, 5, 18, 20)), (3, (1, 4, 15, 13))])
def test_convblock3dsn_layer(cin, shape):
    """ConvBlock3DSN = spectral-norm Conv3d + LeakyReLU (reference modules/networks_3d.py:59-70)"""
    from modules import networks_3d
    m = networks_3d.ConvBlock3DSN(cin, 64, 3, 1, 1)
    sd = _leafs(_fill_module(m, 6))
    m.cuda()
    n, d, h, w = shape
    x = port.det_tensor((n, cin, d, h, w), 3)
    g = port.det_tensor((n, 64, d, h, w), 4)
    if cin >= 64:
        x = x.bfloat16().float()
    x_ref = x.clone().requires_grad_(True)
    y_ref = port.conv_block_sn(sd, '', x_ref, 1)
    y_ref.backward(g)
    x_gpu = x.cuda().requires_grad_(True)
    y = _run_block(m, x_gpu)
    assert rel_err(y, y_ref) < BF16_TOL
    y.backward(g.cuda())
    assert rel_err(x_gpu.grad, x_ref.grad) < LRELU_GRAD_TOL
    _compare_grads(m, sd)
    bufs = dict(m.named_buffers())
    assert rel_err(bufs['conv.weight_u'], sd['conv.weight_u']) < 1e-4
    assert rel_err(bufs['conv.weight_v'], sd['conv.weight_v']) < 1e-4


@pytest.mark.parametrize("dims", [2, 3])
def test_convblock2d_and_bare_heads(dims):
    """ConvBlock2D (networks_2d.py:53-61) and the bn=False, act=None form used by the mu / logvar heads"""
    from modules import networks_2d, networks_3d
    nets = networks_2d if dims == 2 else networks_3d
    cls = nets.ConvBlock2D if dims == 2 else nets.ConvBlock3D
    sp = (22, 26) if dims == 2 else (4, 12, 14)
    for bn, act, cout in ((True, 'lrelu', 64), (False, None, 128)):
        m = cls(64, cout, 3, 1, 1, bn=bn, act=act)
        sd = _leafs(_fill_module(m, 8))
        m.cuda()
        x = port.det_tensor((2, 64) + sp, 5).bfloat16().float()
        g = port.det_tensor((2, cout) + sp, 6)
        x_ref = x.clone().requires_grad_(True)
        y_ref = port.conv_block(sd, '', x_ref, 1)
        y_ref.backward(g)
        x_gpu = x.cuda().requires_grad_(True)
        y = _run_block(m, x_gpu)
        assert y.shape == y_ref.shape
        assert rel_err(y, y_ref) < BF16_TOL
        y.backward(g.cuda())
        assert rel_err(x_gpu.grad, x_ref.grad) < (LRELU_GRAD_TOL if act else BF16_TOL)
        _compare_grads(m, sd, LRELU_GRAD_TOL if act else BF16_TOL)


@pytest.mark.parametrize("cout,pad", [(3, 1), (1, 1), (3, 0)])
def test_tail_conv_layer(cout, pad):
    """the bare tail convolutions 64 -> nc_im / 64 -> 1 (networks_3d.py:175,341,362; pad 0 in GeneratorSG :290)"""
    from hpvg import ops
    w = port.det_tensor((cout, 64, 3, 3, 3), 7, scale=0.05)
    b = port.det_tensor((cout,), 8, scale=0.05)
    x = port.det_tensor((1, 64, 5, 14, 15), 9).bfloat16().float()
    x_ref, w_ref, b_ref = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y_ref = F.conv3d(x_ref, w_ref, b_ref, padding=pad)
    g = port.det_tensor(tuple(y_ref.shape), 10)
    y_ref.backward(g)
    xw = ops.ToWide.apply(x.cuda().requires_grad_(True))
    xw.retain_grad()
    wg, bg = w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    y = ops.conv(xw, wg, bg, pad, False)
    assert rel_err(y, y_ref) < 5e-3          # bf16 input and weights (tcgen05), fp32 accumulation and output
    y.backward(g.cuda())
    assert rel_err(wg.grad, w_ref.grad) < BF16_TOL
    assert rel_err(bg.grad, b_ref.grad) < 1e-4
    assert rel_err(ops.convert_raw(xw.grad, False), x_ref.grad) < BF16_TOL


@pytest.mark.parametrize("in_size,out_size", [((4, 12, 12), (4, 15, 15)), ((6, 54, 54), (16, 64, 64)), ((1, 33, 33), (1, 41, 41)),
                                              ((5, 20, 18), (3, 9, 31))])
def test_resize_matches_interpolate(in_size, out_size):
    """utils.upscale / interpolate_3D: trilinear, align_corners=True, fused '+ amp * noise' (utils/images.py:22-26,83-93)"""
    from hpvg import images
    x = port.det_tensor((2, 3) + in_size, 11)
    noise = port.det_tensor((2, 3) + out_size, 12)
    g = port.det_tensor((2, 3) + out_size, 13)
    x_ref = x.clone().requires_grad_(True)
    y_ref = port.resize(x_ref, out_size) + 0.37 * noise
    y_ref.backward(g)
    x_gpu = x.cuda().requires_grad_(True)
    y = images.resize(x_gpu, out_size, noise=noise.cuda(), amp=0.37)
    assert rel_err(y, y_ref) < 1e-5
    y.backward(g.cuda())
    assert rel_err(x_gpu.grad, x_ref.grad) < 1e-5
    assert rel_err(images.resize(x.cuda(), out_size), port.resize(x, out_size)) < 1e-5


def test_vae_head_kl_and_tanh():
    """reparameterize (networks_3d.py:29-35), kl_criterion (losses.py:7-9), tanh(a + b) (networks_3d.py:377,404)"""
    from hpvg import ops
    from modules.losses import kl_criterion
    from modules import networks_3d
    mu = port.det_tensor((1, 128, 4, 9, 10), 14, scale=0.7).bfloat16().float()
    logvar = port.det_tensor((1, 128, 4, 9, 10), 15, scale=1.5).bfloat16().float()
    eps = port.det_tensor((1, 128, 4, 9, 10), 16, scale=2.0)
    g = port.det_tensor((1, 128, 4, 9, 10), 17)
    mr, lr = mu.clone().requires_grad_(True), logvar.clone().requires_grad_(True)
    z_ref = eps.mul(lr.mul(0.5).exp()).add(mr)
    z_ref.backward(g.bfloat16().float())
    mg, lg = mu.cuda().requires_grad_(True), logvar.cuda().requires_grad_(True)
    z = ops.ToThin.apply(ops.Reparam.apply(ops.ToWide.apply(mg), ops.ToWide.apply(lg), eps.cuda()))
    assert rel_err(z, z_ref) < 4e-3           # one bf16 rounding of the result
    z.backward(g.cuda())
    assert rel_err(mg.grad, mr.grad) < 4e-3 and rel_err(lg.grad, lr.grad) < 4e-3
    # the module-level function of the reference API draws its own eps: compare in distribution only
    z2 = networks_3d.reparameterize(mu.cuda(), logvar.cuda(), True)
    assert z2.shape == mu.shape and abs(((z2.cpu() - mu) / logvar.mul(0.5).exp()).std().item() - 1.0) < 0.05

    mr, lr = mu.clone().requires_grad_(True), logvar.clone().requires_grad_(True)
    kl_ref = port.kl_criterion(mr, lr)
    kl_ref.backward()
    mg, lg = mu.cuda().requires_grad_(True), logvar.cuda().requires_grad_(True)
    kl = kl_criterion(mg, lg)
    assert abs(kl.item() - kl_ref.item()) < 1e-5 * abs(kl_ref.item())
    kl.backward()
    assert rel_err(mg.grad, mr.grad) < 1e-5 and rel_err(lg.grad, lr.grad) < 1e-5

    a, b = port.det_tensor((2, 3, 5, 11, 7), 18, scale=2.0), port.det_tensor((2, 3, 5, 11, 7), 19)
    ar, br = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    t_ref = torch.tanh(ar + br)
    gg = port.det_tensor((2, 3, 5, 11, 7), 20)
    t_ref.backward(gg)
    ag, bg = a.cuda().requires_grad_(True), b.cuda().requires_grad_(True)
    t = ops.TanhAdd.apply(ag, bg)
    assert rel_err(t, t_ref) < 1e-5
    t.backward(gg.cuda())
    assert rel_err(ag.grad, ar.grad) < 1e-4 and rel_err(bg.grad, br.grad) < 1e-4


def test_gradient_penalty_reduction():
    """the penalty term itself (modules/utils.py:18): lambda * mean((||g||_2 over channels - 1)^2), forward and backward"""
    from hpvg import ops
    g = port.det_tensor((2, 3, 4, 10, 9), 21, scale=1.3)
    gr = g.clone().requires_grad_(True)
    p_ref = ((gr.norm(2, dim=1) - 1) ** 2).mean() * 0.1
    p_ref.backward()
    gg = g.cuda().requires_grad_(True)
    p = ops.GpPenalty.apply(gg, 0.1)
    assert abs(p.item() - p_ref.item()) < 1e-5 * abs(p_ref.item())
    p.backward()
    assert rel_err(gg.grad, gr.grad) < 1e-5


# ---------------------------------------------------------------------------------------------------------------
# full BASELINE size (config 2 finest scale: N=1, 64 channels, 16 x 64 x 64): size-independent properties
# ---------------------------------------------------------------------------------------------------------------
def _wide_randn(shape, seed):
    gen = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(shape, device='cuda', generator=gen).to(torch.bfloat16)


def _dot(a, b):
    return (a.double() * b.double()).sum().item()


@pytest.mark.parametrize("cin,cout,vol", [(64, 64, (16, 64, 64)), (128, 64, (4, 32, 32)), (64, 128, (4, 32, 32)), (64, 64, (6, 54, 54))])
def test_full_size_conv_adjoint_identities(cin, cout, vol):
    from hpvg import ops
    d, h, w = vol
    x = _wide_randn((1, d, h, w, cin), 1)
    g = _wide_randn((1, d, h, w, cout), 2)
    gen = torch.Generator(device='cuda').manual_seed(3)
    wt = (torch.randn((cout, cin, 3, 3, 3), device='cuda', generator=gen) * 0.02).bfloat16().float()
    w2 = (torch.randn((cout, cin, 3, 3, 3), device='cuda', generator=gen) * 0.02).bfloat16().float()
    y = ops.conv_raw(x, wt, None, 1, False, True)
    gx = ops.conv_raw(g, wt, None, 1, True, True)
    lhs, rhs = _dot(y, g), _dot(x, gx)
    scale = (y.double().norm() * g.double().norm()).item()
    assert abs(lhs - rhs) < 2e-3 * scale, (lhs, rhs, scale)          # both sides carry one bf16 output rounding
    dw, db = ops.wgrad_raw(x, g, 1, tuple(wt.shape), want_bias=True)
    y2 = ops.conv_raw(x, w2, None, 1, False, True)
    lhs, rhs = _dot(dw, w2), _dot(y2, g)
    scale = (y2.double().norm() * g.double().norm()).item()
    assert abs(lhs - rhs) < 2e-3 * scale, (lhs, rhs, scale)
    assert rel_err(db, g.float().sum((0, 1, 2, 3))) < 1e-4
    # linearity in the weights (exact up to output rounding): conv(x, w) + conv(x, w2) = conv(x, w + w2)
    y12 = ops.conv_raw(x, (wt + w2).bfloat16().float(), None, 1, False, True)
    assert rel_err(y.float() + y2.float(), y12.float()) < 1e-2


def test_full_size_batchnorm_statistics():
    from modules import networks_3d
    torch.manual_seed(0)
    m = networks_3d.ConvBlock3D(64, 64, 3, 1, 1)
    m.cuda()
    x = _wide_randn((1, 16, 64, 64, 64), 4)
    from hpvg import ops
    stats = torch.zeros(128, device='cuda')
    y = ops.conv_raw(x, m.conv.weight.detach(), m.conv.bias.detach(), 1, False, True, stats=stats)
    yf = y.float().view(-1, 64)
    assert rel_err(stats[:64], yf.sum(0)) < 1e-4 and rel_err(stats[64:], (yf * yf).sum(0)) < 1e-4
    out = m.run(x).float().view(-1, 64)
    # LeakyReLU(0.2) of a zero-mean unit-variance channel: compare with the same transform of the normalised conv output
    ref = F.leaky_relu((yf - yf.mean(0)) / torch.sqrt(yf.var(0, unbiased=False) + 1e-5) * m.norm.weight.detach() + m.norm.bias.detach(), 0.2)
    assert rel_err(out, ref) < 5e-3


@pytest.mark.parametrize("cin,cout,vol,pad", [(64, 64, (37, 64, 64), 1), (64, 64, (9, 130, 70), 1), (128, 64, (12, 48, 40), 1),
                                              (64, 64, (22, 50, 50), 0), (64, 3, (37, 64, 64), 1), (64, 3, (6, 46, 46), 1),
                                              (64, 1, (2, 20, 33), 1), (64, 3, (10, 30, 30), 0)])
def test_tcgen05_kernels_with_several_units_per_cta(cin, cout, vol, pad):
    """volumes with more work units than SMs (every CTA loops over several units, with different zero-padding patterns per
    unit: BASELINE configs[4], 32 x 128 x 128, is such a case): tcgen05 kernels against the library's CUDA-core kernels"""
    from hpvg import lib, ops
    d, h, w = vol
    x = _wide_randn((1, d, h, w, cin), 11)
    gen = torch.Generator(device='cuda').manual_seed(12)
    wt = (torch.randn((cout, cin, 3, 3, 3), device='cuda', generator=gen) * 0.02).bfloat16().float()
    bias = torch.randn((cout,), device='cuda', generator=gen)
    wide_out = cout >= 64
    try:
        lib.set_conv_backend(lib.BACKEND_TCGEN05)
        y = ops.conv_raw(x, wt, bias, pad, False, wide_out)
        if wide_out:
            g = _wide_randn(tuple(y.shape), 13)
            gx = ops.conv_raw(g, wt, None, 2 - pad, True, True)
            dw, _ = ops.wgrad_raw(x, g, pad, tuple(wt.shape))
        lib.set_conv_backend(lib.BACKEND_DIRECT)
        y_ref = ops.conv_raw(x, wt, bias, pad, False, wide_out)
        if wide_out:
            gx_ref = ops.conv_raw(g, wt, None, 2 - pad, True, True)
            dw_ref, _ = ops.wgrad_raw(x, g, pad, tuple(wt.shape))
    finally:
        lib.set_conv_backend(lib.BACKEND_AUTO)
    assert rel_err(y.float(), y_ref.float()) < 3e-3          # both round the same fp32 sums to bf16 (thin output: fp32)
    if wide_out:
        assert rel_err(gx.float(), gx_ref.float()) < 3e-3
        assert rel_err(dw, dw_ref) < 1e-3


@pytest.mark.parametrize("col_mode", [0, 1])
@pytest.mark.parametrize("n,cout,vol,pad", [(1, 64, (16, 64, 64), 1), (2, 64, (5, 19, 21), 1), (1, 128, (7, 40, 24), 1), (1, 64, (11, 30, 17), 0),
                                            (1, 64, (19, 16, 8), 1), (1, 64, (1, 9, 9), 1), (1, 64, (37, 64, 64), 1)])
def test_both_tcgen05_conv_kernels_match_cuda_core(col_mode, n, cout, vol, pad):
    """The brick kernel (conv_tc.cu, col_mode 0) and the column-streaming kernel (conv_col.cu, col_mode 1) on the same
    64-input-channel layers — forward with bias + LeakyReLU + BatchNorm sums, data gradient with the fused LeakyReLU' mask —
    against the library's CUDA-core kernels: ragged bricks, batch 2, Cout 128, valid (pad 0) convolution and its pad-2 data
    gradient, columns longer than one 8-slice segment, a single d-slice, and the BASELINE config-2 volume."""
    from hpvg import lib, ops
    d, h, w = vol
    x = _wide_randn((n, d, h, w, 64), 21)
    gen = torch.Generator(device='cuda').manual_seed(22)
    wt = (torch.randn((cout, 64, 3, 3, 3), device='cuda', generator=gen) * 0.02).bfloat16().float()
    bias = torch.randn((cout,), device='cuda', generator=gen)
    res = {}
    prev = lib.set_conv_col_mode(col_mode)
    try:
        for backend in (lib.BACKEND_TCGEN05, lib.BACKEND_DIRECT):
            lib.set_conv_backend(backend)
            stats = torch.zeros(2 * cout, device='cuda')
            y = ops.conv_raw(x, wt, bias, pad, False, True, act_slope=0.2, stats=stats)
            y_plain = ops.conv_raw(x, wt, None, pad, False, True)
            g = _wide_randn(tuple(y.shape), 23)
            gx = ops.conv_raw(g, wt, None, 2 - pad, True, True)
            res[backend] = (y, y_plain, stats, gx)
            if cout == 64:
                msrc = _wide_randn(tuple(x.shape), 24)
                res[backend] += (ops.conv_raw(g, wt, None, 2 - pad, True, True, mask_src=msrc, mask_slope=0.2),)
    finally:
        lib.set_conv_backend(lib.BACKEND_AUTO)
        lib.set_conv_col_mode(prev)
    tc, ref = res[lib.BACKEND_TCGEN05], res[lib.BACKEND_DIRECT]
    assert rel_err(tc[0].float(), ref[0].float()) < 3e-3
    assert rel_err(tc[1].float(), ref[1].float()) < 3e-3
    yf = tc[0].float().reshape(-1, cout)
    assert rel_err(tc[2][:cout], yf.sum(0)) < 1e-4 and rel_err(tc[2][cout:], (yf * yf).sum(0)) < 1e-4
    assert rel_err(tc[3].float(), ref[3].float()) < 3e-3
    if cout == 64:
        assert rel_err(tc[4].float(), ref[4].float()) < 3e-3
