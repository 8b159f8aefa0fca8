"""The tcgen05 kernels against the CPU reference arithmetic AT BASELINE.json's shapes.

configs[1] (train_video.py, 16 frames 64 x 64, nfc 64) runs its 64 -> 64 layers on 1 x 64 x 16 x 64 x 64 (finest level) and
1 x 64 x 6 x 54 x 54 (the level below it).  Every kernel that serves those layers — the brick kernel (conv_tc.cu), the
column-streaming kernel (conv_col.cu), the thin-output tail kernel, the weight-gradient kernel (wgrad_tc.cu) and the
3-channel-end kernels (narrow.cu) — is compared here with PyTorch's CPU convolution (the arithmetic the reference's
nn.Conv3d resolves to, modules/networks_3d.py:51,63,175,341,362) and its functional adjoints
torch.nn.grad.conv3d_input / conv3d_weight, on identical bf16-representable operands.

Tolerances (north_star): 2e-2 relative for bf16 operands is the bar; what is measured here is far tighter because both
sides multiply the same bf16 values and accumulate in fp32 — the only difference is summation order and ONE rounding of
the stored result (bf16: 2^-9 relative per element -> 2.3e-3 in L2; fp32 outputs: 1e-4).
"""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_err

pytestmark = pytest.mark.gpu

WIDE_OUT_TOL = 3e-3      # one bf16 rounding of the stored result
F32_OUT_TOL = 2e-4       # fp32 results (weight gradients, thin outputs): summation order only

VOLUMES = [(16, 64, 64), (6, 54, 54)]


def _bf16_values(shape, seed, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=gen) * scale).bfloat16().float()


def _to_wide(t):
    """float32 NCDHW (CPU) -> bf16 NDHWC (cuda), exact for bf16-representable values"""
    return t.permute(0, 2, 3, 4, 1).contiguous().to(device='cuda', dtype=torch.bfloat16)


def _to_ncdhw(t):
    return t.float().permute(0, 4, 1, 2, 3).contiguous().cpu()


@pytest.mark.parametrize("col_mode", [0, 1], ids=["brick", "column"])
@pytest.mark.parametrize("vol", VOLUMES, ids=["16x64x64", "6x54x54"])
def test_conv_64_to_64_against_cpu_conv3d(vol, col_mode):
    """forward (+bias, BatchNorm sums), forward with fused LeakyReLU, and the data gradient of the 64 -> 64 layer"""
    from hpvg import lib, ops
    d, h, w = vol
    x = _bf16_values((1, 64, d, h, w), 1)
    wt = _bf16_values((64, 64, 3, 3, 3), 2, 0.03)
    bias = _bf16_values((64,), 3, 0.1)
    g = _bf16_values((1, 64, d, h, w), 4)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    y_ref = F.conv3d(x, wt, bias, padding=1)
    gx_ref = torch.nn.grad.conv3d_input(x.shape, wt, g, padding=1)
    prev = lib.set_conv_col_mode(col_mode)
    try:
        lib.set_conv_backend(lib.BACKEND_TCGEN05)
        stats = torch.zeros(128, device='cuda')
        y = ops.conv_raw(_to_wide(x), wt.cuda(), bias.cuda(), 1, False, True, stats=stats)
        y_act = ops.conv_raw(_to_wide(x), wt.cuda(), bias.cuda(), 1, False, True, act_slope=0.2)
        gx = ops.conv_raw(_to_wide(g), wt.cuda(), None, 1, True, True)
    finally:
        lib.set_conv_backend(lib.BACKEND_AUTO)
        lib.set_conv_col_mode(prev)
    assert rel_err(_to_ncdhw(y), y_ref) < WIDE_OUT_TOL
    assert rel_err(_to_ncdhw(y_act), F.leaky_relu(y_ref, 0.2)) < WIDE_OUT_TOL
    assert rel_err(_to_ncdhw(gx), gx_ref) < WIDE_OUT_TOL
    # BatchNorm sums are taken over the values the kernel stores (bf16): against the fp32 reference they carry the rounding
    # noise of V elements, which averages out
    assert rel_err(stats[:64].cpu(), y_ref.sum((0, 2, 3, 4))) < 2e-3
    assert rel_err(stats[64:].cpu(), (y_ref * y_ref).sum((0, 2, 3, 4))) < 1e-3


@pytest.mark.parametrize("vol", VOLUMES, ids=["16x64x64", "6x54x54"])
def test_wgrad_64_to_64_against_cpu_conv3d_weight(vol):
    """hpvg_conv_wgrad (wgrad_tc.cu) against torch.nn.grad.conv3d_weight, and the fused bias gradient"""
    from hpvg import lib, ops
    d, h, w = vol
    x = _bf16_values((1, 64, d, h, w), 5)
    g = _bf16_values((1, 64, d, h, w), 6)
    dw_ref = torch.nn.grad.conv3d_weight(x, (64, 64, 3, 3, 3), g, padding=1)
    try:
        lib.set_conv_backend(lib.BACKEND_TCGEN05)
        dw, db = ops.wgrad_raw(_to_wide(x), _to_wide(g), 1, (64, 64, 3, 3, 3), want_bias=True)
    finally:
        lib.set_conv_backend(lib.BACKEND_AUTO)
    assert rel_err(dw.cpu(), dw_ref) < F32_OUT_TOL
    assert rel_err(db.cpu(), g.sum((0, 2, 3, 4))) < F32_OUT_TOL


@pytest.mark.parametrize("cout", [3, 1])
@pytest.mark.parametrize("vol", VOLUMES, ids=["16x64x64", "6x54x54"])
def test_tail_conv_against_cpu_conv3d(vol, cout):
    """the 64 -> 3 generator tail and the 64 -> 1 critic tail (thin fp32 output), their data gradient (thin -> wide) and weight
    gradient (narrow.cu)"""
    from hpvg import ops
    d, h, w = vol
    x = _bf16_values((1, 64, d, h, w), 7)
    wt = _bf16_values((cout, 64, 3, 3, 3), 8, 0.03)
    bias = _bf16_values((cout,), 9, 0.1)
    g = _bf16_values((1, cout, d, h, w), 10)
    y_ref = F.conv3d(x, wt, bias, padding=1)
    gx_ref = torch.nn.grad.conv3d_input(x.shape, wt, g, padding=1)
    dw_ref = torch.nn.grad.conv3d_weight(x, tuple(wt.shape), g, padding=1)
    y = ops.conv_raw(_to_wide(x), wt.cuda(), bias.cuda(), 1, False, False)
    gx = ops.conv_raw(g.cuda(), wt.cuda(), None, 1, True, True)
    dw, db = ops.wgrad_raw(_to_wide(x), g.cuda(), 1, tuple(wt.shape), want_bias=True)
    assert rel_err(y.cpu(), y_ref) < F32_OUT_TOL
    # the thin -> wide kernel multiplies in TF32 (10-bit mantissa: exact for bf16 weights, g rounded from fp32... g is bf16-exact here)
    assert rel_err(_to_ncdhw(gx), gx_ref) < WIDE_OUT_TOL
    assert rel_err(dw.cpu(), dw_ref) < 1e-3
    assert rel_err(db.cpu(), g.sum((0, 2, 3, 4))) < F32_OUT_TOL


@pytest.mark.parametrize("vol", VOLUMES, ids=["16x64x64", "6x54x54"])
def test_head_conv_against_cpu_conv3d(vol):
    """the 3 -> 64 head (thin fp32 input, wide output), its data gradient (wide -> thin) and weight gradient"""
    from hpvg import ops
    d, h, w = vol
    x = _bf16_values((1, 3, d, h, w), 11)
    wt = _bf16_values((64, 3, 3, 3, 3), 12, 0.1)
    bias = _bf16_values((64,), 13, 0.1)
    g = _bf16_values((1, 64, d, h, w), 14)
    y_ref = F.conv3d(x, wt, bias, padding=1)
    gx_ref = torch.nn.grad.conv3d_input(x.shape, wt, g, padding=1)
    dw_ref = torch.nn.grad.conv3d_weight(x, tuple(wt.shape), g, padding=1)
    stats = torch.zeros(128, device='cuda')
    y = ops.conv_raw(x.cuda(), wt.cuda(), bias.cuda(), 1, False, True, stats=stats)
    gx = ops.conv_raw(_to_wide(g), wt.cuda(), None, 1, True, False)
    dw, db = ops.wgrad_raw(x.cuda(), _to_wide(g), 1, tuple(wt.shape), want_bias=True)
    assert rel_err(_to_ncdhw(y), y_ref) < WIDE_OUT_TOL
    assert rel_err(gx.cpu(), gx_ref) < F32_OUT_TOL
    assert rel_err(dw.cpu(), dw_ref) < 1e-3
    assert rel_err(db.cpu(), g.sum((0, 2, 3, 4))) < F32_OUT_TOL
    assert rel_err(stats[:64].cpu(), y_ref.sum((0, 2, 3, 4))) < 2e-3


# ---------------------------------------------------------------------------------------------------------------
# ConvBlock3D in one launch (hpvg_conv_bn_lrelu_fused): conv + BatchNorm(batch statistics) + LeakyReLU, forward and backward,
# against PyTorch's CPU fp32 arithmetic for the same layer (reference modules/networks_3d.py:48-56)
# ---------------------------------------------------------------------------------------------------------------
BF16_TOL = 2e-2          # north_star: relative 2e-2 for bf16 operands — forward AND gradients of the fused layer


def _convblock_reference(x, wt, bias, gamma, beta, g):
    x = x.clone().requires_grad_(True)
    prm = [t.clone().requires_grad_(True) for t in (wt, bias, gamma, beta)]
    y = F.conv3d(x, prm[0], prm[1], padding=1)
    out = F.leaky_relu(F.batch_norm(y, None, None, prm[2], prm[3], True, 0.1, 1e-5), 0.2)
    out.backward(g)
    return out.detach(), x.grad, [t.grad for t in prm], y.detach()


@pytest.mark.parametrize("vol", VOLUMES + [(4, 32, 32), (4, 39, 39), (6, 46, 46), (13, 64, 64), (3, 20, 9)],
                         ids=["16x64x64", "6x54x54", "4x32x32", "4x39x39", "6x46x46", "13x64x64", "3x20x9"])
def test_fused_convblock_against_cpu_reference(vol):
    from hpvg import lib, ops
    d, h, w = vol
    assert lib.load().hpvg_conv_bn_lrelu_fused_supported(1, 64, 64, d, h, w, 3, 1) == 1
    x = _bf16_values((1, 64, d, h, w), 21)
    wt = _bf16_values((64, 64, 3, 3, 3), 22, 0.03)
    bias = _bf16_values((64,), 23, 0.1)
    gamma = 1.0 + _bf16_values((64,), 24, 0.1)
    beta = _bf16_values((64,), 25, 0.1)
    g = _bf16_values((1, 64, d, h, w), 26)
    out_ref, gx_ref, (gw_ref, gb_ref, gg_ref, gbeta_ref), y_ref = _convblock_reference(x, wt, bias, gamma, beta, g)

    def run(fused):
        xg = _to_wide(x).requires_grad_(True)
        prm = [t.cuda().requires_grad_(True) for t in (wt, bias, gamma, beta)]
        rm, rv, nbt = torch.zeros(64, device='cuda'), torch.ones(64, device='cuda'), torch.zeros((), dtype=torch.int64, device='cuda')
        launches = lib.launch_count()
        with ops.fused_bn(fused):
            out = ops.conv_bn_lrelu(xg, prm[0], prm[1], prm[2], prm[3], rm, rv, nbt, 1)
        fwd_launches = lib.launch_count() - launches
        out.backward(_to_wide(g))
        return out.detach(), xg.grad, [t.grad for t in prm], (rm, rv, nbt), fwd_launches

    out, gx, (gw, gb, gg, gbeta), (rm, rv, nbt), n_fused = run(True)
    out_u, gx_u, (gw_u, _, gg_u, gbeta_u), (rm_u, rv_u, _), n_unfused = run(False)
    assert n_fused < n_unfused            # pack + ONE fused launch against pack + conv + BatchNorm apply
    # forward: the activated output, the running statistics
    assert rel_err(_to_ncdhw(out), out_ref) < 5e-3
    ym = y_ref.mean((0, 2, 3, 4))
    assert rel_err(rm.cpu(), 0.1 * ym) < 1e-3
    assert rel_err(rv.cpu() - 0.9, 0.1 * y_ref.var((0, 2, 3, 4), unbiased=True)) < 1e-3
    assert int(nbt.item()) == 1
    # backward at the bf16-operand tolerance: input gradient, weight gradient, BatchNorm scale / shift gradients
    assert rel_err(_to_ncdhw(gx), gx_ref) < BF16_TOL
    assert rel_err(gw.cpu(), gw_ref) < BF16_TOL
    assert rel_err(gg.cpu(), gg_ref) < BF16_TOL
    assert rel_err(gbeta.cpu(), gbeta_ref) < BF16_TOL
    # the conv bias in front of BatchNorm has a mathematically zero gradient: on the scale of the other gradients
    assert gb.abs().max().item() <= 1e-2 * gbeta_ref.abs().max().item() + 1e-3
    # the fused launch is closer to the fp32 reference than the two-launch path (which re-reads a bf16 conv output)
    assert rel_err(_to_ncdhw(out), out_ref) <= rel_err(_to_ncdhw(out_u), out_ref) * 1.05
    assert rel_err(_to_ncdhw(gx), gx_ref) <= rel_err(_to_ncdhw(gx_u), gx_ref) * 1.05
    assert rel_err(out.float(), out_u.float()) < 1e-2 and rel_err(rm, rm_u) < 1e-3 and rel_err(rv, rv_u) < 1e-3
    print("fused vs fp32: out %.2e gx %.2e gw %.2e | two-launch: out %.2e gx %.2e gw %.2e" % (
        rel_err(_to_ncdhw(out), out_ref), rel_err(_to_ncdhw(gx), gx_ref), rel_err(gw.cpu(), gw_ref),
        rel_err(_to_ncdhw(out_u), out_ref), rel_err(_to_ncdhw(gx_u), gx_ref), rel_err(gw_u.cpu(), gw_ref)))


@pytest.mark.parametrize("vol,c", [((16, 64, 64), 64), ((6, 54, 54), 64), ((3, 20, 9), 64), ((4, 17, 13), 8), ((2, 9, 11), 128)],
                         ids=["16x64x64", "6x54x54", "3x20x9", "8ch", "128ch"])
def test_one_launch_batchnorm_backward_equals_the_two_launch_pair(vol, c):
    """hpvg_bn_lrelu_bwd_fused (y and gout read once into shared memory, grid barrier) against hpvg_bn_lrelu_bwd_reduce + _apply:
    same formulas on the same values, so the results agree to fp32 summation order; with and without the sign-bit mask, with the
    fused bias-gradient sum"""
    from hpvg import lib
    from hpvg.ops import _ptr, _stream
    d, h, w = vol
    nvox = d * h * w
    assert lib.load().hpvg_bn_lrelu_bwd_fused_supported(nvox, c) == 1
    gen = torch.Generator(device='cuda').manual_seed(31)
    y = torch.randn((nvox, c), device='cuda', generator=gen).bfloat16()
    gout = torch.randn((nvox, c), device='cuda', generator=gen).bfloat16()
    scale_shift = torch.cat([1.0 + 0.1 * torch.randn(c, device='cuda', generator=gen), 0.1 * torch.randn(c, device='cuda', generator=gen)])
    mean_invstd = torch.cat([0.05 * torch.randn(c, device='cuda', generator=gen), 1.0 + 0.1 * torch.rand(c, device='cuda', generator=gen)])
    for use_mask in (False, True):
        mask = torch.randint(0, 256, (nvox * c // 8,), device='cuda', dtype=torch.uint8, generator=gen) if use_mask else None
        res = {}
        for fused in (False, True):
            gy = torch.empty_like(y)
            dgamma, dbeta = torch.empty(c, device='cuda'), torch.empty(c, device='cuda')
            if fused:
                sums = torch.zeros(3 * c + 32, device='cuda')
                lib.call("hpvg_bn_lrelu_bwd_fused", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), _ptr(gy),
                         _ptr(dgamma), _ptr(dbeta), nvox, c, 0.2, 1, _ptr(mask), _stream())
            else:
                sums = torch.empty(3 * c, device='cuda')
                lib.call("hpvg_bn_lrelu_bwd_reduce", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), nvox, c, 0.2,
                         _ptr(mask), _stream())
                lib.call("hpvg_bn_lrelu_bwd_apply", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), _ptr(gy),
                         _ptr(dgamma), _ptr(dbeta), nvox, c, 0.2, 1, _ptr(mask), _stream())
            torch.cuda.synchronize()
            res[fused] = (gy.float(), dgamma.clone(), dbeta.clone(), sums[2 * c:3 * c].clone())
        a, b = res[True], res[False]
        assert rel_err(a[1], b[1]) < 1e-4 and rel_err(a[2], b[2]) < 1e-4          # dgamma, dbeta: summation order
        assert rel_err(a[0], b[0]) < 2e-3                                         # gy: bf16 roundings flipped by 1-ulp mean differences
        assert (a[3] - b[3]).abs().max().item() <= 2e-3 * b[0].abs().sum(0).max().item() + 1e-3    # bias-gradient sums
