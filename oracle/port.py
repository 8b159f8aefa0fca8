"""Functional CPU restatement of the reference's per-scale hot path (TEST INFRASTRUCTURE — see oracle/__init__.py).

Everything is a pure function of a state_dict `sd` (same keys as the reference modules) so that the
very same weights can be loaded into the reference modules (fixture generation), into this port and into the CUDA
drop-in.  Buffers that the reference mutates in forward (BatchNorm running statistics, spectral-norm u/v) are
updated in place in `sd` exactly as the reference does.  2-D vs 3-D is decided by the weight rank.

Citations are to files under the reference repository root.
"""
import math

import os

import torch
import torch.nn.functional as F

SLOPE = 0.2          # nn.LeakyReLU(0.2), modules/networks_3d.py:21
BN_EPS = 1e-5        # nn.BatchNorm3d default
BN_MOMENTUM = 0.1
SN_EPS = 1e-12       # torch.nn.utils.spectral_norm default


# ------------------------------------------------------------------------------------------------------------------
# storage-precision emulation (test aid): the CUDA path keeps every 64/128-channel activation AND its gradient in
# bfloat16 between kernels (fp32 accumulation inside).  `with storage('bf16'):` rounds at exactly those points, so a
# CUDA result can be compared with (a) the fp32 reference at the north-star tolerance and (b) this emulation at a
# much tighter one — (b) separates "bf16 rounding" from "bug".
# ------------------------------------------------------------------------------------------------------------------
class _RoundBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.to(torch.bfloat16).to(t.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


_STORAGE = ['f32']


class storage(object):
    def __init__(self, mode):
        assert mode in ('f32', 'bf16')
        self.mode = mode

    def __enter__(self):
        _STORAGE.append(self.mode)

    def __exit__(self, *a):
        _STORAGE.pop()


def store(t):
    """a tensor the CUDA path holds in bf16 (value and gradient)"""
    return _RoundBf16.apply(t) if _STORAGE[-1] == 'bf16' else t


class _RoundGradBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def store_grad(t):
    """a tensor whose value never leaves the kernel but whose gradient is materialised in bf16 (the pre-activation of
    a fused conv + LeakyReLU: its gradient is the output of the leaky_relu_backward kernel)"""
    return _RoundGradBf16.apply(t) if _STORAGE[-1] == 'bf16' else t


def _tf32(t):
    """round to TF32 (10 explicit mantissa bits), as cvt.rna.tf32.f32 does for the operands of the 3 -> 64 head layers"""
    bits = t.detach().contiguous().view(torch.int32)
    return ((bits + 0x1000) & ~0x1FFF).view(torch.float32)


# operand precision of the thin -> wide layers: 'bf16' = the tcgen05 kernel (csrc/narrow_tc.cu, the default), 'tf32' = the
# mma.sync kernel it replaced (HPVG_EXPAND_TC=0)
HEAD_OPERAND = ['tf32' if os.environ.get('HPVG_EXPAND_TC') == '0' else 'bf16']


def head_operand(t, w):
    """operands (input or weight) of a thin -> wide layer (Cin <= 3, Cout = 64): the CUDA path multiplies them on the
    tensor cores in bf16 (TF32 with HPVG_EXPAND_TC=0); gradients pass straight through"""
    if _STORAGE[-1] == 'bf16' and t.dtype == torch.float32 and w.shape[1] <= 3 and w.shape[0] == 64:
        rounded = _tf32(t) if HEAD_OPERAND[0] == 'tf32' else t.detach().to(torch.bfloat16).to(t.dtype)
        return t + (rounded - t).detach()
    return t


def mma_weight(w):
    """weights of the tcgen05 layers (Cin in {64,128} with Cout a multiple of 64, and the 64 -> nc_im / 64 -> 1 tails)
    are fed to the tensor core in bf16; the master weight and its gradient stay fp32 (straight-through)"""
    if _STORAGE[-1] == 'bf16' and ((w.shape[1] in (64, 128) and w.shape[0] % 64 == 0) or (w.shape[1] == 64 and w.shape[0] <= 16)):
        return w + (w.to(torch.bfloat16).to(w.dtype) - w).detach()
    return w


# ------------------------------------------------------------------------------------------------------------------
# primitives
# ------------------------------------------------------------------------------------------------------------------
def conv(x, w, b, pad):
    """nn.Conv3d / nn.Conv2d forward, kernel 3, stride 1 (modules/networks_3d.py:51, networks_2d.py:56)"""
    return F.conv3d(x, w, b, padding=pad) if w.dim() == 5 else F.conv2d(x, w, b, padding=pad)


def batch_norm_train(sd, prefix, y):
    """nn.BatchNorm3d/2d in training mode (modules/networks_3d.py:54): biased batch variance for normalisation,
    running stats updated with momentum 0.1 and the unbiased variance, num_batches_tracked += 1."""
    dims = [0] + list(range(2, y.dim()))
    count = y.numel() // y.shape[1]
    mean = y.mean(dims)
    var = y.var(dims, unbiased=False)
    shape = [1, -1] + [1] * (y.dim() - 2)
    out = (y - mean.view(shape)) / torch.sqrt(var.view(shape) + BN_EPS)
    out = out * sd[prefix + 'weight'].view(shape) + sd[prefix + 'bias'].view(shape)
    with torch.no_grad():
        rm, rv = sd[prefix + 'running_mean'], sd[prefix + 'running_var']
        rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach())
        unbiased = var.detach() * (count / max(count - 1, 1))
        rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * unbiased)
        sd[prefix + 'num_batches_tracked'] += 1
    return out


FUSED_BN_SMS = [148]     # storage emulation: SM count of the device whose fused-layer rule is modelled (0: no fused layers)


def fused_bn_layer(x, w, pad):
    """storage emulation only: would the CUDA path run this ConvBlock as ONE launch (hpvg_conv_bn_lrelu_fused)?  Same rule as
    conv_tc.cu::fused_nacc: 64 -> 64 3x3x3 layers whose 16 x 8-voxel bricks x (2 or 4)-slice units fit the SMs.  Such a layer
    takes BatchNorm statistics, normalisation and the LeakyReLU sign from the fp32 accumulators; the bf16 copy of the conv
    output exists only for the backward pass."""
    if w.dim() != 5 or tuple(w.shape[:2]) != (64, 64) or x.dim() != 5 or FUSED_BN_SMS[0] <= 0:
        return False
    n, _, d, h, wd = x.shape
    do, ho, wo = d + 2 * pad - 2, h + 2 * pad - 2, wd + 2 * pad - 2
    per_slice = n * -(-ho // 16) * -(-wo // 8)
    if do % 2 == 0 and per_slice * (do // 2) <= FUSED_BN_SMS[0]:
        return True
    return per_slice * -(-do // 4) <= FUSED_BN_SMS[0]


def conv_block(sd, prefix, x, pad):
    """ConvBlock3D/2D: conv [+ BatchNorm] [+ LeakyReLU]; the mu/logvar heads have neither (networks_3d.py:48-56, :99-100)"""
    w = sd[prefix + 'conv.weight']
    y = conv(head_operand(x, w), head_operand(mma_weight(w), w), sd[prefix + 'conv.bias'], pad)
    if prefix + 'norm.weight' in sd:
        y = store_grad(y) if (_STORAGE[-1] == 'bf16' and fused_bn_layer(x, w, pad)) else store(y)
        y = store(F.leaky_relu(batch_norm_train(sd, prefix + 'norm.', y), SLOPE))
    else:
        y = store(y)
    return y


def spectral_weight(sd, prefix, training=True):
    """legacy nn.utils.spectral_norm pre-forward hook (networks_3d.py:63): one power iteration on W viewed [Cout, -1]
    (u, v updated in place, no grad), sigma = u^T W v, weight = weight_orig / sigma."""
    w = sd[prefix + 'weight_orig']
    u, v = sd[prefix + 'weight_u'], sd[prefix + 'weight_v']
    wm = w.reshape(w.shape[0], -1)
    if training:
        with torch.no_grad():
            v.copy_(F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS))
            u.copy_(F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS))
    u_, v_ = u.clone(), v.clone()
    sigma = torch.dot(u_, torch.mv(wm, v_))
    return w / sigma


def conv_block_sn(sd, prefix, x, pad):
    """ConvBlock3DSN/2DSN with bn=True: spectral-norm conv + LeakyReLU, no BatchNorm (networks_3d.py:59-70)"""
    w = mma_weight(spectral_weight(sd, prefix + 'conv.'))
    y = conv(head_operand(x, w), head_operand(w, w), sd[prefix + 'conv.bias'], pad)
    return store(F.leaky_relu(store_grad(y), SLOPE))


def resize(x, size):
    """utils.interpolate_3D / utils.interpolate: F.interpolate(align_corners=True) (utils/images.py:9-26)"""
    mode = 'trilinear' if x.dim() == 5 else 'bilinear'
    return F.interpolate(x, size=list(size), mode=mode, align_corners=True)


def scale_size(index, opt):
    """utils.get_scales_by_index (utils/images.py:60-64)"""
    return math.ceil(math.pow(opt.scale_factor, opt.stop_scale - index) * opt.img_size)


def time_depth(index, opt):
    """utils.get_fps_td_by_index (utils/images.py:67-80)"""
    fps_index = int((index / opt.stop_scale_time) * (len(opt.sampling_rates) - 1))
    return opt.fps_lcm // opt.sampling_rates[fps_index] + 1


def upscale(x, index, opt):
    """utils.upscale (5-D, utils/images.py:83-93) / utils.upscale_2d (4-D, :96-105)"""
    assert index > 0
    s = scale_size(index, opt)
    if x.dim() == 5:
        return resize(x, [time_depth(index, opt), int(s * opt.ar), s])
    return resize(x, [int(s * opt.ar), s])


def kl_criterion(mu, logvar):
    """modules/losses.py:7-9"""
    return (-0.5 * (1 + logvar - mu.pow(2) - logvar.exp())).mean()


# ------------------------------------------------------------------------------------------------------------------
# networks
# ------------------------------------------------------------------------------------------------------------------
def _half(opt):
    return opt.ker_size // 2


def encode(sd, opt, x, prefix='encode.'):
    """Encode3DVAE / Encode2DVAE forward (networks_3d.py:88-107): enc_blocks + 1 SN blocks, then mu and logvar convs"""
    h = x
    for i in range(opt.enc_blocks + 1):
        h = conv_block_sn(sd, '%sfeatures.conv_block_%d.' % (prefix, i), h, _half(opt))
    return conv_block(sd, prefix + 'mu.', h, _half(opt)), conv_block(sd, prefix + 'logvar.', h, _half(opt))


def stage(sd, opt, prefix, x, pad, tail_pad):
    """head ConvBlock + num_layer ConvBlocks + bare tail conv (decoder :337-341, refinement stage :355-362)"""
    h = conv_block(sd, prefix + 'head.', x, pad)
    for i in range(opt.num_layer):
        h = conv_block(sd, '%sblock%d.' % (prefix, i), h, pad)
    return conv(h, mma_weight(sd[prefix + 'tail.weight']), sd[prefix + 'tail.bias'], tail_pad)


def num_body(sd):
    k = 0
    while 'body.%d.head.conv.weight' % k in sd:
        k += 1
    return k


def generator(sd, opt, video, noise_amp, noise_init=None, mode='rand', eps=None, noises=None, sample_init=None):
    """GeneratorHPVAEGAN.forward + refinement_layers (networks_3d.py:367-406, networks_2d.py:230-269).

    `eps` (reparameterisation noise) and `noises` ({level: tensor}) may be supplied to make runs comparable across
    devices; when None they are drawn with torch in the reference's order (zeros_like(...).normal_()).
    The 3-D generator adds noise only at levels >= vae_levels; the 2-D one at every level in 'rand' mode.
    """
    three_d = sd['decoder.tail.weight'].dim() == 5
    if noise_init is None:
        mu, logvar = encode(sd, opt, video)
        std = logvar.mul(0.5).exp()                                   # reparameterize, networks_3d.py:29-35
        if eps is None:
            eps = torch.zeros_like(std).normal_()
        z = store(eps.mul(std).add(mu))
    else:
        z = store(noise_init)
    vae_out = torch.tanh(stage(sd, opt, 'decoder.', z, opt.padd_size, _half(opt)))
    start, x = (0, vae_out) if sample_init is None else sample_init
    for idx in range(start, num_body(sd)):
        if opt.vae_levels == idx + 1 and not opt.train_all:
            x = x.detach()          # the reference detaches in place (:391-392); value-wise identical
            if idx == 0 and sample_init is None:
                vae_out = x         # in-place detach also hits the returned vae_out when it is the same tensor
        x_up = upscale(x, idx + 1, opt)
        if mode == 'rand' and (not three_d or opt.vae_levels <= idx + 1):
            noise = noises[idx + 1] if noises is not None else torch.zeros_like(x_up).normal_(0, 1)
            x_in = x_up + noise * noise_amp[idx + 1]
        else:
            x_in = x_up
        x = torch.tanh(stage(sd, opt, 'body.%d.' % idx, x_in, opt.padd_size, _half(opt)) + x_up)
    if noise_init is None:
        return x, vae_out, (mu, logvar)
    return x, vae_out


def discriminator(sd, opt, x):
    """WDiscriminator3D/2D.forward (networks_3d.py:163-181): SN head, num_layer SN blocks, plain tail conv (padding=1)"""
    h = conv_block_sn(sd, 'head.', x, _half(opt))
    for i in range(opt.num_layer):
        h = conv_block_sn(sd, 'body.block%d.' % i, h, _half(opt))
    return conv(h, mma_weight(sd['tail.weight']), sd['tail.bias'], 1)


def gradient_penalty(sd_d, opt, real, fake, lam, alpha=None):
    """calc_gradient_penalty (modules/utils.py:4-19). `alpha` may be supplied; otherwise torch.rand(1, 1) (CPU generator)."""
    if alpha is None:
        alpha = float(torch.rand(1, 1))
    interp = (alpha * real + (1 - alpha) * fake).detach().requires_grad_(True)
    out = discriminator(sd_d, opt, interp)
    grads = torch.autograd.grad(outputs=out, inputs=interp, grad_outputs=torch.ones_like(out), create_graph=True,
                                retain_graph=True, only_inputs=True)[0]
    return ((grads.norm(2, dim=1) - 1) ** 2).mean() * lam


def generator_sg(sd, opt, noise_init, noise_amp, mode='rand', noises=None):
    """GeneratorSG.forward (networks_3d.py:298-322): pad-0 stages on inputs zero-padded by num_layer + 2 voxels"""
    m = opt.num_layer + 2
    p3d = (m,) * 6
    x = stage(sd, opt, 'body.0.', F.pad(noise_init, p3d), 0, 0)
    for idx in range(1, num_body(sd)):
        x = torch.tanh(x)
        x_up = upscale(x, idx, opt)
        if mode == 'rand':
            big = resize(x, [s + 2 * m for s in x_up.shape[-3:]])
            noise = noises[idx] if noises is not None else torch.zeros_like(big).normal_(0, 1)
            x_prev = stage(sd, opt, 'body.%d.' % idx, big + noise * noise_amp[idx], 0, 0)
        else:
            x_prev = stage(sd, opt, 'body.%d.' % idx, F.pad(x_up, p3d), 0, 0)
        x = x_prev + x_up
    return torch.tanh(x)


def generator_csg(sd, opt, noise_init, noise_amp, mode='rand', noises=None):
    """GeneratorCSG.forward (networks_3d.py:246-269): BatchNorm head on the noise padded by 1, stages of num_layer pad-0
    ConvBlocks on inputs zero-padded by num_layer voxels, chained in feature space (resize, + noise, residual add), one
    conv + tanh tail on the result padded by 1"""
    m = opt.num_layer

    def body(k, h):
        for i in range(opt.num_layer):
            h = conv_block(sd, 'body.%d.block%d.' % (k, i), h, 0)
        return h

    x = conv_block(sd, 'head.', F.pad(noise_init, (1,) * 6), 0)
    x = body(0, F.pad(x, (m,) * 6))
    k = 1
    while 'body.%d.block0.conv.weight' % k in sd:
        x_up = store(upscale(x, k, opt))
        if mode == 'rand':
            big = resize(x, [s + 2 * m for s in x_up.shape[-3:]])
            noise = noises[k] if noises is not None else torch.zeros_like(big).normal_(0, 1)
            x_in = store(big + noise * noise_amp[k])
        else:
            x_in = F.pad(x_up, (m,) * 6)
        x = store(body(k, x_in) + x_up)
        k += 1
    return torch.tanh(conv(F.pad(x, (1,) * 6), mma_weight(sd['tail.0.weight']), sd['tail.0.bias'], 0))


def discriminator_baselines(sd, opt, x):
    """WDiscriminatorBaselines.forward (networks_3d.py:204-210): input zero-padded by num_layer + 2, conv + LeakyReLU head,
    num_layer ConvBlocks (BatchNorm), plain tail conv, all with padding opt.padd_size"""
    m = opt.num_layer + 2
    x = F.pad(x, (m,) * 6)
    w = sd['head.conv.weight']
    h = conv(head_operand(x, w), head_operand(mma_weight(w), w), sd['head.conv.bias'], opt.padd_size)
    h = store(F.leaky_relu(store_grad(h), SLOPE))
    for i in range(opt.num_layer):
        h = conv_block(sd, 'body.block%d.' % i, h, opt.padd_size)
    return conv(h, mma_weight(sd['tail.weight']), sd['tail.bias'], opt.padd_size)


# ------------------------------------------------------------------------------------------------------------------
# deterministic weights and options (shared by the fixture generator, the CPU tests and the GPU tests)
# ------------------------------------------------------------------------------------------------------------------
def _hash_uniform(n, seed):
    """n reproducible uniforms in [-1, 1): a 64-bit integer mixer on the element index (exact on every machine)"""
    idx = torch.arange(n, dtype=torch.int64) + int(seed) * 1000003
    h = idx * -7046029254386353131            # 0x9E3779B97F4A7C15 as int64 (wrap-around multiply)
    h = h ^ ((h >> 29) & 0x7FFFFFFFF)
    h = h * -4658895280553007687              # 0xBF58476D1CE4E5B9
    h = h ^ ((h >> 32) & 0xFFFFFFFF)
    u = (h & 0xFFFFFF).to(torch.float64) / float(1 << 24)
    return 2.0 * u - 1.0


def det_fill(sd, seed=0):
    """Overwrite every tensor of a state_dict with reproducible pseudo-random values (no torch RNG), sized like a
    freshly initialised network: conv weights uniform with std 1/sqrt(fan_in), BN scales around 1, running_var around
    1, spectral-norm u/v unit vectors."""
    with torch.no_grad():
        for i, (k, t) in enumerate(sd.items()):
            if k.endswith('num_batches_tracked'):
                t.zero_()
                continue
            n = t.numel()
            base = _hash_uniform(n, seed * 7919 + i)
            if k.endswith('weight_orig') or (k.endswith('weight') and t.dim() >= 4):
                fan_in = n // t.shape[0]
                val = base * (math.sqrt(3.0) / math.sqrt(fan_in))
            elif k.endswith('norm.weight'):
                val = 1.0 + 0.1 * base
            elif k.endswith('running_var'):
                val = 1.0 + 0.2 * base.abs()
            elif k.endswith('weight_u') or k.endswith('weight_v'):
                val = base / base.norm()
            else:
                val = 0.05 * base
            t.copy_(val.to(t.dtype).view_as(t))
    return sd


def det_tensor(shape, seed, scale=1.0, dtype=torch.float32):
    """reproducible pseudo-random tensor, uniform in [-scale, scale)"""
    n = 1
    for s in shape:
        n *= s
    return (_hash_uniform(n, 104729 + seed) * scale).to(dtype).view(*shape)


class Opt(object):
    """the fields of the reference's argparse namespace that forward reads (train_video.py:262-322, utils/images.py)"""

    def __init__(self, **kw):
        self.nc_im, self.nfc, self.latent_dim = 3, 64, 128
        self.ker_size, self.num_layer, self.padd_size, self.enc_blocks = 3, 5, 1, 2
        self.vae_levels, self.train_all = 3, False
        self.img_size, self.min_size, self.max_size, self.scale_factor_init = 64, 32, 256, 0.75
        self.ar = 1.0
        self.sampling_rates, self.org_fps = [4, 3, 2, 1], 24.0
        self.__dict__.update(kw)
        self.fps_lcm = 1
        for r in self.sampling_rates:
            self.fps_lcm = self.fps_lcm * r // math.gcd(self.fps_lcm, r)
        # utils.adjust_scales2image (utils/images.py:29-36)
        size = self.img_size
        self.num_scales = math.ceil(math.log(math.pow(self.min_size / size, 1), self.scale_factor_init)) + 1
        scale2stop = math.ceil(math.log(min([self.max_size, size]) / size, self.scale_factor_init))
        self.stop_scale = self.num_scales - scale2stop
        self.scale_factor = math.pow(self.min_size / size, 1 / self.stop_scale)
        self.stop_scale_time = self.stop_scale        # train_video.py sets opt.stop_scale_time = opt.stop_scale
        self.__dict__.update({k: v for k, v in kw.items() if k in ('stop_scale', 'scale_factor', 'stop_scale_time')})
