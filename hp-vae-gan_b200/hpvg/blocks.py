"""Dimension-agnostic building blocks behind the drop-in `modules.networks_3d` / `modules.networks_2d`.

Every class keeps the reference's parameter containers (nn.ConvNd, nn.BatchNormNd, legacy spectral_norm buffers),
so construction consumes the RNG exactly like the reference, `state_dict()` has identical keys and shapes
(SURVEY.md App. E), and `copy.deepcopy`, `.to()`, `nn.DataParallel.replicate` and optimizers work unchanged.  Only
`forward` differs: it hands the parameters to the libhpvg kernels through hpvg.ops.  Internally activations are
"wide" ([N,D,H,W,C] bf16) between layers and "thin" ([N,C,D,H,W] fp32) at the module boundary.

Reference classes restated here: ConvBlock3D/2D (modules/networks_3d.py:48-56, networks_2d.py:53-61),
ConvBlock3DSN/2DSN (:59-70 / :64-75), FeatureExtractor (:73-85 / :78-90), Encode3DVAE/Encode2DVAE (:88-107 / :93-112),
WDiscriminator3D/2D (:163-181 / :168-185), GeneratorHPVAEGAN (:325-406 / :188-269), GeneratorSG (:272-322),
GeneratorCSG (:213-269), WDiscriminatorBaselines (:184-210).
"""
import copy

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import images, ops

LRELU_SLOPE = 0.2


def _conv_cls(dims):
    return nn.Conv3d if dims == 3 else nn.Conv2d


def _bn_cls(dims):
    return nn.BatchNorm3d if dims == 3 else nn.BatchNorm2d


def as5d(t):
    """[N,C,H,W] -> [N,C,1,H,W] (2-D networks run the same kernels with D == 1)"""
    return t.unsqueeze(2) if t.dim() == 4 else t


def like_input(t5, dims):
    return t5.squeeze(2) if dims == 2 else t5


def _check_device(t):
    if not t.is_cuda:
        raise ops.lib.HpvgError("hpvg-b200 modules run on CUDA tensors only (got %s); there is no CPU fallback" % t.device)


class ConvBlock(nn.Sequential):
    """conv + BatchNorm(train statistics) + LeakyReLU(0.2); with bn=False, act=None it is a bare conv (mu/logvar heads)."""
    dims = 3

    def __init__(self, in_channel, out_channel, ker_size, padding, stride, bn=True, act='lrelu'):
        super().__init__()
        if ker_size != 3 or stride != 1:
            raise NotImplementedError("hpvg-b200 kernels cover 3x3(x3), stride 1 convolutions (the reference's only configuration)")
        if act not in (None, 'lrelu'):
            raise NotImplementedError("only LeakyReLU(0.2) is fused; got act=%r" % (act,))
        self.add_module('conv', _conv_cls(self.dims)(in_channel, out_channel, kernel_size=ker_size, stride=stride, padding=padding))
        if bn:
            self.add_module('norm', _bn_cls(self.dims)(out_channel))
        if act is not None:
            self.add_module(act, nn.LeakyReLU(LRELU_SLOPE, inplace=True))
        self.pad = int(padding)
        self.has_bn, self.has_act = bool(bn), act is not None

    def run(self, x, out_wide=True):
        """x: wide or thin 5-D tensor -> wide (or thin when out_wide=False and no BN/activation)"""
        conv = self.conv
        if self.has_bn:
            norm = self.norm
            if not self.has_act:
                raise NotImplementedError("BatchNorm without LeakyReLU is not on the reference path")
            if norm.training or not norm.track_running_stats:
                return ops.conv_bn_lrelu(x, conv.weight, conv.bias, norm.weight, norm.bias, norm.running_mean, norm.running_var,
                                         norm.num_batches_tracked, self.pad, momentum=norm.momentum, eps=norm.eps, slope=LRELU_SLOPE)
            raise NotImplementedError("eval-mode BatchNorm is never reached by the reference (the generator stays in train mode)")
        return ops.conv(x, conv.weight, conv.bias, self.pad, out_wide, LRELU_SLOPE if self.has_act else None)

    def forward(self, x):
        _check_device(x)
        y = self.run(as5d(x).contiguous(), out_wide=True)
        return like_input(ops.ToThin.apply(y), self.dims)


class ConvBlockSN(nn.Sequential):
    """spectral-normalised conv + LeakyReLU(0.2) (no BatchNorm: in the reference the `bn` flag selects spectral norm)."""
    dims = 3

    def __init__(self, in_channel, out_channel, ker_size, padding, stride, bn=True, act='lrelu'):
        super().__init__()
        if ker_size != 3 or stride != 1:
            raise NotImplementedError("hpvg-b200 kernels cover 3x3(x3), stride 1 convolutions")
        if not bn:
            raise NotImplementedError("the reflect-padded non-spectral branch is never used by the reference networks")
        if act not in (None, 'lrelu'):
            raise NotImplementedError("only LeakyReLU(0.2) is fused; got act=%r" % (act,))
        self.add_module('conv', nn.utils.spectral_norm(
            _conv_cls(self.dims)(in_channel, out_channel, kernel_size=ker_size, stride=stride, padding=padding)))
        if act is not None:
            self.add_module(act, nn.LeakyReLU(LRELU_SLOPE, inplace=True))
        self.pad = int(padding)
        self.has_act = act is not None

    def weight(self):
        """one power iteration (training mode) + W/sigma, as torch's legacy spectral_norm pre-forward hook does"""
        conv = self.conv
        return ops.SpectralWeight.apply(conv.weight_orig, conv.weight_u, conv.weight_v, conv.training, 1e-12)

    def run(self, x, out_wide=True, weight=None, in_link=None, out_link=None):
        """`weight`: W / sigma computed by the caller for all layers of the network at once (ops.spectral_weights);
        in_link / out_link: ops.ChainLink handshakes with the neighbouring blocks of a chain (_run_sn_chain)"""
        w = self.weight() if weight is None else weight
        w, token = ops.deferred_weight(w)        # (w, None) unless HPVG_CRITIC_WSIDE=1: weight gradient on the side stream
        return ops.conv(x, w, self.conv.bias, self.pad, out_wide, LRELU_SLOPE if self.has_act else None, in_link=in_link,
                        out_link=out_link if self.has_act else None, token=token)

    def forward(self, x):
        _check_device(x)
        return like_input(ops.ToThin.apply(self.run(as5d(x))), self.dims)


def _run_chain(seq, x):
    for m in seq:
        x = m.run(x)
    return x


def _sn_weights(blocks):
    """one batched power iteration for all blocks of a spectral-norm chain (the layers are independent, as in the reference where
    every forward pre-hook runs on its own weight) + the bf16 operand images of W / sigma (forward + data-gradient form)"""
    weights = ops.spectral_weights([b.conv for b in blocks])
    ops.prepack_pairs(weights)       # one launch for the bf16 operand images of all layers
    return weights


class sn_prefetch:
    """Spectral-norm prologues ahead of time.  A critic pass starts with ~50 us of small dependent launches that do not depend on
    its input: the power iteration of every layer (u, v updated in place: ONE iteration per forward call, exactly as the
    reference's pre-forward hook), W / sigma, and the operand images.  `sn_prefetch(module_chain, passes, stream)` runs that
    prologue for the next `passes` forward calls of the chain back to back on `stream` (same order, same arithmetic, same
    buffers as the calls themselves would) and queues the results; each of those forward calls then waits for its event and
    consumes its entry.  hpvg.train.ScaleTrainer uses it for the three critic passes of the critic step, whose weights do not
    change until optimizerD.step()."""

    def __init__(self, owner, blocks, passes, stream):
        self.owner, self.blocks, self.passes, self.stream = owner, list(blocks), int(passes), stream

    def run(self):
        queue = []
        main = torch.cuda.current_stream()
        self.stream.wait_stream(main)
        with torch.cuda.stream(self.stream):
            for _ in range(self.passes):
                weights = _sn_weights(self.blocks)
                ev = torch.cuda.Event()
                ev.record(self.stream)
                queue.append((weights, ev))
        self.owner._sn_queue = queue
        return self


def _run_sn_chain(blocks, x, owner=None):
    """a chain of spectral-norm blocks: the batched prologue (_sn_weights; taken from the owner's prefetch queue when one is
    pending), then the convolutions"""
    blocks = list(blocks)
    queue = getattr(owner, '_sn_queue', None) if owner is not None else None
    if queue:
        weights, ev = queue.pop(0)
        cur = torch.cuda.current_stream()
        cur.wait_event(ev)
        for w in weights:
            w.record_stream(cur)
    else:
        weights = _sn_weights(blocks)
    # each block's output feeds exactly the next block, so in the first-order backward the next block's data-gradient launch
    # can apply this block's LeakyReLU derivative (and sum its bias gradient) in its epilogue: see ops.ChainLink
    link = None
    for i, (b, w) in enumerate(zip(blocks, weights)):
        nxt = None
        if b.has_act and i + 1 < len(blocks):
            bias = b.conv.bias
            nxt = ops.ChainLink(LRELU_SLOPE, bias is not None and bias.requires_grad)
        x = b.run(x, weight=w, in_link=link, out_link=nxt)
        link = nxt
    return x


def make_family(dims):
    """Build the 2-D or 3-D family of classes (the reference keeps two copies of this code)."""

    class _ConvBlock(ConvBlock):
        pass

    class _ConvBlockSN(ConvBlockSN):
        pass

    _ConvBlock.dims = dims
    _ConvBlockSN.dims = dims
    Conv = _conv_cls(dims)

    class FeatureExtractor(nn.Sequential):
        def __init__(self, in_channel, out_channel, ker_size, padding, stride, num_blocks=2, return_linear=False):
            super().__init__()
            if return_linear:
                raise NotImplementedError("return_linear=True is never used by the reference networks")
            for i in range(num_blocks + 1):
                cin = in_channel if i == 0 else out_channel
                self.add_module('conv_block_{}'.format(i), _ConvBlockSN(cin, out_channel, ker_size, padding, stride))

        def run(self, x):
            return _run_sn_chain(self, x)

        def forward(self, x):
            _check_device(x)
            return like_input(ops.ToThin.apply(self.run(as5d(x))), dims)

    class EncodeVAE(nn.Module):
        def __init__(self, opt, out_dim=None, num_blocks=2):
            super().__init__()
            if out_dim is None:
                out_dim = opt.nfc
            elif type(out_dim) is not int:
                raise AssertionError("out_dim must be an int")
            half = opt.ker_size // 2
            self.features = FeatureExtractor(opt.nc_im, opt.nfc, opt.ker_size, half, 1, num_blocks=num_blocks)
            self.mu = _ConvBlock(opt.nfc, out_dim, opt.ker_size, half, 1, bn=False, act=None)
            self.logvar = _ConvBlock(opt.nfc, out_dim, opt.ker_size, half, 1, bn=False, act=None)

        def run(self, x5):
            """thin video -> (mu, logvar) wide"""
            feat = self.features.run(x5)
            return self.mu.run(feat), self.logvar.run(feat)

        def forward(self, x):
            _check_device(x)
            mu, logvar = self.run(as5d(x))
            return like_input(ops.ToThin.apply(mu), dims), like_input(ops.ToThin.apply(logvar), dims)

    class WDiscriminator(nn.Module):
        def __init__(self, opt):
            super().__init__()
            self.opt = opt
            nfc = int(opt.nfc)
            half = opt.ker_size // 2
            self.head = _ConvBlockSN(opt.nc_im, nfc, opt.ker_size, half, stride=1, bn=True, act='lrelu')
            self.body = nn.Sequential()
            for i in range(opt.num_layer):
                self.body.add_module('block%d' % i, _ConvBlockSN(nfc, nfc, opt.ker_size, half, stride=1, bn=True, act='lrelu'))
            self.tail = Conv(nfc, 1, kernel_size=opt.ker_size, padding=1, stride=1)

        def prefetch_spectral_weights(self, passes, stream):
            """run the spectral-norm prologue of the next `passes` forward calls ahead of time on `stream` (blocks.sn_prefetch)"""
            return sn_prefetch(self, [self.head] + list(self.body), passes, stream).run()

        def forward(self, x):
            _check_device(x)
            h = _run_sn_chain([self.head] + list(self.body), as5d(x).contiguous(), owner=self)
            out = ops.conv(h, self.tail.weight, self.tail.bias, 1, False)   # thin critic map [N,1,(T,)H,W]
            return like_input(out, dims)

    def _stage(opt, nfc, in_channels, padding):
        """head ConvBlock + num_layer ConvBlocks + bare tail conv: the decoder and every refinement stage"""
        seq = nn.Sequential()
        seq.add_module('head', _ConvBlock(in_channels, nfc, opt.ker_size, padding, stride=1))
        for i in range(opt.num_layer):
            seq.add_module('block%d' % i, _ConvBlock(nfc, nfc, opt.ker_size, padding, stride=1))
        return seq

    def _run_stage(seq, x5, tail_pad):
        """thin (or wide) input -> thin output of the stage's tail conv"""
        h = x5
        for name, m in seq.named_children():
            if name == 'tail':
                return ops.conv(h, m.weight, m.bias, tail_pad, False)
            h = m.run(h)
        raise RuntimeError("stage has no tail conv")

    class GeneratorHPVAEGAN(nn.Module):
        def __init__(self, opt):
            super().__init__()
            self.opt = opt
            self.N = int(opt.nfc)
            self.encode = self._make_encoder(opt)      # first, as in the reference: construction order fixes the initial values
            self.decoder = _stage(opt, self.N, opt.latent_dim, opt.padd_size)
            self.decoder.add_module('tail', Conv(self.N, opt.nc_im, opt.ker_size, 1, opt.ker_size // 2))
            self.body = torch.nn.ModuleList([])

        def _make_encoder(self, opt):
            return EncodeVAE(opt, out_dim=opt.latent_dim, num_blocks=opt.enc_blocks)

        def init_next_stage(self):
            if len(self.body) == 0:
                first = _stage(self.opt, self.N, self.opt.nc_im, self.opt.padd_size)
                first.add_module('tail', Conv(self.N, self.opt.nc_im, self.opt.ker_size, 1, self.opt.ker_size // 2))
                # built on the CPU like the reference's (same initial values for a given seed), then moved to the generator's
                # device: the train scripts call init_next_stage() AFTER netG.to(device) (train_video.py:397,416) and wrap
                # the result in nn.DataParallel, which requires every parameter on that device
                self.body.append(first.to(next(self.decoder.parameters()).device))
            else:
                self.body.append(copy.deepcopy(self.body[-1]))

        def forward(self, video, noise_amp, noise_init=None, sample_init=None, mode='rand'):
            if sample_init is not None:
                assert len(self.body) > sample_init[0], "Strating index must be lower than # of body blocks"
            opt = self.opt
            half = opt.ker_size // 2
            if noise_init is None:
                _check_device(video)
                mu_w, logvar_w = self.encode.run(as5d(video).contiguous())
                mu5, logvar5 = ops.ToThin.apply(mu_w), ops.ToThin.apply(logvar_w)
                mu, logvar = like_input(mu5, dims), like_input(logvar5, dims)
                if self.training:
                    eps = as5d(images.generate_noise(ref=mu))      # same draw as reparameterize (reference :31-32)
                    z = ops.Reparam.apply(mu_w, logvar_w, eps)
                else:
                    z = ops.ToWide.apply(as5d(images.generate_noise(ref=mu)))
            else:
                _check_device(noise_init)
                z = ops.ToWide.apply(as5d(noise_init).contiguous())
            # tanh runs on the caller-visible shape so that the result is not a view (refinement may detach_() it in place)
            vae_out = ops.TanhAdd.apply(like_input(_run_stage(self.decoder, z, half), dims), None)

            if sample_init is not None:
                x_prev_out = self.refinement_layers(sample_init[0], sample_init[1], noise_amp, mode)
            else:
                x_prev_out = self.refinement_layers(0, vae_out, noise_amp, mode)

            if noise_init is None:
                return x_prev_out, vae_out, (mu, logvar)
            return x_prev_out, vae_out

        def refinement_layers(self, start_idx, x_prev_out, noise_amp, mode):
            opt = self.opt
            half = opt.ker_size // 2
            for idx, block in enumerate(self.body[start_idx:], start_idx):
                if opt.vae_levels == idx + 1 and not opt.train_all:
                    x_prev_out.detach_()
                x5 = as5d(x_prev_out)
                size = images.video_target_size(idx + 1, opt) if dims == 3 else [1] + images.image_target_size(idx + 1, opt)
                x_up = images.resize(x5, size)
                # 3-D adds noise only from the first GAN level on; 2-D adds it at every level (networks_2d.py:261-263)
                if mode == 'rand' and (dims == 2 or opt.vae_levels <= idx + 1):
                    noise = as5d(images.generate_noise(ref=like_input(x_up, dims)))
                    x_in = images.resize(x5, size, noise=noise, amp=float(noise_amp[idx + 1]))
                else:
                    x_in = x_up
                x_prev_out = ops.TanhAdd.apply(like_input(_run_stage(block, x_in, half), dims), like_input(x_up, dims))
            return x_prev_out


    Pool = nn.AdaptiveAvgPool3d if dims == 3 else nn.AdaptiveAvgPool2d

    class EncodeVAE_nb(nn.Module):
        """Bernoulli-gated encoder (reference modules/networks_3d.py:110-138, networks_2d.py:115-143): a 1-channel sigmoid gate
        multiplies the features, mu / logvar are global averages of their conv maps.  The convolutions (spectral-norm feature
        chain, the nfc -> 1 gate, the two nfc -> out_dim heads) run on the library's kernels; the gate product, the sigmoid and
        the global average are small torch element-wise ops on the way."""

        def __init__(self, opt, out_dim=None, num_blocks=2):
            super().__init__()
            if out_dim is None:
                out_dim = opt.nfc
            elif type(out_dim) is not int:
                raise AssertionError("out_dim must be an int")
            half = opt.ker_size // 2
            self.features = FeatureExtractor(opt.nc_im, opt.nfc, opt.ker_size, half, 1, num_blocks=num_blocks)
            self.mu = nn.Sequential(_ConvBlock(opt.nfc, out_dim, opt.ker_size, half, 1, bn=False, act=None), Pool(1))
            self.logvar = nn.Sequential(_ConvBlock(opt.nfc, out_dim, opt.ker_size, half, 1, bn=False, act=None), Pool(1))
            self.bern = _ConvBlock(opt.nfc, 1, opt.ker_size, half, 1, bn=False, act=None)

        def forward(self, x):
            _check_device(x)
            feat = self.features.run(as5d(x).contiguous())                            # wide [N,D,H,W,nfc]
            gate = torch.sigmoid(self.bern.run(feat, out_wide=False))                 # thin [N,1,D,H,W]
            gated = (feat * gate.permute(0, 2, 3, 4, 1).to(feat.dtype)).contiguous()

            def pooled(head):
                m = head[0].run(gated)                                                # wide [N,D,H,W,out_dim]
                v = m.float().mean(dim=(1, 2, 3))                                     # AdaptiveAvgPool(1)
                return v.view(v.shape[0], v.shape[1], *([1] * dims))
            return pooled(self.mu), pooled(self.logvar), like_input(gate, dims)

    def reparameterize_bern(x, training):
        """reference modules/networks_3d.py:38-43 (Gumbel-style relaxation of the gate)"""
        if training:
            eps = torch.zeros_like(x).uniform_()
            return torch.log(x + 1e-20) - torch.log(-torch.log(eps + 1e-20) + 1e-20)
        return torch.zeros_like(x).bernoulli_()

    class GeneratorVAE_nb(GeneratorHPVAEGAN):
        """HP-VAE-GAN generator with the Bernoulli-gated encoder (reference modules/networks_3d.py:409-485, networks_2d.py:272-348):
        the latent is z_norm (a [N,latent,1,..] vector) times z_bern (a [N,1,T,H,W] map); every refinement level adds noise in
        'rand' mode and the first GAN level detaches unconditionally.  Decoder and refinement stages are the same kernels."""

        def _make_encoder(self, opt):
            return EncodeVAE_nb(opt, out_dim=opt.latent_dim, num_blocks=opt.enc_blocks)

        def forward(self, video, noise_amp, noise_init_norm=None, noise_init_bern=None, sample_init=None, mode='rand'):
            if sample_init is not None:
                assert len(self.body) > sample_init[0], "Strating index must be lower than # of body blocks"
            half = self.opt.ker_size // 2
            if noise_init_norm is None:
                mu, logvar, bern = self.encode(video)
                if self.training:
                    eps = torch.zeros_like(logvar).normal_()
                    z_norm = eps.mul(logvar.mul(0.5).exp()).add(mu)
                else:
                    z_norm = torch.zeros_like(mu).normal_()
                z_bern = reparameterize_bern(bern, self.training)
            else:
                _check_device(noise_init_norm)
                z_norm, z_bern = noise_init_norm, noise_init_bern
            z = ops.ToWide.apply(as5d(z_norm * z_bern).contiguous())
            vae_out = ops.TanhAdd.apply(like_input(_run_stage(self.decoder, z, half), dims), None)
            if sample_init is not None:
                x_prev_out = self.refinement_layers(sample_init[0], sample_init[1], noise_amp, mode)
            else:
                x_prev_out = self.refinement_layers(0, vae_out, noise_amp, mode)
            if noise_init_norm is None:
                return x_prev_out, vae_out, (mu, logvar, bern)
            return x_prev_out, vae_out

        def refinement_layers(self, start_idx, x_prev_out, noise_amp, mode):
            opt = self.opt
            half = opt.ker_size // 2
            for idx, block in enumerate(self.body[start_idx:], start_idx):
                if opt.vae_levels == idx + 1:
                    x_prev_out.detach_()
                x5 = as5d(x_prev_out)
                size = images.video_target_size(idx + 1, opt) if dims == 3 else [1] + images.image_target_size(idx + 1, opt)
                x_up = images.resize(x5, size)
                if mode == 'rand':
                    noise = as5d(images.generate_noise(ref=like_input(x_up, dims)))
                    x_in = images.resize(x5, size, noise=noise, amp=float(noise_amp[idx + 1]))
                else:
                    x_in = x_up
                x_prev_out = ops.TanhAdd.apply(like_input(_run_stage(block, x_in, half), dims), like_input(x_up, dims))
            return x_prev_out

    class _TorchSNBlock(nn.Sequential):
        """spectral-norm conv + LeakyReLU in plain torch: pass-through for kernel sizes the library does not cover (1x1x1)"""

        def __init__(self, cin, cout, ker_size, padding, stride):
            super().__init__()
            self.add_module('conv', nn.utils.spectral_norm(Conv(cin, cout, kernel_size=ker_size, stride=stride, padding=padding)))
            self.add_module('lrelu', nn.LeakyReLU(LRELU_SLOPE, inplace=True))

    class _TorchConv(nn.Sequential):
        def __init__(self, cin, cout, ker_size, padding, stride):
            super().__init__()
            self.add_module('conv', Conv(cin, cout, kernel_size=ker_size, stride=stride, padding=padding))

    class EncodeVAE1x1(nn.Module):
        """1x1x1 encoder (reference modules/networks_3d.py:141-160; no network of the reference instantiates it): plain torch
        pass-through — a 1x1x1 convolution is a per-voxel matrix product, outside the 3x3x3 kernels of this library."""

        def __init__(self, opt, out_dim=None):
            super().__init__()
            if out_dim is None:
                out_dim = opt.nfc
            elif type(out_dim) is not int:
                raise AssertionError("out_dim must be an int")
            self.features = nn.Sequential()
            for i in range(3):
                self.features.add_module('conv_block_{}'.format(i), _TorchSNBlock(opt.nc_im if i == 0 else opt.nfc, opt.nfc, 1, 0, 1))
            self.mu = _TorchConv(opt.nfc, out_dim, 1, 0, 1)
            self.logvar = _TorchConv(opt.nfc, out_dim, 1, 0, 1)

        def forward(self, x):
            feat = self.features(x)
            return self.mu(feat), self.logvar(feat)

    class GeneratorSG(nn.Module):
        """SinGAN-style baseline: valid (pad 0) convolutions on inputs zero-padded by num_layer + 2 voxels (3-D only)."""

        def __init__(self, opt):
            super().__init__()
            if dims != 3:
                raise NotImplementedError("GeneratorSG exists only in networks_3d")
            self.opt = opt
            nfc = int(opt.nfc)
            self.margin = opt.num_layer + 2
            self.p3d = (self.margin,) * 6
            self.body = nn.ModuleList([])
            first = _stage(opt, nfc, opt.nc_im, 0)
            first.add_module('tail', Conv(nfc, opt.nc_im, kernel_size=opt.ker_size, padding=0, stride=1))
            self.body.append(first)
            self.apply(weights_init)

        def init_next_stage(self):
            self.body.append(copy.deepcopy(self.body[-1]))

        def forward(self, noise_init, noise_amp, mode='rand'):
            _check_device(noise_init)
            x_prev_out = _run_stage(self.body[0], F.pad(noise_init, self.p3d).contiguous(), 0)
            for idx, block in enumerate(self.body[1:], 1):
                x_prev_out = ops.TanhAdd.apply(x_prev_out, None)
                size = images.video_target_size(idx, self.opt)
                x_up = images.resize(x_prev_out, size)
                if mode == 'rand':
                    big = [s + 2 * self.margin for s in size]
                    noise = images.generate_noise(size=[x_prev_out.shape[0], x_prev_out.shape[1]] + big, device=x_prev_out.device)
                    x_in = images.resize(x_prev_out, big, noise=noise, amp=float(noise_amp[idx]))
                else:
                    x_in = F.pad(x_up, self.p3d).contiguous()
                x_prev_out = _run_stage(block, x_in, 0) + x_up
            return ops.TanhAdd.apply(x_prev_out, None)

    class GeneratorCSG(nn.Module):
        """Feature-space SinGAN baseline (reference modules/networks_3d.py:213-269, the default generator of
        train_video_baselines.py): a BatchNorm head lifts the noise to nfc channels, every stage is num_layer valid (pad 0)
        ConvBlocks on an input zero-padded by num_layer voxels, stages are chained in FEATURE space — trilinear resize of the
        nfc-channel map, + noise_amp * noise, residual add — and one conv + tanh tail maps to the image.  3-D only."""

        def __init__(self, opt):
            super().__init__()
            if dims != 3:
                raise NotImplementedError("GeneratorCSG exists only in networks_3d")
            self.opt = opt
            nfc = int(opt.nfc)
            self.margin = int(opt.num_layer)
            self.p3d_once = (1,) * 6
            self.p3d = (self.margin,) * 6
            self.head = _ConvBlock(opt.nc_im, nfc, opt.ker_size, padding=0, stride=1)
            self.body = nn.ModuleList([])
            first = nn.Sequential()
            for i in range(opt.num_layer):
                first.add_module('block%d' % i, _ConvBlock(nfc, nfc, opt.ker_size, padding=0, stride=1))
            self.body.append(first)
            self.tail = nn.Sequential(Conv(nfc, opt.nc_im, kernel_size=opt.ker_size, padding=0, stride=1), nn.Tanh())
            self.apply(weights_init)

        def init_next_stage(self):
            self.body.append(copy.deepcopy(self.body[-1]))

        def forward(self, noise_init, noise_amp, mode='rand'):
            _check_device(noise_init)
            head, norm = self.head.conv, self.head.norm
            # head(F.pad(noise, 1)) with a valid convolution == the same convolution with padding 1 (zero fill in the TMA box /
            # halo loads); BatchNorm statistics run over the output, which is identical
            x = ops.conv_bn_lrelu(noise_init.contiguous(), head.weight, head.bias, norm.weight, norm.bias, norm.running_mean,
                                  norm.running_var, norm.num_batches_tracked, 1, momentum=norm.momentum, eps=norm.eps, slope=LRELU_SLOPE)
            x_prev_out = _run_chain(self.body[0], ops.PadWide.apply(x, self.margin))
            for idx, block in enumerate(self.body[1:], 1):
                size = images.video_target_size(idx, self.opt)
                x_up = ops.UpsampleWide.apply(x_prev_out, tuple(size), None, 0.0)
                if mode == 'rand':
                    big = tuple(s + 2 * self.margin for s in size)
                    n, c = x_prev_out.shape[0], x_prev_out.shape[-1]
                    noise = images.generate_noise(size=[n, c] + list(big), device=x_prev_out.device)
                    x_in = ops.UpsampleWide.apply(x_prev_out, big, noise, float(noise_amp[idx]))
                else:
                    x_in = ops.PadWide.apply(x_up, self.margin)
                x_prev_out = ops.AddWide.apply(_run_chain(block, x_in), x_up)
            tail = self.tail[0]
            out = ops.conv(x_prev_out, tail.weight, tail.bias, 1, False)      # tail(F.pad(x, 1)), valid conv == padding 1
            return ops.TanhAdd.apply(out, None)

    class WDiscriminatorBaselines(nn.Module):
        """BatchNorm critic of the SinGAN-style baselines (reference modules/networks_3d.py:184-210): the input is zero-padded
        by num_layer + 2 voxels, head = conv + LeakyReLU, num_layer ConvBlocks, plain tail conv.  Forward and first-order
        backward run on the library's kernels (fused conv + BatchNorm node); under calc_gradient_penalty the blocks switch to
        the double-differentiable form (hpvg.ops.twice_differentiable: library convolutions, BatchNorm + LeakyReLU composed of
        element-wise torch operations), so the WGAN-GP double backward of train_video_baselines.py --discriminator
        WDiscriminatorBaselines works.  3-D only."""

        def __init__(self, opt):
            super().__init__()
            if dims != 3:
                raise NotImplementedError("WDiscriminatorBaselines exists only in networks_3d")
            self.opt = opt
            nfc = int(opt.nfc)
            self.p3d = (opt.num_layer + 2,) * 6
            self.head = _ConvBlock(opt.nc_im, nfc, opt.ker_size, opt.padd_size, stride=1, bn=False, act='lrelu')
            self.body = nn.Sequential()
            for i in range(opt.num_layer):
                self.body.add_module('block%d' % i, _ConvBlock(nfc, nfc, opt.ker_size, opt.padd_size, stride=1, bn=True, act='lrelu'))
            self.tail = Conv(nfc, 1, kernel_size=opt.ker_size, padding=opt.padd_size, stride=1)
            self.apply(weights_init)

        def forward(self, x):
            _check_device(x)
            h = self.head.run(F.pad(x, self.p3d).contiguous())
            h = _run_chain(self.body, h)
            return ops.conv(h, self.tail.weight, self.tail.bias, int(self.opt.padd_size), False)

    return {
        'ConvBlock': _ConvBlock, 'ConvBlockSN': _ConvBlockSN, 'FeatureExtractor': FeatureExtractor, 'EncodeVAE': EncodeVAE,
        'WDiscriminator': WDiscriminator, 'GeneratorHPVAEGAN': GeneratorHPVAEGAN, 'GeneratorSG': GeneratorSG,
        'GeneratorCSG': GeneratorCSG, 'WDiscriminatorBaselines': WDiscriminatorBaselines,
        'EncodeVAE_nb': EncodeVAE_nb, 'GeneratorVAE_nb': GeneratorVAE_nb, 'EncodeVAE1x1': EncodeVAE1x1,
        'reparameterize_bern': reparameterize_bern,
    }


def weights_init(m):
    """N(0, 0.02) conv weights, N(1, 0.02) norm scales — applied by the SinGAN-style baselines only (reference :9-15)."""
    name = m.__class__.__name__
    if 'Conv2d' in name or 'Conv3d' in name:
        m.weight.data.normal_(0.0, 0.02)
    elif 'Norm' in name:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)
