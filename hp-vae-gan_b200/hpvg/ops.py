"""torch.autograd.Function wrappers over libhpvg.so.

Tensor conventions inside the networks
  wide : torch.bfloat16, shape [N, D, H, W, C], contiguous  (NDHWC_BF16; every 64/128-channel activation)
  thin : torch.float32,  shape [N, C, D, H, W], contiguous  (NCDHW_F32; videos, critic maps, the latent at the API)
2-D networks use D == 1 and 3x3 weights (KD == 1).

The conv family {ConvFwd, ConvDgrad, ConvWgrad, LReluBwd} is closed under differentiation: each backward is
written with the other Functions (never `once_differentiable`), so `torch.autograd.grad(..., create_graph=True)`
in calc_gradient_penalty (reference modules/utils.py:14-16) builds the second-order graph out of the same
kernels.
"""
import os
import threading

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import lib
from .lib import ACT_LRELU, ACT_NONE, FMT_NCDHW_F32, FMT_NDHWC_BF16

_INPUT_ONLY = [False]


class input_grad_only:
    """Context: conv backward passes inside it skip weight/bias gradients (used around the first-order
    autograd.grad of the gradient penalty, which only asks for d/d(interpolates)).
    A process-wide flag, not a thread-local: for CUDA tensors the backward of a Function runs on the autograd engine's
    device thread, not on the thread that called torch.autograd.grad (which blocks until the sweep is done), so a
    thread-local set by the caller is invisible exactly where it is read."""

    def __enter__(self):
        self.prev = _INPUT_ONLY[0]
        _INPUT_ONLY[0] = True

    def __exit__(self, *a):
        _INPUT_ONLY[0] = self.prev


def _input_only():
    return _INPUT_ONLY[0]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise lib.HpvgError("hpvg ops run on CUDA tensors only (got a %s tensor); there is no CPU fallback" % t.device)


def is_wide(t):
    return t.dtype == torch.bfloat16


def fmt_of(t):
    return FMT_NDHWC_BF16 if t.dtype == torch.bfloat16 else FMT_NCDHW_F32


def dims_of(t):
    """-> (N, C, D, H, W) of a wide or thin tensor"""
    if t.dim() != 5:
        raise ValueError("expected a 5-D tensor, got shape %s" % (tuple(t.shape),))
    if is_wide(t):
        n, d, h, w, c = t.shape
    else:
        n, c, d, h, w = t.shape
    return n, c, d, h, w


def _empty(n, c, d, h, w, wide, device):
    if wide:
        return torch.empty((n, d, h, w, c), dtype=torch.bfloat16, device=device)
    return torch.empty((n, c, d, h, w), dtype=torch.float32, device=device)


class _Arena:
    """bump allocator over ONE zero-filled float32 buffer: the small accumulators that kernels add into with atomics
    (BatchNorm sums, fused bias-gradient sums) are carved from it, so an iteration pays one fill launch instead of ~80"""

    def __init__(self, device, capacity):
        self.buf = torch.zeros((int(capacity),), dtype=torch.float32, device=device)
        self.used = 0
        self.lock = threading.Lock()

    def take(self, n):
        n_al = (int(n) + 31) & ~31          # 128-byte granules: no two accumulators share a cache line
        with self.lock:
            if self.used + n_al > self.buf.numel():
                return None
            t = self.buf[self.used:self.used + n]
            self.used += n_al
        return t


_ARENA = [None]


class zero_arena:
    """`with zero_arena(device, capacity_floats):` — zeros_small() inside the block hands out slices of one pre-zeroed buffer.
    The buffer is created on the current stream when the block is entered: enter it BEFORE forking side streams."""

    def __init__(self, device, capacity=16384):
        self.device, self.capacity = device, capacity

    def __enter__(self):
        self.prev = _ARENA[0]
        _ARENA[0] = _Arena(self.device, self.capacity)
        return self

    def __exit__(self, *a):
        _ARENA[0] = self.prev


def zeros_small(n, device):
    """a zero-filled float32 vector of n elements for kernels that accumulate with atomics"""
    a = _ARENA[0]
    if a is not None and a.buf.device == torch.device(device):
        t = a.take(n)
        if t is not None:
            return t
    return torch.zeros((int(n),), dtype=torch.float32, device=device)


def _kd_of(weight):
    return 3 if weight.dim() == 5 else 1


# GEMM + shift-add kernel for the 64 -> <= 4 channel layers (thin_gs.cu).  OFF by default (HPVG_THIN_GS=1 turns it on): parity-green,
# but measured on B200 in a dependent chain of 10 launches at 16 x 64 x 64 it takes 14.6 us per launch against 12.3 us for the
# 16-rows-per-tap tcgen05 kernel — 8 instead of 54 MMAs per slab, but 110 KB of shared-memory traffic per slab for the partial products.
_THIN_GS = os.environ.get('HPVG_THIN_GS', '0') == '1'
THIN_ROWS = 16   # rows per tap of the packed weights of the thin-output tcgen05 kernel (Cout <= 16, zero padded)


def _tc_eligible(cin, cout, x_wide, y_wide, plain=True):
    """which convolutions run on the tcgen05 kernels: wide -> wide with Cin in {64,128} and Cout a multiple of 64, and
    wide -> thin with Cin == 64 and Cout <= 16 (bias only: `plain`)"""
    if not x_wide or lib.get_conv_backend() == lib.BACKEND_DIRECT:
        return False
    if y_wide:
        return cin in (64, 128) and cout % 64 == 0
    return plain and cin == 64 and cout <= THIN_ROWS


def pack_weights(w, cout, cin, taps, transposed, sigma=None, rows=None):
    rows = cout if rows is None else rows
    out = torch.empty((taps, rows, cin), dtype=torch.bfloat16, device=w.device)
    lib.call("hpvg_pack_weights", _ptr(w), _ptr(out), cout, cin, taps, int(transposed), _ptr(sigma), rows, _stream())
    return out


_PACK_GEN = [0]


def invalidate_packed_weights():
    """Packed bf16 weight images are cached on the weight tensor, stamped with its version counter.  A CUDA-graph replay
    updates weights without touching version counters, so whoever replays a graph that contains optimizer steps calls
    this (hpvg.train.ScaleTrainer.replay does)."""
    _PACK_GEN[0] += 1


def packed_for(w, cout, cin, taps, transposed, rows):
    """bf16 operand image of `w` for the tcgen05 kernels, repacked only when the weight changed: the same weight is used
    by the 'rec' and 'rand' generator passes and, transposed, by their data-gradient passes within one iteration"""
    key = (bool(transposed), rows)
    stamp = (w._version, _PACK_GEN[0], w.data_ptr())
    owner = w._base if (w._base is not None and w._base.shape == w.shape) else w    # WeightProxy hands out full views
    cache = getattr(owner, '_hpvg_packs', None)
    if cache is not None:
        hit = cache.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
    packed = pack_weights(w, cout, cin, taps, transposed, rows=rows)
    if cache is None:
        try:
            owner._hpvg_packs = cache = {}
        except (AttributeError, RuntimeError):
            return packed
    cache[key] = (stamp, packed)
    return packed


def prepack_pairs(weights):
    """Pack the forward and the data-gradient operand image of every tcgen05-eligible weight in `weights` (float32,
    [Cout, Cin, (3,)3,3], e.g. the W / sigma tensors of one critic pass) in ONE launch and plant them in the per-weight cache,
    so that the packed_for() calls of the convolutions that follow hit.  Same values as pack_weights()."""
    if lib.get_conv_backend() == lib.BACKEND_DIRECT:
        return
    todo = [w for w in weights if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.dim() in (4, 5)
            and w.shape[1] in (64, 128) and w.shape[0] % 64 == 0 and w._base is None]
    for i in range(0, len(todo), lib.SN_MAX_LAYERS):
        chunk = todo[i:i + lib.SN_MAX_LAYERS]
        fwd, tr, couts, cins, taps = [], [], [], [], []
        for w in chunk:
            cout, cin = w.shape[0], w.shape[1]
            t = _kd_of(w) * 9
            fwd.append(torch.empty((t, cout, cin), dtype=torch.bfloat16, device=w.device))
            tr.append(torch.empty((t, cin, cout), dtype=torch.bfloat16, device=w.device))
            couts.append(cout); cins.append(cin); taps.append(t)
        lib.call("hpvg_pack_weights_pair_batched", len(chunk), lib.ptr_array(chunk), lib.ptr_array(fwd), lib.ptr_array(tr),
                 lib.int_array(couts), lib.int_array(cins), lib.int_array(taps), _stream())
        for w, f, r, cout, cin in zip(chunk, fwd, tr, couts, cins):
            stamp = (w._version, _PACK_GEN[0], w.data_ptr())
            cache = getattr(w, '_hpvg_packs', None)
            if cache is None:
                try:
                    w._hpvg_packs = cache = {}
                except (AttributeError, RuntimeError):
                    continue
            cache[(False, cout)] = (stamp, f)        # packed_for(w, cout, cin, taps, False, rows=cout)
            cache[(True, cin)] = (stamp, r)          # packed_for(w, cin, cout, taps, True, rows=cin): the data-gradient form


def prepack_module(module, backward=True):
    """Bring the cached operand images of every convolution weight of `module` up to date on the CURRENT stream: the bf16
    [tap][Cout][Cin] image of the tcgen05 layers (and its transposed data-gradient form when `backward`), the 16-row image of the
    thin-output tails, the float32 filter image of the 3-channel heads.  conv_raw() packs lazily on whatever stream first needs an
    image and later users take the cached tensor without a stream dependency — fine on one stream, a race when two passes that
    share weights run on two streams.  hpvg.train.ScaleTrainer calls this before it forks the generator's passes."""
    if lib.get_conv_backend() == lib.BACKEND_DIRECT:
        return
    for w in module.parameters():
        if w.dim() not in (4, 5) or not w.is_cuda or w.dtype != torch.float32:
            continue
        cout, cin = w.shape[0], w.shape[1]
        taps = _kd_of(w) * 9
        wc = w            # the Parameter object itself: the cache lives on it (packed_for), the kernels only see its pointer
        if cin in (64, 128) and cout % 64 == 0:
            packed_for(wc, cout, cin, taps, False, cout)
            if backward and w.requires_grad:
                packed_for(wc, cin, cout, taps, True, cin)
        elif cin == 64 and cout <= THIN_ROWS:
            packed_for(wc, cout, cin, taps, False, cout if (_THIN_GS and cout <= 4) else THIN_ROWS)
            if backward and w.requires_grad:
                expand_image_for(wc, cout, taps, True)
        elif cin <= 4 and cout == 64:
            expand_image_for(wc, cin, taps, False)
            if backward and w.requires_grad:
                packed_for(wc, cin, cout, taps, True, cin if _THIN_GS else THIN_ROWS)


def expand_image_for(w, cin, taps, transposed):
    """float32 [taps][cin][64] filter image for the thin -> wide kernel, cached on the weight like packed_for()"""
    key = ('expand', bool(transposed))
    stamp = (w._version, _PACK_GEN[0], w.data_ptr())
    owner = w._base if (w._base is not None and w._base.shape == w.shape) else w
    cache = getattr(owner, '_hpvg_packs', None)
    if cache is not None:
        hit = cache.get(key)
        if hit is not None and hit[0] == stamp:
            return hit[1]
    out = torch.empty((taps, cin, 64), dtype=torch.float32, device=w.device)
    lib.call("hpvg_pack_weights_expand", _ptr(w), _ptr(out), cin, taps, int(transposed), _stream())
    if cache is None:
        try:
            owner._hpvg_packs = cache = {}
        except (AttributeError, RuntimeError):
            return out
    cache[key] = (stamp, out)
    return out


def conv_raw(x, w, bias, pad, transposed, out_wide, act_slope=None, stats=None, mask_src=None, mask_slope=None,
             stats_per_sample=False):
    """One hpvg_conv_forward call.  `w` is the float32 weight of the *forward* convolution ([Cout_f, Cin_f, (3,)3,3]);
    transposed=True computes the data gradient form with it (input channels = Cout_f, output channels = Cin_f).
    mask_src / mask_slope: the epilogue multiplies the result by the LeakyReLU derivative read from `mask_src` (a wide
    tensor of the output's extents): "dgrad then leaky_relu_backward" in one launch.
    stats_per_sample: `stats` is [N, 2*Cout] and every sample accumulates into its own row (tcgen05 / expand kernels only)."""
    _require_cuda(x, w)
    x = x.contiguous()
    w = w.contiguous()
    if w.dtype != torch.float32:
        raise TypeError("conv weights must be float32 masters")
    n, cx, d, h, wd = dims_of(x)
    kd = _kd_of(w)
    taps = kd * 9
    cout_f, cin_f = w.shape[0], w.shape[1]
    cin, cout = (cout_f, cin_f) if transposed else (cin_f, cout_f)
    if cx != cin:
        raise ValueError("conv: input has %d channels, weight expects %d" % (cx, cin))
    if kd == 1 and d != 1:
        raise ValueError("2-D convolution needs D == 1")
    pad_d = pad if kd == 3 else 0
    do, ho, wo = d + 2 * pad_d - (kd - 1), h + 2 * pad - 2, wd + 2 * pad - 2
    y = _empty(n, cout, do, ho, wo, out_wide, x.device)
    packed = None
    plain = act_slope is None and stats is None and mask_src is None
    thin_gs = (_THIN_GS and plain and is_wide(x) and not out_wide and cin == 64 and cout <= 4
               and lib.get_conv_backend() != lib.BACKEND_DIRECT)
    if thin_gs:
        packed = packed_for(w, cout, cin, taps, transposed, cout)      # [taps][Cout][64]: thin_gs.cu's resident operand tile
    elif _tc_eligible(cin, cout, is_wide(x), out_wide, plain=plain):
        packed = packed_for(w, cout, cin, taps, transposed, cout if out_wide else THIN_ROWS)
    elif (not is_wide(x)) and out_wide and cin <= 4 and cout == 64 and mask_src is None and lib.get_conv_backend() != lib.BACKEND_DIRECT:
        packed = expand_image_for(w, cin, taps, transposed)
    if bias is not None:
        bias = bias.contiguous()
    if mask_src is not None:
        if act_slope is not None or mask_slope is None or not out_wide or tuple(mask_src.shape) != tuple(y.shape):
            raise ValueError("conv: mask_src needs a wide output of the same extents, a slope and no activation")
        mask_src = mask_src.contiguous()
    slope = act_slope if act_slope is not None else (mask_slope if mask_src is not None else 0.0)
    if stats_per_sample:
        if stats is None or tuple(stats.shape) != (n, 2 * cout):
            raise ValueError("per-sample statistics need a [N, 2*Cout] float32 accumulator")
        if packed is None:
            raise lib.HpvgError("per-sample statistics need the tcgen05 or the expand kernel (Cin=%d Cout=%d)" % (cin, cout))
    lib.call("hpvg_conv_forward_ex", _ptr(x), fmt_of(x), _ptr(w), _ptr(packed), _ptr(bias), _ptr(y), fmt_of(y), n, cin, cout, d, h, wd,
             kd, pad, int(transposed) | (2 if thin_gs else 0), ACT_LRELU if act_slope is not None else ACT_NONE, float(slope), _ptr(stats),
             int(bool(stats_per_sample)), _ptr(mask_src), _stream())
    return y


def per_sample_stats_supported(x, w):
    """whether conv_raw(x, w, ..., out_wide=True, stats_per_sample=True) has a kernel: the tcgen05 and the thin -> wide layers"""
    cout, cin = w.shape[0], w.shape[1]
    if lib.get_conv_backend() == lib.BACKEND_DIRECT:
        return False
    if is_wide(x):
        return cin in (64, 128) and cout % 64 == 0
    return cin <= 4 and cout == 64


def wgrad_raw(x, gy, pad, wshape, want_bias=False):
    """dw (float32, `wshape` = forward weight shape) [and dbias] from input x and output gradient gy."""
    _require_cuda(x, gy)
    x = x.contiguous()
    gy = gy.contiguous()
    n, cin, d, h, wd = dims_of(x)
    n2, cout, do, ho, wo = dims_of(gy)
    kd = 3 if len(wshape) == 5 else 1
    pad_d = pad if kd == 3 else 0
    if (n2, do, ho, wo) != (n, d + 2 * pad_d - (kd - 1), h + 2 * pad - 2, wd + 2 * pad - 2) or cout != wshape[0] or cin != wshape[1]:
        raise ValueError("wgrad: inconsistent shapes x=%s gy=%s w=%s pad=%d" % (tuple(x.shape), tuple(gy.shape), tuple(wshape), pad))
    dw = torch.empty(tuple(wshape), dtype=torch.float32, device=x.device)
    db = torch.empty((cout,), dtype=torch.float32, device=x.device) if want_bias else None
    nbytes = lib.load().hpvg_conv_wgrad_workspace(n, cin, cout, d, h, wd, kd, pad, fmt_of(x), fmt_of(gy))
    ws = torch.empty((max(int(nbytes), 16),), dtype=torch.uint8, device=x.device)
    lib.call("hpvg_conv_wgrad", _ptr(x), fmt_of(x), _ptr(gy), fmt_of(gy), _ptr(dw), _ptr(db), n, cin, cout, d, h, wd, kd, pad,
             _ptr(ws), int(nbytes), _stream())
    return dw, db


def channel_sum(t):
    n, c, d, h, w = dims_of(t)
    out = torch.empty((c,), dtype=torch.float32, device=t.device)
    lib.call("hpvg_channel_sum", _ptr(t), fmt_of(t), _ptr(out), n, c, d * h * w, _stream())
    return out


# ---------------------------------------------------------------------------------------------------------------
# convolution family (double differentiable)
# ---------------------------------------------------------------------------------------------------------------
class ChainLink:
    """Handshake between two consecutive LeakyReLU conv blocks of a chain (critic, encoder features) for the plain
    first-order backward.  The consumer's data-gradient launch applies the producer's LeakyReLU derivative in its epilogue
    (mask read from the producer's stored activation = the consumer's input) and, if the producer has a bias, sums the
    result per channel there: aten::leaky_relu_backward and the bias reduction cost no launch and no extra pass over the
    8.4 MB gradient.  The producer's backward then finds `premasked` set and uses the gradient as is.  Valid only when the
    producer's output has exactly one consumer (how _run_sn_chain builds the chain); never used under create_graph=True,
    where the differentiable LReluBwd / ConvDgrad nodes are needed."""
    __slots__ = ("slope", "want_gb", "premasked", "gb")

    def __init__(self, slope, want_gb):
        self.slope, self.want_gb, self.premasked, self.gb = slope, want_gb, False, None


class _MaskLink:
    """second-order twin of ChainLink, shared by the LReluBwd and ConvDgrad nodes that ConvFwd.backward creates in the
    gradient-penalty sweep: ConvDgrad.backward applies LReluBwd.backward's mask in its own epilogue"""
    __slots__ = ("y", "slope", "premasked")

    def __init__(self, y, slope):
        self.y, self.slope, self.premasked = y, slope, False


_FUSE_MASK = [os.environ.get('HPVG_FUSE_MASK', '1') != '0']      # development switch (tests compare the fused and the unfused backward)


class ConvFwd(Function):
    """y = [lrelu](conv(x, w) + b); optional BatchNorm sums of y accumulate into `stats` (a side output).
    in_link: x is the LeakyReLU output of the block owning that ChainLink; out_link: this block's own link."""

    @staticmethod
    def forward(ctx, x, w, bias, pad, out_wide, act_slope, stats, in_link=None, out_link=None, token=None):
        y = conv_raw(x, w, bias, pad, False, out_wide, act_slope=act_slope, stats=stats)
        ctx.pad, ctx.act_slope, ctx.has_bias = pad, act_slope, bias is not None
        ctx.in_link, ctx.out_link, ctx.token = in_link, out_link, token
        ctx.save_for_backward(x, w, y if act_slope is not None else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        want_gb = ctx.has_bias and ctx.needs_input_grad[2] and not _input_only()
        plain = not torch.is_grad_enabled()          # first-order backward without create_graph
        gx = gw = gb = None
        out_link = ctx.out_link
        if out_link is not None and out_link.premasked:
            # the consumer's dgrad epilogue already applied this block's LeakyReLU derivative (and summed the bias gradient)
            gz, out_link.premasked = gy.contiguous(), False
            if want_gb and out_link.gb is not None:
                gb = out_link.gb
            out_link.gb = None
        elif ctx.act_slope is not None:
            if want_gb and _chsum_fusable(y.shape[-1]):
                gz, gb = LReluBwd.apply(gy.contiguous(), y, ctx.act_slope, True)    # bias gradient from the same pass
            else:
                link = _MaskLink(y, ctx.act_slope) if (_FUSE_MASK[0] and not plain and _input_only() and is_wide(x)) else None
                gz = LReluBwd.apply(gy.contiguous(), y, ctx.act_slope, False, link)
                if link is not None and ctx.needs_input_grad[0]:
                    gx = ConvFwd._dgrad_node(ctx, gz, w, is_wide(x), link, plain)
        else:
            gz = gy.contiguous()
        if ctx.needs_input_grad[0] and gx is None:
            in_link = ctx.in_link
            if _FUSE_MASK[0] and plain and in_link is not None and is_wide(x) and is_wide(gz):
                stats = None
                if in_link.want_gb and not _input_only():
                    stats = zeros_small(2 * x.shape[-1], x.device)
                gx = conv_raw(gz, w, None, 2 - ctx.pad, True, True, stats=stats, mask_src=x, mask_slope=in_link.slope)
                in_link.premasked = True
                in_link.gb = stats[:x.shape[-1]] if stats is not None else None
            else:
                gx = ConvFwd._dgrad_node(ctx, gz, w, is_wide(x), None, plain)
        if not _input_only():
            if ctx.needs_input_grad[1]:
                if ctx.token is not None and plain:
                    ctx.token.slot = (x, gz, ctx.pad)            # WeightProxy.backward computes it on the side stream
                    gw = torch.empty(tuple(w.shape), dtype=torch.float32, device=w.device)
                else:
                    gw = ConvWgrad.apply(x, gz, ctx.pad, tuple(w.shape))
            if want_gb and gb is None:
                gb = ChannelSum.apply(gz)
        return gx, gw, gb, None, None, None, None, None, None, None


    @staticmethod
    def _dgrad_node(ctx, gz, w, out_wide, link, plain):
        """the differentiable data-gradient node.  With a deferred weight (ctx.token) and a differentiated sweep (not plain:
        the gradient penalty's create_graph pass) the node gets its OWN proxy of the weight in front of this block's proxy: a
        proxy must have exactly one consumer, because what reaches it is a placeholder, not a gradient."""
        token = None
        if ctx.token is not None and not plain:
            w, token = deferred_weight(ctx.token.base)
        return ConvDgrad.apply(gz, w, ctx.pad, out_wide, link, token)


class ConvDgrad(Function):
    """gx = data gradient of conv(x, w) w.r.t. x given gz (a forward-type conv with flipped/transposed weights)."""

    @staticmethod
    def forward(ctx, gz, w, pad, out_wide, link=None, token=None):
        gx = conv_raw(gz, w, None, 2 - pad, True, out_wide)
        ctx.pad, ctx.link, ctx.token = pad, link, token
        ctx.save_for_backward(gz, w)
        return gx

    @staticmethod
    def backward(ctx, ggx):
        gz, w = ctx.saved_tensors
        ggx = ggx.contiguous()
        g_gz = g_w = None
        link = ctx.link
        if ctx.needs_input_grad[0]:
            if link is not None and not torch.is_grad_enabled() and is_wide(ggx) and is_wide(gz):
                # gz came out of LReluBwd(., link.y): its backward is the same mask, applied here in the conv epilogue
                g_gz = conv_raw(ggx, w, None, ctx.pad, False, True, mask_src=link.y, mask_slope=link.slope)
                link.premasked = True
            else:
                g_gz = ConvFwd.apply(ggx, w, None, ctx.pad, is_wide(gz), None, None)
        if ctx.needs_input_grad[1] and not _input_only():
            if ctx.token is not None and not torch.is_grad_enabled():
                ctx.token.slot = (ggx, gz, ctx.pad)              # WeightProxy.backward computes it on the side stream
                g_w = torch.empty(tuple(w.shape), dtype=torch.float32, device=w.device)
            else:
                g_w = ConvWgrad.apply(ggx, gz, ctx.pad, tuple(w.shape))
        return g_gz, g_w, None, None, None, None


class ConvWgrad(Function):
    """gw[co][ci][k] = sum gz[.., co] * x[.. + k - pad, ci]"""

    @staticmethod
    def forward(ctx, x, gz, pad, wshape):
        gw, _ = wgrad_raw(x, gz, pad, wshape)
        ctx.pad = pad
        ctx.save_for_backward(x, gz)
        return gw

    @staticmethod
    def backward(ctx, ggw):
        x, gz = ctx.saved_tensors
        ggw = ggw.contiguous()
        g_x = g_gz = None
        if ctx.needs_input_grad[0]:
            g_x = ConvDgrad.apply(gz, ggw, ctx.pad, is_wide(x))
        if ctx.needs_input_grad[1]:
            g_gz = ConvFwd.apply(x, ggw, None, ctx.pad, is_wide(gz), None, None)
        return g_x, g_gz, None, None


def _chsum_fusable(c):
    """channel counts for which the elementwise backward kernels can also emit per-channel sums (C/8 a power of two)"""
    return c % 8 == 0 and c <= 256 and (256 % (c // 8)) == 0


class LReluBwd(Function):
    """gz = gy * lrelu'(y) with the derivative read from the saved activation output (sign(y) == sign(pre-activation)).
    With want_sum the kernel also returns the per-channel sum of gz (the bias gradient of the convolution in front)."""

    @staticmethod
    def forward(ctx, gy, y, slope, want_sum=False, link=None):
        _require_cuda(gy, y)
        if not (is_wide(gy) and is_wide(y)):
            raise TypeError("LReluBwd works on wide (bf16 NDHWC) tensors")
        gz = torch.empty_like(gy)
        c = gy.shape[-1]
        gb = torch.empty((c,), dtype=torch.float32, device=gy.device) if want_sum else None
        lib.call("hpvg_lrelu_bwd", _ptr(gy), _ptr(y), _ptr(gz), gy.numel(), float(slope), c, _ptr(gb), _stream())
        ctx.slope, ctx.want_sum, ctx.link = slope, want_sum, link
        ctx.save_for_backward(y)
        if want_sum:
            ctx.mark_non_differentiable(gb)
            return gz, gb
        return gz

    @staticmethod
    def backward(ctx, ggz, *unused):
        link = ctx.link
        if link is not None and link.premasked:
            link.premasked = False
            return ggz, None, None, None, None        # ConvDgrad.backward applied this mask in its epilogue
        (y,) = ctx.saved_tensors
        return LReluBwd.apply(ggz.contiguous(), y, ctx.slope, False), None, None, None, None


class ChannelSum(Function):
    """bias gradient: per-channel sum of a wide or thin tensor"""

    @staticmethod
    def forward(ctx, t):
        ctx.shape, ctx.wide = tuple(t.shape), is_wide(t)
        return channel_sum(t.contiguous())

    @staticmethod
    def backward(ctx, g):
        if ctx.wide:
            return g.to(torch.bfloat16).view(1, 1, 1, 1, -1).expand(ctx.shape).contiguous()
        return g.view(1, -1, 1, 1, 1).expand(ctx.shape).contiguous()


def conv(x, w, bias, pad, out_wide, act_slope=None, stats=None, in_link=None, out_link=None, token=None):
    return ConvFwd.apply(x, w, bias, pad, out_wide, act_slope, stats, in_link, out_link, token)


# ---------------------------------------------------------------------------------------------------------------
# BatchNorm(train) + LeakyReLU on the conv output whose per-channel sums the conv epilogue already produced
# ---------------------------------------------------------------------------------------------------------------
class BnLrelu(Function):
    @staticmethod
    def forward(ctx, y, stats, gamma, beta, running_mean, running_var, nbt, momentum, eps, slope):
        _require_cuda(y, stats, gamma, beta)
        n, c, d, h, w = dims_of(y)
        nvox = n * d * h * w
        dev = y.device
        scale_shift = torch.empty((2 * c,), dtype=torch.float32, device=dev)
        mean_invstd = torch.empty((2 * c,), dtype=torch.float32, device=dev)
        lib.call("hpvg_bn_finalize", _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), _ptr(nbt),
                 float(momentum), float(eps), nvox, _ptr(scale_shift), _ptr(mean_invstd), c, _stream())
        out = torch.empty_like(y)
        lib.call("hpvg_bn_apply_lrelu", _ptr(y), _ptr(scale_shift), _ptr(out), nvox, c, float(slope), _stream())
        ctx.slope, ctx.c, ctx.nvox = slope, c, nvox
        ctx.save_for_backward(y, scale_shift, mean_invstd)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        y, scale_shift, mean_invstd = ctx.saved_tensors
        gout = gout.contiguous()
        c, nvox = ctx.c, ctx.nvox
        sums = torch.empty((3 * c,), dtype=torch.float32, device=y.device)
        lib.call("hpvg_bn_lrelu_bwd_reduce", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), nvox, c,
                 float(ctx.slope), None, _stream())
        gy = torch.empty_like(y)
        dgamma = torch.empty((c,), dtype=torch.float32, device=y.device)
        dbeta = torch.empty((c,), dtype=torch.float32, device=y.device)
        lib.call("hpvg_bn_lrelu_bwd_apply", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), _ptr(gy),
                 _ptr(dgamma), _ptr(dbeta), nvox, c, float(ctx.slope), 0, None, _stream())
        return gy, None, dgamma, dbeta, None, None, None, None, None, None


_BN_TRACK = [True]


class bn_running_stats:
    """Context: BatchNorm layers inside it do not advance running_mean / running_var / num_batches_tracked.  Used by the
    multi-stream sampler, where concurrent draws would race on those buffers; they are never read by the path (the
    generator is never put in eval mode)."""

    def __init__(self, track):
        self.track = track

    def __enter__(self):
        self.prev = _BN_TRACK[0]
        _BN_TRACK[0] = self.track

    def __exit__(self, *a):
        _BN_TRACK[0] = self.prev


# ---------------------------------------------------------------------------------------------------------------
# weight gradients on a second stream.  A weight gradient is needed only by the optimizer, while the data gradient is on
# the critical path of the backward sweep.  WeightProxy is an identity node created under the side stream's context, so
# autograd runs ITS backward on that stream (with the engine's own event synchronisation before it and before the
# gradient is accumulated); the convolution node only deposits (x, gy) and returns a placeholder.
# ---------------------------------------------------------------------------------------------------------------
_WGRAD_STREAM = [None]


class wgrad_stream:
    """Context: convolution blocks built inside it compute their weight gradients on a side stream during backward.  `stream` may
    be a list: the blocks are then dealt out round-robin, so that the weight-gradient chains (kernel -> split-K reduction ->
    gradient accumulation, each a few small launches) of different layers run side by side instead of queueing on one stream —
    measured on B200 (configs[1]): with one stream the queue drained ~150 us after the data-gradient sweep had finished."""

    def __init__(self, stream):
        self.streams = [s for s in stream if s is not None] if isinstance(stream, (list, tuple)) else ([stream] if stream is not None else [])

    def __enter__(self):
        self.prev = _WGRAD_STREAM[0]
        _WGRAD_STREAM[0] = _StreamRing(self.streams) if self.streams else None

    def __exit__(self, *a):
        _WGRAD_STREAM[0] = self.prev


class _StreamRing:
    def __init__(self, streams):
        self.streams, self.i = streams, 0

    def next(self):
        s = self.streams[self.i]
        self.i = (self.i + 1) % len(self.streams)
        return s


def _next_wgrad_stream():
    ring = _WGRAD_STREAM[0]
    return ring.next() if ring is not None else None


class _Deferred:
    __slots__ = ("slot", "base")

    def __init__(self, base=None):
        self.slot = None
        self.base = base       # the weight in front of the proxy (for nodes that must not share this proxy: see ConvFwd.backward)


class WeightProxy(Function):
    @staticmethod
    def forward(ctx, w, token):
        ctx.token = token
        ctx.wshape = tuple(w.shape)
        return w.view_as(w)

    @staticmethod
    @once_differentiable
    def backward(ctx, placeholder):
        job, ctx.token.slot = ctx.token.slot, None
        if job is None:
            return placeholder, None      # nothing deposited (the consumer computed the gradient itself): plain identity
        x, gy, pad = job
        side = torch.cuda.current_stream()
        x.record_stream(side)
        gy.record_stream(side)
        gw, _ = wgrad_raw(x, gy, pad, ctx.wshape)
        return gw, None


# On by default (HPVG_CRITIC_WSIDE=0 turns it off; measured on B200, config 2: 5.38 -> 5.22 ms per iteration): the same
# deferral for the spectral-norm blocks (critic, encoder features).  Their weight gradients sit between the data-gradient
# launches of the backward sweep without feeding them: ~10 wgrad_tc launches of the critic's real / fake passes per iteration.
# In calc_gradient_penalty's critic pass a weight is referenced by its ConvFwd node AND by the ConvDgrad node that the
# create_graph sweep creates; a proxy's placeholder gradient must not meet a real one, so that ConvDgrad node gets a proxy of
# its own (ConvFwd._dgrad_node).  no_wgrad_proxy() switches the deferral off for a region.
_CRITIC_WSIDE = [os.environ.get('HPVG_CRITIC_WSIDE', '1') == '1']
_NO_PROXY = [False]


class no_wgrad_proxy:
    def __enter__(self):
        self.prev = _NO_PROXY[0]
        _NO_PROXY[0] = True

    def __exit__(self, *a):
        _NO_PROXY[0] = self.prev


_SMALL_GRADS_SIDE = [os.environ.get('HPVG_SMALL_GRADS_SIDE', '0') == '1']      # measured slower (3.74 vs 3.63 ms per iteration): off


class SideGrad(Function):
    """identity; created under a side stream so that autograd accumulates the gradient of `p` on that stream"""

    @staticmethod
    def forward(ctx, p):
        return p.view_as(p)

    @staticmethod
    def backward(ctx, g):
        return g


def deferred_weight(w):
    """-> (w', token): w' = w behind a WeightProxy created under the side stream, or (w, None) when the deferral does not apply"""
    if not _CRITIC_WSIDE[0] or _NO_PROXY[0] or _WGRAD_STREAM[0] is None or not torch.is_grad_enabled() or not w.requires_grad:
        return w, None
    side = _next_wgrad_stream()
    token = _Deferred(w)
    with torch.cuda.stream(side):
        w = WeightProxy.apply(w, token)
    return w, token


# conv + BatchNorm + LeakyReLU in one launch (hpvg_conv_bn_lrelu_fused).  The kernel spins on a grid-wide barrier, so two of
# them must never be partially resident at the same time: the training iteration serialises the generator's passes (they
# share BatchNorm buffers), the multi-stream sampler switches the fusion off (fused_bn(False)).
_FUSED_BN = [os.environ.get('HPVG_FUSED_BN', '1') != '0']
_FUSED_BN_BWD = [os.environ.get('HPVG_FUSED_BN_BWD', '1') != '0']      # BatchNorm backward: reduce + apply in one launch (grid barrier)


class fused_bn:
    """Context: allow / forbid the one-launch conv + BatchNorm + LeakyReLU kernel inside the block"""

    def __init__(self, on):
        self.on = bool(on)

    def __enter__(self):
        self.prev = _FUSED_BN[0]
        _FUSED_BN[0] = self.on

    def __exit__(self, *a):
        _FUSED_BN[0] = self.prev


def _fused_bn_ok(x, w, pad):
    if not _FUSED_BN[0] or not is_wide(x) or w.dim() != 5 or w.shape[0] != 64 or w.shape[1] != 64:
        return False
    n, c, d, h, wd = dims_of(x)
    return bool(lib.load().hpvg_conv_bn_lrelu_fused_supported(n, 64, 64, d, h, wd, 3, int(pad)))


_BN_LOG = [None]


class bn_stat_log:
    """`with bn_stat_log() as log:` — BatchNorm blocks inside the block do NOT touch running_mean / running_var /
    num_batches_tracked; each appends what its update needs to `log.entries`.  `flush_bn_stats(entries)` applies them later, in
    list order, with one launch.  hpvg.train.ScaleTrainer runs the generator's 'rec' and 'rand' passes on two streams inside two
    logs and flushes rec + rand in the reference's order."""

    def __init__(self):
        self.entries = []

    def __enter__(self):
        self.prev = _BN_LOG[0]
        _BN_LOG[0] = self
        return self

    def __exit__(self, *a):
        _BN_LOG[0] = self.prev


def flush_bn_stats(entries):
    """apply logged running-statistics updates in order (reference semantics of nn.BatchNorm in training mode)"""
    for i in range(0, len(entries), lib.BN_LOG_MAX):
        chunk = entries[i:i + lib.BN_LOG_MAX]
        lib.call("hpvg_bn_running_update_batched", len(chunk), lib.ptr_array_opt([e[0] for e in chunk]), lib.ptr_array_opt([e[1] for e in chunk]),
                 lib.ptr_array_opt([e[2] for e in chunk]), lib.ptr_array([e[3] for e in chunk]), lib.longlong_array([e[4] for e in chunk]),
                 lib.int_array([e[5] for e in chunk]), lib.float_array([e[6] for e in chunk]), lib.float_array([e[7] for e in chunk]), _stream())


class ConvBnLrelu(Function):
    """ConvBlock3D/2D as ONE autograd node (reference modules/networks_3d.py:48-56).
    forward : conv (+bias) with BatchNorm sums fused into its epilogue -> finalize + normalise + affine + LeakyReLU (1 launch)
    backward: BN/LReLU backward reduce -> apply (also emits the conv-bias gradient) -> data gradient -> weight gradient.
    First-order only: BatchNorm blocks live in the generators, which are never differentiated twice."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, running_mean, running_var, nbt, pad, momentum, eps, slope, token=None):
        _require_cuda(x, w, gamma, beta)
        ctx.token = token
        cout = w.shape[0]
        track = _BN_TRACK[0]
        log = _BN_LOG[0] if track else None
        if log is not None:
            track = False          # the update is logged and applied later, in order (bn_stat_log)
        need_bwd = any(ctx.needs_input_grad[:5])       # all False when the caller runs under no_grad
        mask = None
        if _fused_bn_ok(x, w, pad):
            # ONE launch: the convolution's accumulators stay in TMEM across a grid barrier on the channel sums; statistics,
            # normalisation and the LeakyReLU sign all come from the fp32 values (hpvg_conv_bn_lrelu_fused)
            x = x.contiguous()
            n, cin, d, h, wd = dims_of(x)
            kd = _kd_of(w)
            packed = packed_for(w.contiguous(), cout, cin, kd * 9, False, cout)
            do, ho, wo = d + 2 * pad - 2, h + 2 * pad - 2, wd + 2 * pad - 2
            c, nvox = cout, n * do * ho * wo
            stats = zeros_small(2 * cout + 32, x.device)        # sums + the grid barrier's arrival counter
            out = _empty(n, cout, do, ho, wo, True, x.device)
            y = torch.empty_like(out) if need_bwd else None
            mask = torch.empty((nvox * (cout // 8),), dtype=torch.uint8, device=x.device) if need_bwd else None
            scale_shift = torch.empty((2 * c,), dtype=torch.float32, device=x.device)
            mean_invstd = torch.empty((2 * c,), dtype=torch.float32, device=x.device)
            lib.call("hpvg_conv_bn_lrelu_fused", _ptr(x), _ptr(packed), _ptr(bias.contiguous() if bias is not None else None), _ptr(y),
                     _ptr(out), n, cin, cout, d, h, wd, kd, pad, float(slope), _ptr(gamma), _ptr(beta),
                     _ptr(running_mean if track else None), _ptr(running_var if track else None), _ptr(nbt if track else None),
                     float(momentum), float(eps), _ptr(stats), _ptr(scale_shift), _ptr(mean_invstd), _ptr(mask), _stream())
        else:
            stats = zeros_small(2 * cout, x.device)
            y = conv_raw(x, w, bias, pad, False, True, stats=stats)
            n, c, d, h, wd = dims_of(y)
            nvox = n * d * h * wd
            scale_shift = torch.empty((2 * c,), dtype=torch.float32, device=y.device)
            mean_invstd = torch.empty((2 * c,), dtype=torch.float32, device=y.device)
            out = torch.empty_like(y)
            lib.call("hpvg_bn_finalize_apply_lrelu", _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), _ptr(running_mean if track else None),
                     _ptr(running_var if track else None), _ptr(nbt if track else None), float(momentum), float(eps), _ptr(scale_shift),
                     _ptr(mean_invstd), _ptr(out), nvox, c, float(slope), _stream())
        if log is not None:
            log.entries.append((running_mean, running_var, nbt, mean_invstd, nvox, c, float(momentum), float(eps)))
        ctx.pad, ctx.slope, ctx.c, ctx.nvox, ctx.has_bias = pad, slope, c, nvox, bias is not None
        ctx.save_for_backward(x, w, y, scale_shift, mean_invstd, mask)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        x, w, y, scale_shift, mean_invstd, mask = ctx.saved_tensors
        gout = gout.contiguous()
        c, nvox = ctx.c, ctx.nvox
        want_gb = ctx.has_bias and ctx.needs_input_grad[2]
        fuse_gb = want_gb and _chsum_fusable(c)
        gy = torch.empty_like(y)
        dgamma = torch.empty((c,), dtype=torch.float32, device=y.device)
        dbeta = torch.empty((c,), dtype=torch.float32, device=y.device)
        if _FUSED_BN_BWD[0] and lib.load().hpvg_bn_lrelu_bwd_fused_supported(nvox, c):
            # reduce + apply in ONE launch: y and gout are read once into shared memory, the sums cross a grid barrier
            sums = zeros_small(3 * c + 32, y.device)
            lib.call("hpvg_bn_lrelu_bwd_fused", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), _ptr(gy),
                     _ptr(dgamma), _ptr(dbeta), nvox, c, float(ctx.slope), int(fuse_gb), _ptr(mask), _stream())
        else:
            sums = torch.empty((3 * c,), dtype=torch.float32, device=y.device)
            lib.call("hpvg_bn_lrelu_bwd_reduce", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), nvox, c,
                     float(ctx.slope), _ptr(mask), _stream())
            lib.call("hpvg_bn_lrelu_bwd_apply", _ptr(y), _ptr(gout), _ptr(scale_shift), _ptr(mean_invstd), _ptr(sums), _ptr(gy),
                     _ptr(dgamma), _ptr(dbeta), nvox, c, float(ctx.slope), int(fuse_gb), _ptr(mask), _stream())
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = conv_raw(gy, w, None, 2 - ctx.pad, True, is_wide(x))
        if ctx.needs_input_grad[1]:
            if ctx.token is not None:
                ctx.token.slot = (x, gy, ctx.pad)            # WeightProxy.backward computes it on the side stream
                gw = torch.empty(tuple(w.shape), dtype=torch.float32, device=w.device)
            else:
                gw, _ = wgrad_raw(x, gy, ctx.pad, tuple(w.shape))
        if want_gb:
            gb = sums[2 * c:3 * c] if fuse_gb else channel_sum(gy)
        return gx, gw, gb, dgamma, dbeta, None, None, None, None, None, None, None, None


_BN_PER_SAMPLE = [False]


class bn_per_sample:
    """`with bn_per_sample(True):` — BatchNorm blocks of a no-grad forward normalise every sample of the batch with its OWN
    statistics, i.e. a batch-B forward computes what B batch-1 forwards compute (the reference generates each draw with batch
    size 1 and keeps G in train mode, train_video.py:226-235).  Running statistics are not advanced in this mode."""

    def __init__(self, on):
        self.on = bool(on)

    def __enter__(self):
        self.prev = _BN_PER_SAMPLE[0]
        _BN_PER_SAMPLE[0] = self.on

    def __exit__(self, *a):
        _BN_PER_SAMPLE[0] = self.prev


def _conv_bn_lrelu_per_sample(x, w, bias, gamma, beta, pad, eps, slope):
    _require_cuda(x, w, gamma, beta)
    n = x.shape[0]
    cout = w.shape[0]
    if not per_sample_stats_supported(x, w):
        # channel counts without a per-sample kernel: sample by sample through the same kernels
        outs = []
        for i in range(n):
            stats = torch.zeros((2 * cout,), dtype=torch.float32, device=x.device)
            y = conv_raw(x[i:i + 1].contiguous(), w, bias, pad, False, True, stats=stats)
            o = torch.empty_like(y)
            _, c, d, h, wd = dims_of(y)
            lib.call("hpvg_bn_apply_lrelu_per_sample", _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), float(eps), _ptr(o), 1,
                     d * h * wd, c, float(slope), _stream())
            outs.append(o)
        return torch.cat(outs, 0)
    stats = zeros_small(n * 2 * cout, x.device).view(n, 2 * cout)
    y = conv_raw(x, w, bias, pad, False, True, stats=stats, stats_per_sample=True)
    _, c, d, h, wd = dims_of(y)
    out = torch.empty_like(y)
    lib.call("hpvg_bn_apply_lrelu_per_sample", _ptr(y), _ptr(stats), _ptr(gamma), _ptr(beta), float(eps), _ptr(out), n, d * h * wd, c,
             float(slope), _stream())
    return out


_TWICE = [False]


class twice_differentiable:
    """Context: ConvBlocks with BatchNorm built inside it can be differentiated twice.  calc_gradient_penalty wraps the critic's
    pass over the interpolates in it: the fused ConvBnLrelu node is first-order only (BatchNorm blocks live in the generators,
    which never see a double backward), but WDiscriminatorBaselines (reference modules/networks_3d.py:184-210) is a critic WITH
    BatchNorm.  In this mode the convolution stays on the library's double-differentiable family (ConvFwd / ConvDgrad /
    ConvWgrad) and the BatchNorm + LeakyReLU of the block is composed of element-wise torch operations, whose double
    backward autograd derives."""

    def __enter__(self):
        self.prev = _TWICE[0]
        _TWICE[0] = True

    def __exit__(self, *a):
        _TWICE[0] = self.prev


def _conv_bn_lrelu_twice(x, w, bias, gamma, beta, running_mean, running_var, nbt, pad, momentum, eps, slope):
    y = ConvFwd.apply(x, w, bias, pad, True, None, None).float()             # wide [N,D,H,W,C] -> fp32 for the statistics
    dims = (0, 1, 2, 3)
    mean = y.mean(dims)
    var = y.var(dims, unbiased=False)
    out = torch.nn.functional.leaky_relu((y - mean) * torch.rsqrt(var + eps) * gamma + beta, slope)
    if _BN_TRACK[0] and running_mean is not None:
        with torch.no_grad():
            count = y.numel() // y.shape[-1]
            running_mean.mul_(1 - momentum).add_(momentum * mean.detach())
            running_var.mul_(1 - momentum).add_(momentum * var.detach() * (count / max(count - 1, 1)))
            nbt.add_(1)
    return out.to(torch.bfloat16)


def conv_bn_lrelu(x, w, bias, gamma, beta, running_mean, running_var, nbt, pad, momentum=0.1, eps=1e-5, slope=0.2):
    """ConvBlock3D/2D (reference modules/networks_3d.py:48-56) through the fused node"""
    if _TWICE[0] and torch.is_grad_enabled():
        return _conv_bn_lrelu_twice(x, w, bias, gamma, beta, running_mean, running_var, nbt, pad, momentum, eps, slope)
    if _BN_PER_SAMPLE[0]:
        if torch.is_grad_enabled() and (x.requires_grad or w.requires_grad):
            raise lib.HpvgError("per-sample BatchNorm statistics are an inference mode: run it under torch.no_grad()")
        return _conv_bn_lrelu_per_sample(x, w, bias, gamma, beta, pad, eps, slope)
    token = None
    if _WGRAD_STREAM[0] is not None and torch.is_grad_enabled() and w.requires_grad:
        side = _next_wgrad_stream()
        token = _Deferred()
        with torch.cuda.stream(side):
            w = WeightProxy.apply(w, token)
            if _SMALL_GRADS_SIDE[0]:
                # bias, gamma, beta: their gradients come out of the BatchNorm-backward kernel on the data-gradient chain, and a layer used
                # by both generator passes accumulates them with one tiny add each — inside that chain.  Behind an identity node created
                # under the side stream the accumulation happens there (the engine orders it after the producing kernel by an event).
                bias, gamma, beta = (SideGrad.apply(t) if (t is not None and t.requires_grad) else t for t in (bias, gamma, beta))
    return ConvBnLrelu.apply(x, w, bias, gamma, beta, running_mean, running_var, nbt, pad, momentum, eps, slope, token)


# ---------------------------------------------------------------------------------------------------------------
# thin-tensor ops: resize(+noise), tanh(+residual), KL, GP penalty, format conversion
# ---------------------------------------------------------------------------------------------------------------
class UpsampleLinear(Function):
    """F.interpolate(mode='trilinear'|'bilinear', align_corners=True) [+ amp*noise] (reference utils/images.py:9-26)."""

    @staticmethod
    def forward(ctx, x, size, noise, amp):
        _require_cuda(x, noise)
        x = x.contiguous()
        n, c, d, h, w = dims_of(x)
        do, ho, wo = size
        out = torch.empty((n, c, do, ho, wo), dtype=torch.float32, device=x.device)
        if noise is not None:
            noise = noise.contiguous()
            if tuple(noise.shape) != tuple(out.shape):
                raise ValueError("noise shape %s != output shape %s" % (tuple(noise.shape), tuple(out.shape)))
        lib.call("hpvg_upsample_linear_fwd", _ptr(x), _ptr(out), _ptr(noise), float(amp), n * c, d, h, w, do, ho, wo, _stream())
        ctx.in_shape = (n, c, d, h, w)
        ctx.size = (do, ho, wo)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        n, c, d, h, w = ctx.in_shape
        do, ho, wo = ctx.size
        gout = gout.contiguous()
        gx = torch.empty((n, c, d, h, w), dtype=torch.float32, device=gout.device)
        lib.call("hpvg_upsample_linear_bwd", _ptr(gout), _ptr(gx), n * c, d, h, w, do, ho, wo, _stream())
        return gx, None, None, None


class UpsampleWide(Function):
    """the same resize on a wide tensor [N,D,H,W,C] (GeneratorCSG resizes its nfc-channel feature maps, reference
    modules/networks_3d.py:252-261); `noise`: float32 [N,C,Do,Ho,Wo] as the reference draws it, or None"""

    @staticmethod
    def forward(ctx, x, size, noise, amp):
        _require_cuda(x, noise)
        if not is_wide(x):
            raise TypeError("UpsampleWide takes a wide (bf16 NDHWC) tensor")
        x = x.contiguous()
        n, c, d, h, w = dims_of(x)
        do, ho, wo = size
        out = _empty(n, c, do, ho, wo, True, x.device)
        if noise is not None:
            noise = noise.contiguous()
            if tuple(noise.shape) != (n, c, do, ho, wo) or noise.dtype != torch.float32:
                raise ValueError("noise must be float32 of shape %s, got %s %s" % ((n, c, do, ho, wo), noise.dtype, tuple(noise.shape)))
        lib.call("hpvg_upsample_linear_wide_fwd", _ptr(x), _ptr(out), _ptr(noise), float(amp), n, c, d, h, w, do, ho, wo, _stream())
        ctx.geom = (n, c, d, h, w, do, ho, wo)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        n, c, d, h, w, do, ho, wo = ctx.geom
        gout = gout.contiguous()
        gx = _empty(n, c, d, h, w, True, gout.device)
        lib.call("hpvg_upsample_linear_wide_bwd", _ptr(gout), _ptr(gx), n, c, d, h, w, do, ho, wo, _stream())
        return gx, None, None, None


def pad_wide_raw(x, pad):
    n, c, d, h, w = dims_of(x)
    out = _empty(n, c, d + 2 * pad, h + 2 * pad, w + 2 * pad, True, x.device)
    lib.call("hpvg_pad_wide", _ptr(x.contiguous()), _ptr(out), n, c, d, h, w, int(pad), _stream())
    return out


class PadWide(Function):
    """F.pad(x, (pad,) * 6) with zeros on a wide tensor (reference modules/networks_3d.py:205,248,264); the adjoint crops."""

    @staticmethod
    def forward(ctx, x, pad):
        _require_cuda(x)
        if not is_wide(x):
            raise TypeError("PadWide takes a wide (bf16 NDHWC) tensor")
        ctx.pad = int(pad)
        return pad_wide_raw(x, ctx.pad)

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return pad_wide_raw(g.contiguous(), -ctx.pad), None


class AddWide(Function):
    """a + b on wide tensors: the stage residual of GeneratorCSG (reference modules/networks_3d.py:265)"""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        if not (is_wide(a) and is_wide(b)) or a.shape != b.shape:
            raise ValueError("AddWide takes two wide tensors of equal shape")
        a, b = a.contiguous(), b.contiguous()
        out = torch.empty_like(a)
        lib.call("hpvg_add_wide", _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream())
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        return g, g


class TanhAdd(Function):
    """tanh(a + b) (b optional) on thin tensors (reference modules/networks_3d.py:377,404)."""

    @staticmethod
    def forward(ctx, a, b):
        _require_cuda(a, b)
        a = a.contiguous()
        b = b.contiguous() if b is not None else None
        out = torch.empty_like(a)
        lib.call("hpvg_tanh_add_fwd", _ptr(a), _ptr(b), _ptr(out), a.numel(), _stream())
        ctx.has_b = b is not None
        ctx.save_for_backward(out)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (out,) = ctx.saved_tensors
        g = torch.empty_like(out)
        lib.call("hpvg_tanh_bwd", _ptr(gout.contiguous()), _ptr(out), _ptr(g), out.numel(), _stream())
        return g, (g if ctx.has_b else None)


class Reparam(Function):
    """z = eps * exp(0.5*logvar) + mu on wide mu/logvar with thin eps (reference modules/networks_3d.py:29-35)."""

    @staticmethod
    def forward(ctx, mu, logvar, eps):
        _require_cuda(mu, logvar, eps)
        n, c, d, h, w = dims_of(mu)
        z = torch.empty_like(mu)
        lib.call("hpvg_reparam_fwd", _ptr(mu), _ptr(logvar), _ptr(eps), _ptr(z), n, c, d * h * w, _stream())
        ctx.save_for_backward(logvar, eps)
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, gz):
        logvar, eps = ctx.saved_tensors
        n, c, d, h, w = dims_of(logvar)
        gz = gz.contiguous()
        gmu = torch.empty_like(logvar)
        glv = torch.empty_like(logvar)
        lib.call("hpvg_reparam_bwd", _ptr(gz), _ptr(logvar), _ptr(eps), _ptr(gmu), _ptr(glv), n, c, d * h * w, _stream())
        return gmu, glv, None


class KlCriterion(Function):
    """mean(-0.5*(1 + logvar - mu^2 - exp(logvar))) (reference modules/losses.py:7-9)."""

    @staticmethod
    def forward(ctx, mu, logvar):
        _require_cuda(mu, logvar)
        mu, logvar = mu.contiguous(), logvar.contiguous()
        out = torch.empty((1,), dtype=torch.float32, device=mu.device)
        lib.call("hpvg_kl_fwd", _ptr(mu), _ptr(logvar), _ptr(out), mu.numel(), _stream())
        ctx.save_for_backward(mu, logvar)
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        mu, logvar = ctx.saved_tensors
        gmu, glv = torch.empty_like(mu), torch.empty_like(logvar)
        lib.call("hpvg_kl_bwd", _ptr(gout.contiguous().view(1)), _ptr(mu), _ptr(logvar), _ptr(gmu), _ptr(glv), mu.numel(), _stream())
        return gmu, glv


class GpPenalty(Function):
    """lambda * mean((||g||_2 over channels - 1)^2) (reference modules/utils.py:18)."""

    @staticmethod
    def forward(ctx, g, lam):
        _require_cuda(g)
        g = g.contiguous()
        n, c, d, h, w = dims_of(g)
        out = torch.empty((1,), dtype=torch.float32, device=g.device)
        lib.call("hpvg_gp_penalty_fwd", _ptr(g), _ptr(out), n, c, d * h * w, float(lam), _stream())
        ctx.lam = lam
        ctx.save_for_backward(g)
        return out.view(())

    @staticmethod
    @once_differentiable
    def backward(ctx, gout):
        (g,) = ctx.saved_tensors
        n, c, d, h, w = dims_of(g)
        gg = torch.empty_like(g)
        lib.call("hpvg_gp_penalty_bwd", _ptr(gout.contiguous().view(1)), _ptr(g), _ptr(gg), n, c, d * h * w, float(ctx.lam), _stream())
        return gg, None


def convert_raw(t, to_wide):
    n, c, d, h, w = dims_of(t)
    out = _empty(n, c, d, h, w, to_wide, t.device)
    lib.call("hpvg_convert_format", _ptr(t.contiguous()), fmt_of(t), _ptr(out), fmt_of(out), n, c, d * h * w, _stream())
    return out


class ToWide(Function):
    @staticmethod
    def forward(ctx, t):
        _require_cuda(t)
        return convert_raw(t, True)

    @staticmethod
    def backward(ctx, g):
        return ToThin.apply(g.contiguous())


class ToThin(Function):
    @staticmethod
    def forward(ctx, t):
        _require_cuda(t)
        return convert_raw(t, False)

    @staticmethod
    def backward(ctx, g):
        return ToWide.apply(g.contiguous())


class _GpAlpha:
    """The WGAN-GP mixing coefficient: ONE scalar per call drawn from torch's CPU generator (reference modules/utils.py:5),
    handed to the kernels through a device float.  In eager mode every call draws; under CUDA-graph capture/replay
    (hpvg.train.ScaleTrainer.capture) the owner sets `external` and calls draw() itself before each replay, so the captured
    lerp kernel reads a fresh value."""
    external = False
    tensors = {}
    generator = None      # CPU generator for the draw (None: torch's default one, as the reference); see ScaleTrainer(distributed=True)

    @classmethod
    def tensor(cls, device):
        key = (device.type, device.index)
        t = cls.tensors.get(key)
        if t is None:
            t = cls.tensors[key] = torch.zeros(1, dtype=torch.float32, device=device)
        return t

    @classmethod
    def draw(cls, device):
        t = cls.tensor(device)
        a = torch.rand(1, 1) if cls.generator is None else torch.rand(1, 1, generator=cls.generator)
        t.fill_(float(a))     # scalar travels as a kernel argument: no host buffer to race with
        return t


def gp_alpha(device):
    device = torch.device(device) if not isinstance(device, torch.device) else device
    if device.index is None:
        device = torch.device(device.type, torch.cuda.current_device())
    return _GpAlpha.tensor(device) if _GpAlpha.external else _GpAlpha.draw(device)


def lerp(a, b, alpha):
    """alpha*a + (1-alpha)*b on thin tensors, no autograd (the GP interpolates are a detached leaf); alpha is a device float"""
    _require_cuda(a, b, alpha)
    a, b = a.detach().contiguous(), b.detach().contiguous()
    out = torch.empty_like(a)
    lib.call("hpvg_lerp", _ptr(a), _ptr(b), _ptr(out), _ptr(alpha), a.numel(), _stream())
    return out


# ---------------------------------------------------------------------------------------------------------------
# spectral normalisation (legacy torch.nn.utils.spectral_norm semantics: 1 power iteration per training forward)
# ---------------------------------------------------------------------------------------------------------------
class SpectralWeight(Function):
    """w_sn = w_orig / sigma(w_orig; u, v); u and v are updated in place when `update_uv` (training mode)."""

    @staticmethod
    def forward(ctx, w_orig, u, v, update_uv, eps):
        _require_cuda(w_orig, u, v)
        w_orig = w_orig.contiguous()
        cout = w_orig.shape[0]
        k = w_orig.numel() // cout
        dev = w_orig.device
        sigma = torch.empty((1,), dtype=torch.float32, device=dev)
        w_sn = torch.empty_like(w_orig)
        scratch = torch.empty((k + cout + 4,), dtype=torch.float32, device=dev)
        lib.call("hpvg_sn_power_iter", _ptr(w_orig), _ptr(u), _ptr(v), _ptr(sigma), _ptr(w_sn), _ptr(scratch), cout, k,
                 int(bool(update_uv)), float(eps), _stream())
        # u, v are buffers mutated in place; the backward must see the values used for this sigma
        ctx.save_for_backward(w_sn, u.clone(), v.clone(), sigma)
        return w_sn

    @staticmethod
    @once_differentiable
    def backward(ctx, gw_sn):
        w_sn, u, v, sigma = ctx.saved_tensors
        cout = w_sn.shape[0]
        k = w_sn.numel() // cout
        gw = torch.empty_like(w_sn)
        scratch = torch.empty((lib.SN_DOT_PARTS,), dtype=torch.float32, device=w_sn.device)
        lib.call("hpvg_sn_backward", _ptr(gw_sn.contiguous()), _ptr(w_sn), _ptr(u), _ptr(v), _ptr(sigma), _ptr(gw), _ptr(scratch),
                 cout, k, _stream())
        return gw, None, None, None, None


class SpectralWeights(Function):
    """SpectralWeight for all spectral-norm layers of one network in 4 launches (forward) / 2 (backward) instead of 4 / 2
    per layer.  apply(update_uv, eps, w0, u0, v0, w1, u1, v1, ...) -> (w_sn0, w_sn1, ...)"""

    @staticmethod
    def forward(ctx, update_uv, eps, *tensors):
        n = len(tensors) // 3
        if n == 0 or n > lib.SN_MAX_LAYERS or len(tensors) != 3 * n:
            raise ValueError("SpectralWeights takes 1..%d (w_orig, u, v) triples" % lib.SN_MAX_LAYERS)
        ws = [tensors[3 * i].contiguous() for i in range(n)]
        us = [tensors[3 * i + 1] for i in range(n)]
        vs = [tensors[3 * i + 2] for i in range(n)]
        _require_cuda(*ws)
        dev = ws[0].device
        couts = [w.shape[0] for w in ws]
        ks = [w.numel() // w.shape[0] for w in ws]
        sigmas = torch.empty((n,), dtype=torch.float32, device=dev)
        sig = [sigmas[i:i + 1] for i in range(n)]
        outs = [torch.empty_like(w) for w in ws]
        offs, total = [], 0
        for c, k in zip(couts, ks):
            offs.append(total)
            total += k + c + 4
        scratch_all = torch.empty((total,), dtype=torch.float32, device=dev)
        scratch = [scratch_all[o:] for o in offs]
        # u, v are buffers mutated in place; the backward must see the values used for these sigmas.  In training mode the
        # kernel that normalises them writes the copies as well (no clone launches); otherwise they are cloned.
        if update_uv:
            saved = torch.empty((sum(couts) + sum(ks),), dtype=torch.float32, device=dev)
            u_saved, v_saved, off = [], [], 0
            for c, k in zip(couts, ks):
                u_saved.append(saved[off:off + c]); off += c
                v_saved.append(saved[off:off + k]); off += k
            lib.call("hpvg_sn_power_iter_batched_ex", n, lib.ptr_array(ws), lib.ptr_array(us), lib.ptr_array(vs), lib.ptr_array(sig),
                     lib.ptr_array(outs), lib.ptr_array(scratch), lib.int_array(couts), lib.int_array(ks), 1, float(eps),
                     lib.ptr_array(u_saved), lib.ptr_array(v_saved), _stream())
        else:
            lib.call("hpvg_sn_power_iter_batched", n, lib.ptr_array(ws), lib.ptr_array(us), lib.ptr_array(vs), lib.ptr_array(sig),
                     lib.ptr_array(outs), lib.ptr_array(scratch), lib.int_array(couts), lib.int_array(ks), 0, float(eps), _stream())
            u_saved, v_saved = [u.clone() for u in us], [v.clone() for v in vs]
        ctx.n, ctx.couts, ctx.ks = n, couts, ks
        ctx.save_for_backward(sigmas, *outs, *u_saved, *v_saved)
        return tuple(outs)

    @staticmethod
    @once_differentiable
    def backward(ctx, *grads):
        n = ctx.n
        saved = ctx.saved_tensors
        sigmas, outs, us, vs = saved[0], saved[1:1 + n], saved[1 + n:1 + 2 * n], saved[1 + 2 * n:1 + 3 * n]
        idx = [i for i in range(n) if grads[i] is not None and ctx.needs_input_grad[2 + 3 * i]]
        result = [None, None] + [None] * (3 * n)
        if idx:
            dev = sigmas.device
            gs = [grads[i].contiguous() for i in idx]
            gws = [torch.empty_like(outs[i]) for i in idx]
            scratch_all = torch.empty((len(idx) * lib.SN_DOT_PARTS,), dtype=torch.float32, device=dev)
            scratch = [scratch_all[j * lib.SN_DOT_PARTS:] for j in range(len(idx))]
            lib.call("hpvg_sn_backward_batched", len(idx), lib.ptr_array(gs), lib.ptr_array([outs[i] for i in idx]),
                     lib.ptr_array([us[i] for i in idx]), lib.ptr_array([vs[i] for i in idx]),
                     lib.ptr_array([sigmas[i:i + 1] for i in idx]), lib.ptr_array(gws), lib.ptr_array(scratch),
                     lib.int_array([ctx.couts[i] for i in idx]), lib.int_array([ctx.ks[i] for i in idx]), _stream())
            for j, i in enumerate(idx):
                result[2 + 3 * i] = gws[j]
        return tuple(result)


def spectral_weights(convs, eps=1e-12):
    """w_sn for a list of legacy-spectral_norm convolutions (attributes weight_orig / weight_u / weight_v), batched"""
    out = []
    for i in range(0, len(convs), lib.SN_MAX_LAYERS):
        chunk = convs[i:i + lib.SN_MAX_LAYERS]
        args = []
        for c in chunk:
            args += [c.weight_orig, c.weight_u, c.weight_v]
        out += list(SpectralWeights.apply(chunk[0].training, eps, *args))
    return out
