"""ctypes binding of libhpvg.so.  Signatures mirror include/hpvg.h one to one."""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
# HPVG_LIB: another build of the same library (A/B runs of two kernel versions); there is still no fallback
LIB_PATH = os.environ.get("HPVG_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libhpvg.so")

FMT_NCDHW_F32 = 0
FMT_NDHWC_BF16 = 1
ACT_NONE = 0
ACT_LRELU = 1
BACKEND_AUTO, BACKEND_DIRECT, BACKEND_TCGEN05 = 0, 1, 2

# name -> (restype, argtypes); the single source for the "every symbol exports" test
PROTOTYPES = {
    "hpvg_last_error": (c_char_p, []),
    "hpvg_version": (c_int, []),
    "hpvg_set_conv_backend": (c_int, [c_int]),
    "hpvg_get_conv_backend": (c_int, []),
    "hpvg_launch_count": (c_longlong, []),
    "hpvg_debug_set_clock_buffer": (c_int, [c_void_p]),
    "hpvg_profile_enable": (c_int, [c_int]),
    "hpvg_set_pdl": (c_int, [c_int]),
    "hpvg_set_conv_col_mode": (c_int, [c_int]),
    "hpvg_set_wgrad_mode": (c_int, [c_int]),
    "hpvg_profile_dump": (c_int, [c_void_p, c_int]),
    "hpvg_conv_kernel_choice": (c_int, [c_int] * 10),
    "hpvg_narrow_kernel_choice": (c_int, [c_int] * 2),
    "hpvg_conv_forward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "hpvg_conv_forward_ex": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_void_p, c_int, c_void_p, c_void_p]),
    "hpvg_bn_apply_lrelu_per_sample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_longlong, c_int,
                                               c_float, c_void_p]),
    "hpvg_conv_wgrad_workspace": (c_size_t, [c_int] * 10),
    "hpvg_conv_wgrad": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "hpvg_pack_weights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "hpvg_pack_weights_pair_batched": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hpvg_pack_weights_expand": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "hpvg_channel_sum": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_longlong, c_void_p]),
    "hpvg_bn_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_longlong,
                                 c_void_p, c_void_p, c_int, c_void_p]),
    "hpvg_bn_apply_lrelu": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]),
    "hpvg_bn_lrelu_bwd_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p,
                                         c_void_p]),
    "hpvg_bn_lrelu_bwd_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_longlong, c_int, c_float, c_int, c_void_p, c_void_p]),
    "hpvg_bn_running_update_batched": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p]),
    "hpvg_bn_lrelu_bwd_fused_supported": (c_int, [c_longlong, c_int]),
    "hpvg_bn_lrelu_bwd_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int,
                                        c_float, c_int, c_void_p, c_void_p]),
    "hpvg_conv_bn_lrelu_fused_supported": (c_int, [c_int] * 8),
    "hpvg_conv_bn_lrelu_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_int, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hpvg_bn_finalize_apply_lrelu": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float,
                                             c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]),
    "hpvg_lrelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_float, c_int, c_void_p, c_void_p]),
    "hpvg_upsample_linear_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "hpvg_upsample_linear_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hpvg_upsample_linear_wide_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                              c_int, c_void_p]),
    "hpvg_upsample_linear_wide_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hpvg_pad_wide": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hpvg_add_wide": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "hpvg_tanh_add_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "hpvg_tanh_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "hpvg_reparam_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_void_p]),
    "hpvg_reparam_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_void_p]),
    "hpvg_kl_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "hpvg_kl_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "hpvg_gp_penalty_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_float, c_void_p]),
    "hpvg_gp_penalty_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_longlong, c_float, c_void_p]),
    "hpvg_convert_format": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_longlong, c_void_p]),
    "hpvg_lerp": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]),
    "hpvg_clip_from_frames": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "hpvg_frames_to_uint8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "hpvg_frames_to_uint8_batched": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "hpvg_sn_power_iter": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                   c_void_p]),
    "hpvg_sn_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "hpvg_sn_power_iter_batched": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_float, c_void_p]),
    "hpvg_sn_power_iter_batched_ex": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "hpvg_sn_backward_batched": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p]),
    "hpvg_grad_clip_coef": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "hpvg_adam_step": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_double, c_double,
                               c_int, c_int, c_void_p, c_void_p]),
    "hpvg_peer_alloc": (c_int, [c_size_t, c_void_p]),
    "hpvg_peer_free": (c_int, [c_void_p]),
    "hpvg_peer_export": (c_int, [c_void_p, c_void_p]),
    "hpvg_peer_import": (c_int, [c_void_p, c_void_p]),
    "hpvg_peer_close": (c_int, [c_void_p]),
    "hpvg_peer_can_access": (c_int, [c_int, c_int]),
    "hpvg_peer_allreduce_avg": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_void_p]),
    "hpvg_peer_bucket_numel": (c_longlong, [c_int, c_void_p, c_int]),
    "hpvg_peer_allreduce_avg_tensors": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class HpvgError(RuntimeError):
    pass


def load():
    """Load libhpvg.so (built by hp-vae-gan_b200/build.py).  Fails loudly: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HpvgError("libhpvg.so not found at %s - run `python hp-vae-gan_b200/build.py` (needs nvcc); "
                        "hpvg-b200 has no CPU or library fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def call(name, *args):
    """Call an int-returning entry point and raise HpvgError(hpvg_last_error()) on a non-zero return."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise HpvgError("%s failed (%d): %s" % (name, rc, lib.hpvg_last_error().decode(errors="replace")))


OPT_MAX_TENSORS = 32      # HPVG_OPT_MAX_TENSORS: tensors per hpvg_adam_step / hpvg_grad_clip_coef call
OPT_BLOCKS = 64           # HPVG_OPT_BLOCKS: partial sums per tensor
OPT_STATE_FLOATS = 8
SN_MAX_LAYERS = 8
BN_LOG_MAX = 48           # HPVG_BN_LOG_MAX: entries per hpvg_bn_running_update_batched call
PEER_MAX_RANKS, PEER_HANDLE_BYTES, PEER_SIGNAL_BYTES, PEER_MAX_TENSORS = 8, 64, 8192, 64      # HPVG_PEER_*
SN_DOT_PARTS = 32     # HPVG_SN_DOT_PARTS: floats of scratch per layer of the spectral-norm backward


def ptr_array(tensors):
    """host array of device pointers (for the *_batched entry points)"""
    return (c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def longlong_array(values):
    return (c_longlong * len(values))(*[int(v) for v in values])


def int_array(values):
    return (c_int * len(values))(*[int(v) for v in values])


def float_array(values):
    return (c_float * len(values))(*[float(v) for v in values])


def ptr_array_opt(tensors):
    """host array of device pointers where entries may be None (NULL)"""
    return (c_void_p * len(tensors))(*[(None if t is None else t.data_ptr()) for t in tensors])


def launch_count():
    return int(load().hpvg_launch_count())


def set_conv_backend(backend):
    call("hpvg_set_conv_backend", int(backend))


def get_conv_backend():
    return int(load().hpvg_get_conv_backend())


PROF_KINDS = {0: "conv_tc", 1: "wgrad_tc", 2: "conv_direct", 3: "wgrad_direct", 4: "conv_expand", 5: "wgrad_narrow", 6: "conv_bn_fused", 7: "conv_thin"}


def set_conv_col_mode(mode):
    """-1 = brick or column-streaming tcgen05 kernel per layer (default), 0 = brick kernel always, 1 = column kernel whenever
    supported; returns the previous mode"""
    return int(load().hpvg_set_conv_col_mode(int(mode)))


KERNEL_DIRECT, KERNEL_EXPAND, KERNEL_TC_BRICK, KERNEL_TC_COLUMN = 0, 1, 2, 3


def conv_kernel_choice(n, cin, cout, d, h, w, kd=3, pad=1, x_wide=True, y_wide=True):
    """which kernel hpvg_conv_forward picks for this layer (host logic only: works without a GPU)"""
    return int(load().hpvg_conv_kernel_choice(n, cin, cout, d, h, w, kd, pad, FMT_NDHWC_BF16 if x_wide else FMT_NCDHW_F32,
                                              FMT_NDHWC_BF16 if y_wide else FMT_NCDHW_F32))


def set_wgrad_mode(mode):
    """0 = measured weight-gradient kernel (default), 1 = kd-stacked N = 192 form (experimental); returns the previous mode"""
    return int(load().hpvg_set_wgrad_mode(int(mode)))


def set_pdl(on):
    """programmatic dependent launch of the library's kernels on/off; returns the previous setting"""
    return int(load().hpvg_set_pdl(int(bool(on))))


def profile_enable(on):
    call("hpvg_profile_enable", int(bool(on)))


def profile_dump(max_rows=256):
    """-> list of {kind, work (FLOPs per launch), launches, ms} aggregated over the launches recorded since the last dump"""
    buf = (ctypes.c_double * (4 * max_rows))()
    n = load().hpvg_profile_dump(ctypes.cast(buf, c_void_p), max_rows)
    return [{"kind": PROF_KINDS.get(int(buf[4 * i]), str(int(buf[4 * i]))), "work": buf[4 * i + 1], "launches": int(buf[4 * i + 2]),
             "ms": buf[4 * i + 3]} for i in range(n)]
