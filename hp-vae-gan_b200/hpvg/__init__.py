"""Host side of hpvg-b200: ctypes binding of libhpvg.so (the C-ABI kernel library for sm_100a), the
torch.autograd.Function wrappers around it, and the scale-pyramid helpers the generator calls in forward.

PyTorch is used for device memory, streams, autograd bookkeeping and torch.distributed only; every replaced
operator of the reference's hot path runs in libhpvg.so.  There is no CPU or cuDNN fallback: calling an op
without the library or without a CUDA tensor raises.
"""
from . import lib  # noqa: F401
