"""Gradient buckets in NVLink peer memory (multi-GPU mode, one process per GPU).

`PeerBucket` owns this rank's flat fp32 gradient bucket and signal pad (allocations of libhpvg's own, exported to the other ranks
as CUDA IPC handles over the process group) and the mapped buckets / pads of every peer; `allreduce_tensors(grads)` launches
hpvg_peer_allreduce_avg_tensors (csrc/peer.cu) on the current stream: ONE kernel per backward — it gathers the gradients into the
bucket, exchanges, and scatters the means back — recorded into the iteration's CUDA graph like any other (`allreduce()`: the
exchange alone, on a bucket the caller packed).  It stands where nn.DataParallel's backward reduces the replicas' gradients (train_video.py:91-94, :182, :200).

`PeerBucket.create` returns None — on every rank alike — when the ranks cannot map each other's memory (no peer access between
two of the GPUs, more than 8 ranks, a rank on another host): `train.GradBucket` then keeps the NCCL all-reduce, the collective
this kernel replaces.  HPVG_PEER_ALLREDUCE=0 selects NCCL outright."""
import ctypes
import os
import socket
import sys

import torch

from . import lib


class _Raw:
    """a device allocation of libhpvg's seen through __cuda_array_interface__ (torch.as_tensor aliases it, no copy)"""

    def __init__(self, ptr, numel):
        self.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": "<f4", "data": (int(ptr), False), "version": 2}


def enabled():
    return os.environ.get("HPVG_PEER_ALLREDUCE", "1") != "0"


class PeerBucket:
    def __init__(self):
        self.numel = 0            # floats in the bucket (a multiple of 4 * world)
        self.flat = None          # this rank's bucket as a torch tensor
        self.rank = self.world = 0
        self._local = []          # (bucket, pad) pointers allocated here
        self._mapped = []         # pointers imported from the peers
        self._bufs = self._sigs = None

    @classmethod
    def create(cls, numel, device, group=None):
        """collective over `group`: every rank calls it with the same numel.  Returns a PeerBucket, or None on EVERY rank when
        any rank could not set it up (the reason goes to stderr once, on rank 0)."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        self = cls()
        self.rank, self.world = rank, world
        unit = 4 * world
        self.numel = (int(numel) + unit - 1) // unit * unit
        err, mine = None, None
        try:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("buckets must be set up outside a stream capture (run one eager iteration first)")
            if world > lib.PEER_MAX_RANKS:
                raise RuntimeError("%d ranks (the kernel pairs up to %d)" % (world, lib.PEER_MAX_RANKS))
            buf, sig = ctypes.c_void_p(), ctypes.c_void_p()
            with torch.cuda.device(device):
                lib.call("hpvg_peer_alloc", self.numel * 4, ctypes.byref(buf))
                self._local.append(buf.value)
                lib.call("hpvg_peer_alloc", lib.PEER_SIGNAL_BYTES, ctypes.byref(sig))
                self._local.append(sig.value)
                hb, hs = ctypes.create_string_buffer(lib.PEER_HANDLE_BYTES), ctypes.create_string_buffer(lib.PEER_HANDLE_BYTES)
                lib.call("hpvg_peer_export", buf, hb)
                lib.call("hpvg_peer_export", sig, hs)
            mine = (socket.gethostname(), torch.device(device).index, hb.raw, hs.raw)
        except Exception as e:      # noqa: BLE001 — reported below, on every rank alike
            err = "rank %d: %s" % (rank, e)
        everyone = [None] * world
        dist.all_gather_object(everyone, (err, mine), group=group)
        errs = [e for e, _ in everyone if e]
        bufs, sigs = [None] * world, [None] * world
        if not errs:
            try:
                me = torch.device(device).index
                for q, (_, (host, dev_q, hb_q, hs_q)) in enumerate(everyone):
                    if q == rank:
                        bufs[q], sigs[q] = self._local[0], self._local[1]
                        continue
                    if host != mine[0]:
                        raise RuntimeError("rank %d runs on another host (%s)" % (q, host))
                    if not lib.load().hpvg_peer_can_access(me, dev_q):
                        raise RuntimeError("GPU %d cannot access GPU %d's memory" % (me, dev_q))
                    with torch.cuda.device(device):
                        for handle, out in ((hb_q, bufs), (hs_q, sigs)):
                            p = ctypes.c_void_p()
                            lib.call("hpvg_peer_import", ctypes.create_string_buffer(handle, lib.PEER_HANDLE_BYTES), ctypes.byref(p))
                            self._mapped.append(p.value)
                            out[q] = p.value
            except Exception as e:      # noqa: BLE001
                errs = ["rank %d: %s" % (rank, e)]
        second = [None] * world
        dist.all_gather_object(second, errs[0] if errs else None, group=group)
        errs = [e for e in second if e]
        if errs:
            self.close()
            if rank == 0:
                sys.stderr.write("hpvg: gradient buckets stay on NCCL (%s)\n" % errs[0])
            return None
        self._bufs = (ctypes.c_void_p * world)(*bufs)
        self._sigs = (ctypes.c_void_p * world)(*sigs)
        self.flat = torch.as_tensor(_Raw(self._local[0], self.numel), device=device)
        # every pad is zero-filled (hpvg_peer_alloc synchronises) before any rank's first flag can arrive
        torch.cuda.synchronize(device)
        dist.barrier(group=group)
        return self

    def allreduce(self):
        """flat <- mean over ranks of flat, on the current stream (every rank must call it)"""
        lib.call("hpvg_peer_allreduce_avg", self._bufs, self._sigs, self.rank, self.world, self.numel,
                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))

    @staticmethod
    def numel_for(grads, world):
        """floats a bucket needs for these gradients in the slot layout of hpvg_peer_allreduce_avg_tensors (>= their total size)"""
        return int(lib.load().hpvg_peer_bucket_numel(len(grads), lib.longlong_array([g.numel() for g in grads]), world))

    @staticmethod
    def can_gather(grads):
        """the kernel reads and writes the gradient tensors themselves: fp32, contiguous, 16-byte aligned, at most 64 per call"""
        return (os.environ.get("HPVG_PEER_FUSED_PACK", "1") != "0" and 0 < len(grads) <= lib.PEER_MAX_TENSORS and
                all(g.dtype == torch.float32 and g.is_contiguous() and g.numel() > 0 and g.data_ptr() % 16 == 0 for g in grads))

    def allreduce_tensors(self, grads):
        """grads[i] <- mean over ranks of grads[i], in place, ONE launch on the current stream: the kernel fills the bucket from the
        tensors, exchanges, and writes the means back (every rank must call it with the same shapes)"""
        lib.call("hpvg_peer_allreduce_avg_tensors", self._bufs, self._sigs, self.rank, self.world, self.numel, len(grads),
                 lib.ptr_array(grads), lib.longlong_array([g.numel() for g in grads]),
                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))

    def close(self):
        handle = lib.load()
        for p in self._mapped:
            handle.hpvg_peer_close(ctypes.c_void_p(p))
        self.flat = None
        for p in self._local:
            handle.hpvg_peer_free(ctypes.c_void_p(p))
        self._mapped, self._local = [], []
