"""Host-side logic of the multi-GPU modes on CPU with the gloo backend, world size 2 (the kernels themselves need a GPU:
the N > 1 GPU path is exercised by bench.py --gpus N)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hpvg import train
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(3, 4)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
        for i, p in enumerate(params[:2]):            # the third parameter has no gradient on any rank (frozen stage)
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        bucket = train.GradBucket()
        nbytes = bucket.average(params)
        ok = nbytes == (12 + 7) * 4
        ok = ok and torch.allclose(params[0].grad, torch.full((3, 4), 1.5)) and torch.allclose(params[1].grad, torch.full((7,), 3.0))
        ok = ok and params[2].grad is None
        # a second call reuses the flat buffer
        for p in params[:2]:
            p.grad.fill_(float(rank))
        bucket.average(params)
        ok = ok and torch.allclose(params[0].grad, torch.full((3, 4), 0.5))
        # the noise-amplitude reduction of iteration 0 (train_video.py:143-144 over ranks)
        mse = torch.tensor(float(rank + 1))
        dist.all_reduce(mse)
        ok = ok and abs(mse.item() / world - 1.5) < 1e-6
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_allreduce_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}


def test_draws_partition_over_ranks():
    from hpvg import train
    for total in (4096, 10, 7, 1):
        for world in (1, 2, 4, 8):
            spans = [train.draws_for_rank(total, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == total
            pos = 0
            for start, count in spans:
                assert start == pos and count >= 0
                pos += count
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1


def _peer_refusal_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hpvg import peer, train
        # no CUDA device here: the allocation fails on every rank, and every rank must learn that the bucket stays on the library collective
        bucket = peer.PeerBucket.create(1000, "cuda:0")
        # ... and a GradBucket that was refused once does not try again (no set-up collective per backward)
        gb = train.GradBucket()
        gb.peer_refused = bucket is None
        out[rank] = (bucket is None, gb._peer_bucket(1000, "cuda:0") is None)
    finally:
        dist.destroy_process_group()


def test_peer_bucket_refusal_is_collective():
    """hpvg.peer.PeerBucket.create: when a rank cannot set its bucket up, ALL ranks return None from the same call (two all-gathers of
    the error state), so that no rank launches the peer kernel while another one calls NCCL"""
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_refusal_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: (True, True), 1: (True, True)}
