"""Multi-GPU modes on real devices: needs >= 2 visible GPUs (skipped otherwise; run with `gpurun --gpus 2`).

Batched-noise data-parallel training: one clip per rank, replicated weights, the gradients averaged over ranks once per backward
(one libhpvg kernel over NVLink peer memory, hpvg/peer.py; NCCL where peer memory cannot be mapped).  The critic has no BatchNorm, so the gradient of the two-clip batch is exactly the mean of the per-clip gradients: the
averaged gradients a rank holds after the distributed iteration must equal the mean of the gradients two single-GPU iterations
produce on the two clips, and both ranks must take the same optimizer steps."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "hp-vae-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from helpers import state_d_from, state_from, train_opt_from
        from hpvg import train
        from modules import networks_3d
        fx = torch.load(os.path.join(ROOT, "tests", "golden", "train_gan_tiny.pt"), map_location="cpu", weights_only=False)

        def build(distributed):
            opt = train_opt_from(fx)
            g = networks_3d.GeneratorHPVAEGAN(opt)
            for _ in range(fx['stages']):
                g.init_next_stage()
            g.load_state_dict(state_from(fx), strict=True)
            d = networks_3d.WDiscriminator3D(opt)
            d.load_state_dict(state_d_from(fx), strict=True)
            g.to(dev); d.to(dev)
            return opt, g, d, train.ScaleTrainer(opt, g, d, distributed=distributed)

        dr = fx['draws'][0]
        draws = [dr['noise_init'], dr['eps_amp'], dr['eps']] + [dr['noises'][l] for l in sorted(dr['noises'])]
        clips = [(fx['real'] * (1.0 - 0.3 * r) + 0.05 * r, fx['real_zero'] * (1.0 - 0.3 * r) + 0.05 * r) for r in range(world)]
        # single-GPU iterations on every clip (each rank computes all of them): per-clip critic gradients
        singles = []
        for r in range(world):
            opt_s, g_s, d_s, tr_s = build(False)
            opt_s.Noise_Amps = list(fx['amps_before']) + [0.1]        # fixed amplitude: the distributed run averages the MSE over ranks
            feed = train.NoiseFeed(dev)
            with feed:
                feed.load(draws, dr['alpha'])
                tr_s.iterations = 1                                    # skip the amplitude computation (and its eps_amp draw)
                feed.load([draws[0]] + draws[2:], dr['alpha'])
                tr_s.iteration(clips[r][0].to(dev), clips[r][1].to(dev))
            singles.append({k: p.grad.detach().clone() for k, p in d_s.named_parameters()})
        opt_d, g_d, d_d, tr_d = build(True)
        opt_d.Noise_Amps = list(fx['amps_before']) + [0.1]
        feed = train.NoiseFeed(dev)
        with feed:
            tr_d.iterations = 1
            feed.load([draws[0]] + draws[2:], dr['alpha'])
            tr_d.iteration(clips[rank][0].to(dev), clips[rank][1].to(dev))
        torch.cuda.synchronize()
        worst = 0.0
        for k, p in d_d.named_parameters():
            mean = sum(s[k] for s in singles) / world
            err = (p.grad - mean).norm().item() / (mean.norm().item() + 1e-12)
            worst = max(worst, err)
        # both ranks took the same steps: weights identical across ranks
        flat = torch.cat([p.detach().flatten() for p in list(g_d.parameters()) + list(d_d.parameters())])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        spread = max((t - gathered[0]).abs().max().item() for t in gathered)
        out[rank] = (worst, spread, tr_d.allreduce_bytes_per_iter, tr_d.bucketD.kind)
    finally:
        dist.destroy_process_group()


def test_distributed_critic_gradients_equal_the_mean_of_single_gpu_gradients():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = dict(out)
    print("worst relative error of the averaged critic gradients, weight spread across ranks, all-reduce bytes:", res)
    assert set(res) == {0, 1}
    for worst, spread, nbytes, kind in res.values():
        assert worst < 5e-3, res          # same kernels on the same inputs: atomics' summation order only
        assert spread == 0.0, res         # identical averaged gradients -> bit-identical steps on every rank
        assert nbytes > 0
        assert kind == "peer", res        # the buckets of a single-node NVLink box live in peer memory


def _peer_worker(rank, world, port, out):
    for p in (os.path.join(ROOT, "hp-vae-gan_b200"), ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from hpvg import train
        sizes = [(64, 64, 3, 3, 3), (64,), (3, 64, 3, 3, 3), (7,), (1, 5, 1)]      # ragged: 110 592 + 64 + 5 184 + 7 + 5 floats

        def fill(params, seed):
            g = torch.Generator(device=dev).manual_seed(1000 * seed + rank)
            for p in params:
                p.grad = torch.randn(p.shape, generator=g, device=dev) * (1.0 + rank)

        def expected(params):
            flat = torch.cat([p.grad.flatten() for p in params])
            both = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(both, flat)
            acc = both[0].clone()
            for t in both[1:]:
                acc += t          # rank order, as the kernel sums
            return acc * (1.0 / world)

        res = {}
        for mode in ("1", "0"):      # the kernel gathers / scatters the gradients itself (default); pack and unpack copies around the exchange
            os.environ["HPVG_PEER_FUSED_PACK"] = mode
            params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in sizes]
            bucket = train.GradBucket()
            worst = 0.0
            for it in range(6):      # eager calls: the flags count calls, nothing is reset in between
                fill(params, it)
                want = expected(params)
                nbytes = bucket.average(params)
                got = torch.cat([p.grad.flatten() for p in params])
                worst = max(worst, (got - want).abs().max().item())
            # the same call recorded into a CUDA graph and replayed on fresh gradients
            static = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in sizes]
            for p in static:
                p.grad = torch.zeros_like(p)
            graph_bucket = train.GradBucket()
            graph_bucket.average(static)          # set-up (allocation, handle exchange) happens outside the capture
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                graph_bucket.average(static)
            worst_replay = 0.0
            for it in range(5):
                fill(params, 100 + it)
                for p, q in zip(static, params):
                    p.grad.copy_(q.grad)
                want = expected(static)
                graph.replay()
                got = torch.cat([p.grad.flatten() for p in static])
                worst_replay = max(worst_replay, (got - want).abs().max().item())
            torch.cuda.synchronize()
            # every rank holds the same bits
            got = torch.cat([p.grad.flatten() for p in static])
            both = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(both, got)
            spread = max((t - both[0]).abs().max().item() for t in both)
            res[mode] = (bucket.kind, graph_bucket.kind, worst, worst_replay, spread, nbytes)
            del graph
        out[rank] = res
    finally:
        dist.destroy_process_group()


def test_peer_memory_gradient_bucket_equals_the_mean_over_ranks():
    """hpvg_peer_allreduce_avg_tensors / hpvg_peer_allreduce_avg (csrc/peer.cu) through train.GradBucket: ragged gradient lists, repeated eager calls and replays of a
    recorded call against the mean formed from an NCCL all-gather in the kernel's (rank) order — bit-exact, identical on every rank"""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    res = dict(out)
    print("bucket kinds, worst |difference| eager / replayed, spread across ranks, bytes:", res)
    assert set(res) == {0, 1}
    for per_mode in res.values():
        assert set(per_mode) == {"1", "0"}
        for kind, kind_graph, worst, worst_replay, spread, nbytes in per_mode.values():
            assert kind == "peer" and kind_graph == "peer", res
            assert worst == 0.0 and worst_replay == 0.0, res
            assert spread == 0.0, res
            assert nbytes == 4 * (110592 + 64 + 5184 + 7 + 5)
