"""Gradient-bucket averaging: hpvg_peer_allreduce_avg (csrc/peer.cu) against one NCCL all-reduce of the same flat bucket.
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 experiments/peer_bench.py
Each measurement: 50 dependent calls between two CUDA events (max over ranks), after 10 warm-up calls."""
import os
import sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "hp-vae-gan_b200"), ROOT]
import torch
import torch.distributed as dist
from hpvg import peer

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def timed(fn, calls=50, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(calls):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item() * 1e3 / calls


for floats in ((16 * 1024, 331_000, 690_000, 1_380_000, 4_000_000) if world <= 2 else (331_000, 690_000, 1_380_000)):
    b = peer.PeerBucket.create(floats, dev)
    if b is None:
        if rank == 0:
            print("peer memory not available")
        break
    flat = torch.randn(b.numel, device=dev)
    b.flat.copy_(flat)
    us_peer = timed(b.allreduce)
    us_nccl = timed(lambda: dist.all_reduce(flat, op=dist.ReduceOp.AVG))
    # in a recorded graph (what the iteration replays): 20 calls per replay
    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g1):
        for _ in range(20):
            b.allreduce()
    with torch.cuda.graph(g2):
        for _ in range(20):
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    us_peer_g = timed(g1.replay, calls=10, warm=3) / 20
    us_nccl_g = timed(g2.replay, calls=10, warm=3) / 20
    if rank == 0:
        print("%d ranks, %8.2f MB bucket: peer kernel %6.1f us (graph %6.1f)   NCCL all-reduce %6.1f us (graph %6.1f)" %
              (world, b.numel * 4 / 1e6, us_peer, us_peer_g, us_nccl, us_nccl_g), flush=True)
    del g1, g2
# the generator's gradient list at BASELINE configs[1] (two trained stages, 36 tensors, 2.7 MB): ONE launch that gathers, exchanges and
# scatters (hpvg_peer_allreduce_avg_tensors) against multi-tensor pack + NCCL all-reduce + multi-tensor unpack
stage = [(64, 3, 3, 3, 3), (64,), (64,), (64,)] + 3 * [(64, 64, 3, 3, 3), (64,), (64,), (64,)] + [(3, 64, 3, 3, 3), (3,)]
grads = [torch.randn(s, device=dev) for s in stage + stage]
b = peer.PeerBucket.create(peer.PeerBucket.numel_for(grads, world), dev)
if b is not None:
    n = sum(g.numel() for g in grads)
    flat = torch.empty(n, device=dev)
    views, off = [], 0
    for g in grads:
        views.append(flat[off:off + g.numel()].view_as(g))
        off += g.numel()

    def nccl_path():
        torch._foreach_copy_(views, grads)
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
        torch._foreach_copy_(grads, views)

    def fused_path():
        b.allreduce_tensors(grads)

    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
    fused_path(); nccl_path()
    torch.cuda.synchronize()
    with torch.cuda.graph(g1):
        for _ in range(20):
            fused_path()
    with torch.cuda.graph(g2):
        for _ in range(20):
            nccl_path()
    us_f, us_n = timed(g1.replay, calls=10, warm=3) / 20, timed(g2.replay, calls=10, warm=3) / 20
    if rank == 0:
        print("%d ranks, %d gradient tensors (%.2f MB), inside a graph: gather + exchange + scatter in one launch %6.1f us   "
              "pack + NCCL all-reduce + unpack %6.1f us" % (world, len(grads), n * 4 / 1e6, us_f, us_n), flush=True)
    del g1, g2
torch.cuda.synchronize()
dist.barrier()
sys.stdout.flush()
os._exit(0)
