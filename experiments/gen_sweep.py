"""generation leg (BASELINE configs[3]) over draws-per-forward x streams (development aid)"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "hp-vae-gan_b200"), ROOT]
import torch
import bench
from hpvg import train


class A:
    no_graph = False; warmup = 3; graph_candidates = 1; settle_steps = 0; steps = 5


dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
leg = bench.TrainLeg("cfg2", A, 0, 1, dev, False)
G, o = leg.G, leg.o
while len(o.Noise_Amps) < o.stop_scale + 1:
    o.Noise_Amps.append(0.07)      # the finest level's amplitude is normally computed at iteration 0
for batch, streams in [(32, 2), (32, 3), (64, 2), (64, 1), (16, 4), (128, 1)]:
    s = train.Sampler(G, o, dev, batch=batch, graph=True, streams=streams, static_weights=True)
    calls = max(1, 2048 // batch)

    def run():
        s.begin()
        for _ in range(calls):
            s.sample()
        s.wait()
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("batch %3d x %d streams: %.0f frames/s" % (batch, streams, calls * batch * 16 / (ms * 1e-3)), flush=True)
    del s
    torch.cuda.empty_cache()
