"""does running several batch-1 sampler graphs on separate streams raise generated frames/s? (development aid)"""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200")); sys.path.insert(0, ROOT)
import torch
import bench
from hpvg import train
from modules import networks_3d
o = bench.make_opt()
o.Noise_Amps = [1.0] + [0.07] * o.stop_scale
sg, sd = bench.fresh_states(o)
dev = torch.device("cuda", 0)
G = networks_3d.GeneratorHPVAEGAN(o)
for _ in range(o.scale_idx): G.init_next_stage()
G.load_state_dict(sg); G.to(dev)
for S in (1, 2, 3, 4, 6):
    samplers, streams = [], []
    for i in range(S):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            samplers.append(train.Sampler(G, o, dev, batch=1, graph=True))
        streams.append(st)
    torch.cuda.synchronize()
    draws = 240
    def run():
        for i in range(draws):
            k = i % S
            with torch.cuda.stream(streams[k]):
                samplers[k].sample()
        torch.cuda.synchronize()
    run()
    t0 = time.perf_counter(); run(); dt = time.perf_counter() - t0
    print("streams %d: %.3f ms/draw  %.0f frames/s" % (S, dt / draws * 1e3, draws * 16 / dt), flush=True)
    del samplers
