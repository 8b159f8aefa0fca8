"""diagnosis: bench.parity_leg's pieces one by one (eager vs oracle with default-init weights, then replay variants)"""
import os, sys, copy, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "hp-vae-gan_b200"), ROOT]
import torch
import bench
from hpvg import train, ops
from modules import networks_3d
from oracle import train_ref

bench.WORKLOAD["name"] = "cfg2"
dev = torch.device("cuda", 0)
o0 = bench.make_opt()
sg, sd = bench.fresh_states(o0)
real, real_zero = bench.synthetic_clip(o0, 0)
gen = torch.Generator().manual_seed(123)
z = tuple(o0.Z_init_size)
levels = list(range(o0.vae_levels, o0.stop_scale + 1))
draws = []
for it in range(3):
    dr = {"noise_init": torch.randn(z, generator=gen)}
    if it == 0:
        dr["eps_amp"] = torch.randn(z, generator=gen)
    dr["eps"] = torch.randn(z, generator=gen)
    dr["noises"] = {l: torch.randn(bench.level_shape(o0, l), generator=gen) for l in levels}
    dr["alpha"] = 0.25 + 0.25 * it
    draws.append(dr)
flat = lambda dr: [dr["noise_init"]] + ([dr["eps_amp"]] if "eps_amp" in dr else []) + [dr["eps"]] + [dr["noises"][l] for l in levels]

oc = bench.make_opt()
oracle = train_ref.ScaleTrainer(oc, {k: v.clone().float() for k, v in sg.items()}, {k: v.clone().float() for k, v in sd.items()})
ref = []
for dr in draws:
    out = oracle.iteration(real, real_zero, noise_init=dr["noise_init"], eps=dr["eps"], noises=dr["noises"], alpha=dr["alpha"], eps_amp=dr.get("eps_amp"))
    ref.append({k: v.item() for k, v in out.items()})


def nets():
    o = bench.make_opt()
    G = networks_3d.GeneratorHPVAEGAN(o)
    for _ in range(o.scale_idx):
        G.init_next_stage()
    G.load_state_dict(sg); D = networks_3d.WDiscriminator3D(o); D.load_state_dict(sd)
    return o, G.to(dev), D.to(dev)


def show(tag, got, r):
    print(tag, " ".join("%s %.5f/%.5f(%+.2f%%)" % (k, got[k], r[k], 100 * (got[k] - r[k]) / abs(r[k])) for k in ("rec_loss", "gradient_penalty", "errD_real", "errD_fake", "errG")), flush=True)


rd, rzd = real.to(dev), real_zero.to(dev)
for knobs in ({}, {"overlap": False}):
    o, G, D = nets()
    tr = train.ScaleTrainer(o, G, D, **knobs)
    feed = train.NoiseFeed(dev)
    with feed:
        for it, dr in enumerate(draws):
            feed.load(flat(dr), dr["alpha"])
            got = {k: v.item() for k, v in tr.iteration(rd, rzd).items()}
            show("eager %s it%d" % (knobs, it), got, ref[it])

# replay without restore: it0 eager, it1 = capture warm-up, it2 replayed
o, G, D = nets()
tr = train.ScaleTrainer(o, G, D, capturable=True)
feed = train.NoiseFeed(dev)
with feed:
    feed.load(flat(draws[0]), draws[0]["alpha"]); tr.iteration(rd, rzd)
    feed.load(flat(draws[1]), draws[1]["alpha"])
    stock = tr.iteration
    def rew(a, b):
        feed.rewind(); return stock(a, b)
    tr.iteration = rew
    tr.capture(rd, rzd, warmup=1)
    tr.iteration = stock
    feed.load(flat(draws[2]), draws[2]["alpha"])
    got = {k: v.item() for k, v in tr.replay().items()}
    show("replay it2 (no restore)", got, ref[2])
