"""BASELINE configs[0] (train_image.py, 2-D HP-VAE-GAN, 128 px, vae-levels 3) finest-level iteration on the GPU (sanity + timing)"""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200")); sys.path.insert(0, ROOT)
import torch
from hpvg import train
from hpvg.options import Options
from modules import networks_2d
dev = torch.device("cuda", 0)
o = Options(img_size=128, vae_levels=3, nfc=64, latent_dim=128, num_layer=5, batch_size=1)
o.scale_idx = o.stop_scale
o.Noise_Amps = [1.0] + [0.07] * (o.stop_scale - 1)
_, h0, w0 = o.level_size(0)
o.Z_init_size = [1, o.latent_dim, h0, w0]
torch.manual_seed(0)
G = networks_2d.GeneratorHPVAEGAN(o)
for _ in range(o.scale_idx): G.init_next_stage()
D = networks_2d.WDiscriminator2D(o)
G.to(dev); D.to(dev)
_, h, w = o.level_size(o.scale_idx)
real = torch.rand(1, 3, h, w, device=dev) * 2 - 1
real_zero = torch.rand(1, 3, h0, w0, device=dev) * 2 - 1
tr = train.ScaleTrainer(o, G, D, capturable=True, dims=2)
out = tr.capture(real, real_zero, warmup=3)
for _ in range(3): tr.replay()
torch.cuda.synchronize()
t0 = time.perf_counter()
n = 50
for _ in range(n): tr.replay()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / n
print("2-D finest level %dx%d: %.3f ms/iteration  %.1f iter/s  losses %s" % (h, w, dt * 1e3, 1 / dt, {k: round(v.item(), 4) for k, v in out.items()}))
