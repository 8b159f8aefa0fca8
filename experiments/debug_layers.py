"""layer-by-layer comparison of the CUDA decoder with the bf16-emulating oracle (development aid)"""
import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from helpers import opt_from, state_from, rel_err
from hpvg import ops
from modules import networks_3d
from oracle import port
name = sys.argv[1] if len(sys.argv) > 1 else "hp3d_tiny"
fx = torch.load(os.path.join(ROOT, "tests/golden", name + ".pt"), weights_only=False)
opt = opt_from(fx)
g = networks_3d.GeneratorHPVAEGAN(opt)
for _ in range(fx['stages']): g.init_next_stage()
g.load_state_dict(state_from(fx)); g.cuda()
sd = state_from(fx)
z = fx['rand']['z']
with torch.no_grad(), port.storage('bf16'):
    h_ref = port.store(z)
    h = ops.ToWide.apply(z.cuda())
    print("z", rel_err(ops.convert_raw(h, False), h_ref))
    for nm, m in g.decoder.named_children():
        if nm == 'tail':
            break
        # conv only
        stats = torch.zeros(2 * m.conv.weight.shape[0], device='cuda')
        y = ops.conv_raw(h, m.conv.weight, m.conv.bias, 1, False, True, stats=stats)
        y_ref = port.store(port.conv(h_ref, port.mma_weight(sd['decoder.%s.conv.weight' % nm]), sd['decoder.%s.conv.bias' % nm], 1))
        yt = ops.convert_raw(y, False).cpu()
        nflip = (yt != y_ref).float().mean().item()
        cnt = y_ref.numel() // y_ref.shape[1]
        mean_ref = y_ref.mean((0, 2, 3, 4)); var_ref = y_ref.var((0, 2, 3, 4), unbiased=False)
        mean = stats[:len(mean_ref)].cpu() / cnt; var = stats[len(mean_ref):].cpu() / cnt - mean * mean
        print(nm, "conv rel", rel_err(yt, y_ref), "flip frac", nflip, "mean rel", rel_err(mean, mean_ref), "var rel", rel_err(var, var_ref))
        h = m.run(h)
        h_ref = port.conv_block(sd, 'decoder.%s.' % nm, h_ref, 1)
        ht = ops.convert_raw(h, False).cpu()
        print(nm, "block rel", rel_err(ht, h_ref), "flip frac", (ht != h_ref).float().mean().item())
