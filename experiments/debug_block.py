import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "hp-vae-gan_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn.functional as F
from hpvg import ops
from modules import networks_3d
from oracle import port
cin, cout, shape = 64, 64, (1, 6, 20, 24)
m = networks_3d.ConvBlock3D(cin, cout, 3, 1, 1)
sd = m.state_dict(); port.det_fill(sd, 5); sd = {k: v.detach().clone() for k, v in sd.items()}
m.cuda()
n, d, h, w = shape
x = port.det_tensor((n, cin, d, h, w), 1).bfloat16().float()
xw = ops.ToWide.apply(x.cuda())
stats = torch.zeros(128, device='cuda')
y = ops.conv_raw(xw, m.conv.weight.detach(), m.conv.bias.detach(), 1, False, True, stats=stats)
yt = ops.convert_raw(y, False).cpu()
y_ref = F.conv3d(x, sd['conv.weight'].bfloat16().float(), sd['conv.bias'], padding=1)
err = (yt - y_ref).abs()
print("conv max abs err", err.max().item(), "rel", ((yt - y_ref).norm() / y_ref.norm()).item())
bad = (err > 0.02).nonzero()
print("bad count", len(bad), bad[:20].tolist())
cnt = y_ref.numel() // 64
print("stats mean rel", ((stats[:64].cpu() / cnt - yt.mean((0, 2, 3, 4))).norm() / yt.mean((0, 2, 3, 4)).norm()).item())
out = ops.ToThin.apply(m.run(xw)).cpu()
ref = port.conv_block({k: v.clone() for k, v in sd.items()}, '', x, 1)
e2 = (out - ref).abs()
print("block rel", ((out - ref).norm() / ref.norm()).item(), "max", e2.max().item())
bad = (e2 > 0.05).nonzero()
print("bad count", len(bad), bad[:20].tolist())
# per-channel error
pc = ((out - ref) ** 2).sum((0, 2, 3, 4)).sqrt() / (ref ** 2).sum((0, 2, 3, 4)).sqrt()
print("per-channel rel err", [round(v, 4) for v in pc.tolist()])
