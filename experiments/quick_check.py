"""Ad-hoc GPU check of the raw kernels against torch functional ops (development aid; the real parity tests are in tests/)."""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
import torch.nn.functional as F
from hpvg import ops, lib

torch.manual_seed(0)
dev = "cuda"

def wide(t):   # NCDHW fp32 -> wide bf16
    return t.permute(0, 2, 3, 4, 1).contiguous().to(torch.bfloat16)
def thin(t):   # wide -> NCDHW fp32
    return t.float().permute(0, 4, 1, 2, 3).contiguous()
def rel(a, b):
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

def check_conv(n, cin, cout, d, h, w, pad, x_wide, y_wide, act=None, backend=lib.BACKEND_AUTO, transposed=False, tag=""):
    lib.set_conv_backend(backend)
    x = torch.randn(n, cin, d, h, w, device=dev)
    if transposed:
        wt = torch.randn(cin, cout, 3, 3, 3, device=dev) * 0.05   # forward weight [Cout_f=cin, Cin_f=cout]
    else:
        wt = torch.randn(cout, cin, 3, 3, 3, device=dev) * 0.05
    b = torch.randn(cout, device=dev) if not transposed else None
    xq = wide(x) if x_wide else x
    xr = thin(xq) if x_wide else x
    tc = x_wide and backend != lib.BACKEND_DIRECT and ((y_wide and cin % 64 == 0 and cout % 64 == 0) or (not y_wide and cin == 64 and cout <= 16 and act is None))
    wr = wt.to(torch.bfloat16).float() if tc else wt
    stats = torch.zeros(2 * cout, device=dev) if y_wide else None
    y = ops.conv_raw(xq, wt, b, pad, transposed, y_wide, act_slope=act, stats=stats)
    torch.cuda.synchronize()
    if transposed:
        ref = F.conv_transpose3d(xr, wr, None, padding=2 - pad)   # dgrad of forward conv with padding p=2-pad ... pad here is the dgrad-call pad
    else:
        ref = F.conv3d(xr, wr, b, padding=pad)
    if act is not None:
        ref = F.leaky_relu(ref, act)
    yt = thin(y) if y_wide else y
    e = rel(yt, ref)
    msg = "conv%s n%d %d->%d %dx%dx%d pad%d xw%d yw%d be%d T%d: rel %.3e" % (tag, n, cin, cout, d, h, w, pad, x_wide, y_wide, backend, transposed, e)
    if stats is not None:
        s_ref = torch.stack([yt.sum((0, 2, 3, 4)), (yt * yt).sum((0, 2, 3, 4))]).flatten()
        msg += " stats rel %.3e" % rel(stats, s_ref)
    print(msg, flush=True)
    return e

def check_wgrad(n, cin, cout, d, h, w, pad, x_wide, g_wide, backend=lib.BACKEND_AUTO):
    lib.set_conv_backend(backend)
    x = torch.randn(n, cin, d, h, w, device=dev)
    do, ho, wo = d + 2 * pad - 2, h + 2 * pad - 2, w + 2 * pad - 2
    gy = torch.randn(n, cout, do, ho, wo, device=dev)
    xq = wide(x) if x_wide else x
    gq = wide(gy) if g_wide else gy
    xr = thin(xq) if x_wide else x
    gr = thin(gq) if g_wide else gy
    dw, db = ops.wgrad_raw(xq, gq, pad, (cout, cin, 3, 3, 3), want_bias=True)
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv3d_weight(xr, (cout, cin, 3, 3, 3), gr, padding=pad)
    print("wgrad n%d %d->%d %dx%dx%d pad%d xw%d gw%d be%d: rel %.3e  bias rel %.3e" % (
        n, cin, cout, d, h, w, pad, x_wide, g_wide, backend, rel(dw, ref), rel(db, gr.sum((0, 2, 3, 4)))), flush=True)

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "direct"):
    check_conv(1, 3, 64, 4, 9, 11, 1, False, True, act=0.2)
    check_conv(2, 64, 3, 3, 10, 7, 1, True, False)
    check_conv(1, 64, 1, 3, 10, 7, 1, True, False)
    check_conv(1, 16, 24, 3, 6, 7, 1, True, True, backend=lib.BACKEND_DIRECT)
    check_conv(1, 64, 64, 3, 6, 7, 1, True, True, backend=lib.BACKEND_DIRECT)
    check_conv(1, 3, 8, 5, 9, 9, 0, False, False)
    check_conv(1, 8, 3, 5, 9, 9, 2, False, False)
    check_conv(1, 1, 64, 4, 9, 11, 1, False, True, transposed=True)     # tail dgrad: forward weight [1,64,...]
    check_conv(1, 64, 3, 4, 9, 11, 1, True, False, transposed=True)     # head dgrad
    check_wgrad(1, 3, 64, 4, 9, 11, 1, False, True)
    check_wgrad(1, 64, 1, 4, 9, 11, 1, True, False)
    check_wgrad(2, 16, 8, 3, 6, 7, 1, True, True, backend=lib.BACKEND_DIRECT)
    check_wgrad(1, 8, 8, 5, 8, 7, 0, False, False)
if which in ("all", "tc"):
    check_conv(1, 64, 64, 4, 16, 8, 1, True, True, tag="[tc]")
    check_conv(1, 64, 64, 4, 32, 32, 1, True, True, act=0.2, tag="[tc]")
    check_conv(1, 64, 64, 6, 54, 54, 1, True, True, tag="[tc]")
    check_conv(2, 64, 128, 4, 32, 32, 1, True, True, tag="[tc]")
    check_conv(1, 128, 64, 4, 32, 32, 1, True, True, tag="[tc]")
    check_conv(1, 64, 64, 16, 64, 64, 1, True, True, tag="[tc]")
    check_conv(1, 64, 64, 7, 20, 21, 0, True, True, tag="[tc]")
    check_conv(1, 64, 64, 5, 20, 21, 2, True, True, tag="[tc]")
    check_conv(1, 64, 64, 4, 16, 16, 1, True, True, transposed=True, tag="[tc]")
if which in ("all", "narrow"):
    check_conv(1, 3, 64, 4, 9, 11, 1, False, True, act=0.2, tag="[expand]")
    check_conv(2, 3, 64, 5, 17, 40, 1, False, True, tag="[expand]")
    check_conv(1, 1, 64, 4, 9, 11, 1, False, True, transposed=True, tag="[expand]")
    check_conv(1, 3, 64, 7, 20, 21, 0, False, True, tag="[expand]")
    check_conv(1, 3, 64, 16, 64, 64, 1, False, True, tag="[expand]")
    check_wgrad(1, 3, 64, 4, 9, 11, 1, False, True)
    check_wgrad(2, 3, 64, 5, 17, 40, 1, False, True)
    check_wgrad(1, 3, 64, 7, 20, 21, 0, False, True)
    check_wgrad(1, 3, 64, 16, 64, 64, 1, False, True)
    check_wgrad(1, 64, 1, 4, 9, 11, 1, True, False)
    check_wgrad(2, 64, 3, 5, 17, 40, 1, True, False)
    check_wgrad(1, 64, 3, 7, 20, 21, 0, True, False)
    check_wgrad(1, 64, 3, 16, 64, 64, 1, True, False)
if which in ("all", "thin"):
    check_conv(1, 64, 3, 4, 16, 8, 1, True, False, tag="[tc-thin]")
    check_conv(2, 64, 3, 3, 10, 7, 1, True, False, tag="[tc-thin]")
    check_conv(1, 64, 1, 6, 54, 54, 1, True, False, tag="[tc-thin]")
    check_conv(1, 64, 3, 16, 64, 64, 1, True, False, tag="[tc-thin]")
    check_conv(1, 64, 3, 7, 20, 21, 0, True, False, tag="[tc-thin]")
    check_conv(1, 64, 3, 4, 9, 11, 1, True, False, transposed=True, tag="[tc-thin]")
if which in ("all", "wtc"):
    check_wgrad(1, 64, 64, 4, 16, 8, 1, True, True)
    check_wgrad(1, 64, 64, 4, 32, 32, 1, True, True)
    check_wgrad(1, 64, 64, 6, 54, 54, 1, True, True)
    check_wgrad(1, 64, 128, 4, 32, 32, 1, True, True)
    check_wgrad(1, 128, 64, 4, 32, 32, 1, True, True)
    check_wgrad(1, 64, 64, 16, 64, 64, 1, True, True)
    check_wgrad(1, 64, 64, 7, 20, 21, 0, True, True)
print("launches", lib.launch_count())
