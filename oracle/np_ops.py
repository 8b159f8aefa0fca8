"""numpy restatement of the primitive operators the path resolves to inside PyTorch (TEST INFRASTRUCTURE — see
oracle/__init__.py).  The reference's arithmetic lives in its third-party dependency torch (pinned pytorch==1.4.0 in
env.sh:3; installed here: 2.11.0); these functions restate the published definitions of those operators in plain
float64 numpy so that oracle/port.py (which calls torch.nn.functional on the CPU) is itself checked against an
independent implementation (tests/test_oracle.py::test_numpy_primitives_match_torch).  Small sizes only.
"""
import numpy as np


def conv_nd(x, w, b=None, pad=0):
    """nn.Conv3d / nn.Conv2d forward, stride 1, zero padding `pad` (cross-correlation, as in torch): x [N,Cin,*sp],
    w [Cout,Cin,*k]  (modules/networks_3d.py:51, networks_2d.py:56)"""
    nd = x.ndim - 2
    xp = np.pad(x, [(0, 0), (0, 0)] + [(pad, pad)] * nd)
    k = w.shape[2:]
    out_sp = tuple(xp.shape[2 + i] - k[i] + 1 for i in range(nd))
    y = np.zeros((x.shape[0], w.shape[0]) + out_sp, dtype=np.float64)
    for idx in np.ndindex(*k):
        sl = tuple(slice(idx[i], idx[i] + out_sp[i]) for i in range(nd))
        patch = xp[(slice(None), slice(None)) + sl]                      # [N,Cin,*out]
        y += np.einsum('nc...,oc->no...', patch, w[(slice(None), slice(None)) + idx])
    if b is not None:
        y += b.reshape((1, -1) + (1,) * nd)
    return y


def batch_norm_train(y, gamma, beta, eps=1e-5):
    """nn.BatchNorm3d/2d in training mode: per-channel mean and BIASED variance over N and the spatial axes
    (modules/networks_3d.py:54); also returns (mean, unbiased variance) which feed the running statistics"""
    axes = (0,) + tuple(range(2, y.ndim))
    mean = y.mean(axes)
    var = y.var(axes)
    shape = (1, -1) + (1,) * (y.ndim - 2)
    out = (y - mean.reshape(shape)) / np.sqrt(var.reshape(shape) + eps) * gamma.reshape(shape) + beta.reshape(shape)
    count = y.size // y.shape[1]
    return out, mean, var * count / max(count - 1, 1)


def leaky_relu(x, slope=0.2):
    """nn.LeakyReLU(0.2) (modules/networks_3d.py:21)"""
    return np.where(x > 0, x, slope * x)


def resize_linear(x, size):
    """F.interpolate(mode='trilinear'|'bilinear', align_corners=True) (utils/images.py:13-24): separable linear
    interpolation where output index o maps to input coordinate o * (in - 1) / (out - 1)"""
    y = x.astype(np.float64)
    for axis, out_n in zip(range(2, x.ndim), size):
        in_n = y.shape[axis]
        pos = np.zeros(out_n) if out_n == 1 else np.arange(out_n) * (in_n - 1) / (out_n - 1)
        lo = np.minimum(np.floor(pos).astype(int), in_n - 1)
        hi = np.minimum(lo + 1, in_n - 1)
        frac = (pos - lo).reshape([-1 if a == axis else 1 for a in range(y.ndim)])
        y = np.take(y, lo, axis=axis) * (1 - frac) + np.take(y, hi, axis=axis) * frac
    return y


def spectral_norm_step(w, u, eps=1e-12):
    """one power iteration of nn.utils.spectral_norm (modules/networks_3d.py:63): v = normalize(W^T u),
    u = normalize(W v), sigma = u^T W v; returns (W / sigma, u, v)"""
    wm = w.reshape(w.shape[0], -1)
    v = wm.T @ u
    v = v / max(np.linalg.norm(v), eps)
    u = wm @ v
    u = u / max(np.linalg.norm(u), eps)
    sigma = u @ wm @ v
    return w / sigma, u, v


def kl(mu, logvar):
    """modules/losses.py:7-9"""
    return np.mean(-0.5 * (1 + logvar - mu ** 2 - np.exp(logvar)))


def gp_penalty(grads, lam):
    """modules/utils.py:18: L2 norm over the CHANNEL axis per voxel"""
    return np.mean((np.sqrt((grads ** 2).sum(1)) - 1) ** 2) * lam


def clip_grad_norm(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_(G_curr.parameters(), opt.grad_clip) (train_video.py:201, train_image.py:216):
    total = 2-norm over ALL gradients, every gradient scaled by min(1, max_norm / (total + 1e-6)); returns (scaled, total)"""
    total = np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads))
    coef = min(1.0, max_norm / (total + 1e-6))
    return [g * coef for g in grads], total


def adam_step(p, g, m, v, step, lr, beta1, beta2=0.999, eps=1e-8):
    """optim.Adam(lr, betas=(opt.beta1, 0.999)).step() (train_video.py:55,88,183,202; eps 1e-8, no weight decay, no amsgrad):
    `step` is the count BEFORE this update; returns (p, m, v) after it"""
    t = step + 1
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    denom = np.sqrt(v) / np.sqrt(1 - beta2 ** t) + eps
    return p - (lr / (1 - beta1 ** t)) * m / denom, m, v
