"""numpy model of wgrad_tc_kdstack_kernel's index arithmetic (hp-vae-gan_b200/csrc/wgrad_tc.cu): which slab rows, gy slices,
TMEM lanes / columns and partial slots the kernel combines — checked against the definition
    dw[co][ci][kd,kh,kw] = sum gy[n,od,oh,ow,co] * x[n,od+kd-p,oh+kh-p,ow+kw-p,ci].
It validates the mapping (work items, out-of-range slices, tap pairing, drain), not the hardware descriptor semantics.
Runs on the CPU:  python experiments/model_wgrad_stack.py
"""
import numpy as np

BH, BW = 16, 8
SLAB_W, SLAB_H = BW + 2, BH + 2


def tma_box(t, n, d, h0, w0, hh, ww):
    """[hh][ww][C] box of t[n, d] starting at (h0, w0); everything outside the tensor reads as zero (TMA zero fill)"""
    N, D, H, W, C = t.shape
    out = np.zeros((hh, ww, C))
    if d < 0 or d >= D:
        return out
    for i in range(hh):
        for j in range(ww):
            h, w = h0 + i, w0 + j
            if 0 <= h < H and 0 <= w < W:
                out[i, j] = t[n, d, h, w]
    return out


def kernel_model(x, gy, pad, splits=3):
    N, Di, Hi, Wi, Cin = x.shape
    _, Do, Ho, Wo, Cout = gy.shape
    pad_d = pad
    bricks_h, bricks_w = -(-Ho // BH), -(-Wo // BW)
    num_items = N * Di * bricks_h * bricks_w
    per_split = -(-num_items // splits)
    taps = 27
    partial = np.zeros((splits, taps, Cin, Cout))
    for split in range(splits):
        for kh in range(3):                                   # blockIdx.z
            tmem = np.zeros((2, 128, 192))                    # [accumulator][lane][column]
            for b in range(split * per_split, min(num_items, (split + 1) * per_split)):
                w0 = (b % bricks_w) * BW
                r = b // bricks_w
                h0 = (r % bricks_h) * BH
                r //= bricks_h
                d, n = r % Di, r // Di
                slab = tma_box(x, n, d, h0 - pad, w0 - pad, SLAB_H, SLAB_W).reshape(SLAB_H * SLAB_W, Cin)   # rows of 128 B
                atoms = [tma_box(gy, n, d + pad_d - kd, h0, w0, BH, BW).reshape(BH * BW, Cout) for kd in range(3)]
                bmat = np.concatenate(atoms, axis=1)          # [voxel][kd*64 + co]: N-atoms LBO apart
                for pr in range(2):
                    start = kh * SLAB_W + 2 * pr              # descriptor start row; second M-atom `lbo` rows further
                    lbo_rows = 1 if pr == 0 else 0
                    for v in range(BH * BW):                  # K index = brick voxel: row v // 8, column v % 8
                        srow = (v // 8) * SLAB_W + (v % 8)
                        a = np.concatenate([slab[start + srow], slab[start + lbo_rows + srow]])   # lanes 0-63, 64-127
                        tmem[pr] += np.outer(a, bmat[v])
            for pr in range(2):                               # drain
                for m in range(128):
                    kw = 2 * pr + (m >> 6)
                    if kw > 2:
                        continue
                    for kd in range(3):
                        partial[split, kd * 9 + kh * 3 + kw, m & 63] = tmem[pr, m, kd * 64:kd * 64 + Cout]
    dw = partial.sum(0)                                       # [tap][ci][co]
    return dw.transpose(2, 1, 0).reshape(Cout, Cin, 3, 3, 3)


def definition(x, gy, pad):
    N, Di, Hi, Wi, Cin = x.shape
    _, Do, Ho, Wo, Cout = gy.shape
    xp = np.pad(x, [(0, 0), (pad, pad), (pad, pad), (pad, pad), (0, 0)])
    dw = np.zeros((Cout, Cin, 3, 3, 3))
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                patch = xp[:, kd:kd + Do, kh:kh + Ho, kw:kw + Wo]
                dw[:, :, kd, kh, kw] = np.einsum('ndhwo,ndhwi->oi', gy, patch)
    return dw


def main():
    rng = np.random.default_rng(0)
    C = 64          # the model keeps the kernel's 64-lane / 64-column atoms; small volumes keep it fast
    for (n, d, h, w), pad in [((1, 3, 5, 9), 1), ((2, 2, 17, 6), 1), ((1, 5, 20, 11), 0)]:
        x = rng.standard_normal((n, d, h, w, C))
        gy = rng.standard_normal((n, d + 2 * pad - 2, h + 2 * pad - 2, w + 2 * pad - 2, C))
        got, want = kernel_model(x, gy, pad), definition(x, gy, pad)
        err = np.abs(got - want).max() / np.abs(want).max()
        print((n, d, h, w), "pad", pad, "max rel err %.2e" % err, "ok" if err < 1e-12 else "MISMATCH")
        assert err < 1e-12


if __name__ == "__main__":
    main()
