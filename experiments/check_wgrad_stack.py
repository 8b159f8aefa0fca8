"""Parity + timing of the kd-stacked weight-gradient kernel (wgrad_tc.cu, hpvg_set_wgrad_mode(1)) against the measured
default (mode 0) and the CUDA-core kernel.  The stacked kernel was written after round 1's GPU budget was spent and has not run
yet: run this FIRST (under `timeout 60`: a barrier-protocol bug traps the launch, it does not hang), then promote the cases
into tests/test_gpu_layers.py and flip the default if it is faster.

    gpurun -- 'timeout 90 python experiments/check_wgrad_stack.py > gpurun_out/wgrad_stack.txt 2>&1'

What to look at if it fails:
  * launch failure / trap in the first case  -> fully out-of-bounds gy boxes (slices -1 and Do) did not deliver their bytes:
    skip those loads and zero the atom with st.shared instead, or clamp the N range as conv_tc's partial units do;
  * wrong values only in taps kd = 1, 2      -> the N-atom stride of an MN-major B operand is not the LBO field (try SBO/LBO swapped
    as experiments/umma_desc_probe.cu does for A);
  * wrong values only at kw = 2              -> the duplicated upper half (LBO = 0) of the second accumulator.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import lib, ops

CASES = [  # cin, cout, (n, d, h, w), pad
    (64, 64, (1, 16, 64, 64), 1),      # BASELINE config 2, finest level
    (64, 64, (1, 4, 32, 32), 1),       # coarsest level: every slab touches a volume edge in d
    (64, 64, (1, 6, 54, 54), 1),       # ragged bricks
    (64, 64, (1, 13, 64, 64), 1),      # default sampling rates
    (128, 64, (1, 4, 32, 32), 1),      # decoder input
    (64, 128, (1, 4, 32, 32), 1),      # mu / logvar heads
    (64, 64, (1, 22, 50, 50), 0),      # GeneratorSG: pad 0, Di = Do + 2
    (64, 64, (3, 5, 20, 24), 1),       # batch > 1
]


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def run(mode, backend, x, g, pad, wshape):
    lib.set_conv_backend(backend)
    prev = lib.set_wgrad_mode(mode)
    try:
        dw, _ = ops.wgrad_raw(x, g, pad, wshape)
        torch.cuda.synchronize()
    finally:
        lib.set_wgrad_mode(prev)
        lib.set_conv_backend(lib.BACKEND_AUTO)
    return dw


def timed(mode, x, g, pad, wshape, reps=20):
    flush = torch.empty(256 * 2**20 // 4, device="cuda")
    prev = lib.set_wgrad_mode(mode)
    try:
        ts = []
        for i in range(reps + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.wgrad_raw(x, g, pad, wshape)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1) * 1e3)
    finally:
        lib.set_wgrad_mode(prev)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ok = True
    for cin, cout, (n, d, h, w), pad in CASES:
        gen = torch.Generator(device="cuda").manual_seed(1)
        x = torch.randn((n, d, h, w, cin), device="cuda", generator=gen).bfloat16()
        do, ho, wo = d + 2 * pad - 2, h + 2 * pad - 2, w + 2 * pad - 2
        g = torch.randn((n, do, ho, wo, cout), device="cuda", generator=gen).bfloat16()
        wshape = (cout, cin, 3, 3, 3)
        ref = run(0, lib.BACKEND_DIRECT, x, g, pad, wshape)
        base = run(0, lib.BACKEND_TCGEN05, x, g, pad, wshape)
        new = run(1, lib.BACKEND_TCGEN05, x, g, pad, wshape)
        staged = run(2, lib.BACKEND_TCGEN05, x, g, pad, wshape)       # mode 2: the measured kernel with the staged drain
        if not torch.equal(staged, base):
            ok = False
            print("  mode 2 (staged drain) differs from mode 0: rel %.2e  MISMATCH (must be bit-identical)" % rel(staged, base), flush=True)
        e_base, e_new = rel(base, ref), rel(new, ref)
        per_kd = [rel(new[:, :, kd], ref[:, :, kd]) for kd in range(3)]
        per_kw = [rel(new[..., kw], ref[..., kw]) for kw in range(3)]
        good = e_new < 1e-3
        ok &= good
        print("%3d->%3d %s pad %d: mode0 %.2e  stacked %.2e  per kd %s per kw %s  %s" % (
            cin, cout, (n, d, h, w), pad, e_base, e_new, ["%.1e" % v for v in per_kd], ["%.1e" % v for v in per_kw],
            "ok" if good else "MISMATCH"), flush=True)
    if ok:
        for cin, cout, (n, d, h, w), pad in CASES[:4]:
            x = torch.randn((n, d, h, w, cin), device="cuda").bfloat16()
            g = torch.randn((n, d + 2 * pad - 2, h + 2 * pad - 2, w + 2 * pad - 2, cout), device="cuda").bfloat16()
            t0, t1, t2 = (timed(m_, x, g, pad, (cout, cin, 3, 3, 3)) for m_ in (0, 1, 2))
            print("%s: mode0 %.1f us  stacked %.1f us  mode0 + staged drain %.1f us (kernel + reduction, cold L2)" % ((n, d, h, w), t0, t1, t2),
                  flush=True)
    if ok:
        # per-CTA phase clocks (cycles): MMA issue loop, of which waiting for TMA stages, drain warps waiting for the accumulators, drain
        n, d, h, w = CASES[0][2]
        x = torch.randn((n, d, h, w, 64), device="cuda").bfloat16()
        g = torch.randn((n, d, h, w, 64), device="cuda").bfloat16()
        dbg = torch.zeros(160 * 8, dtype=torch.int64, device="cuda")
        lib.call("hpvg_debug_set_clock_buffer", dbg.data_ptr())
        try:
            for mode in (0, 1, 2):
                prev = lib.set_wgrad_mode(mode)
                try:
                    for _ in range(2):
                        dbg.zero_()
                        ops.wgrad_raw(x, g, 1, (64, 64, 3, 3, 3))
                        torch.cuda.synchronize()
                finally:
                    lib.set_wgrad_mode(prev)
                rows = dbg.view(160, 8)[:, :4].float().cpu()
                rows = rows[rows[:, 0] > 0]
                print("mode %d: %d CTAs, mean cycles [mma loop, wait tma, wait acc, drain] = %s" % (
                    mode, rows.shape[0], [int(v) for v in rows.mean(0).tolist()]), flush=True)
        finally:
            lib.call("hpvg_debug_set_clock_buffer", None)
    print("ALL OK" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
