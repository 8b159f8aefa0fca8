"""CUDA-event timing of single libhpvg kernels at the finest-level shapes of BASELINE config 2 (development aid)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "hp-vae-gan_b200"))
import torch
from hpvg import ops, lib

dev = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
D, H, W = (32, 128, 128) if os.environ.get('BENCH_BIG') == '1' else (int(os.environ.get('BENCH_D', '16')), 64, 64)     # BASELINE config 5 / config 2 finest volume
V = D * H * W
flush = torch.empty(256 * 2**20 // 4, device=dev)


GRAPHED = os.environ.get("BENCH_GRAPHED") == "1"


def timeit(fn, name, flops=None, nbytes=None):
    for _ in range(3):
        fn()
    if GRAPHED:
        # 10 launches recorded into a CUDA graph and replayed: no host launch latency between the events (inputs stay in L2)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            fn()
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for _ in range(10):
                fn()
        gr.replay()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); gr.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 100.0)
        ts.sort()
        med = ts[len(ts) // 2]
        msg = "%-34s graphed x10: %8.1f us per launch" % (name, med)
        if flops:
            msg += "  %7.1f TFLOP/s" % (flops / med / 1e6)
        if nbytes:
            msg += "  %7.1f GB/s" % (nbytes / med / 1e3)
        print(msg, flush=True)
        return
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    med = ts[len(ts) // 2]
    msg = "%-34s median %8.1f us  min %8.1f us" % (name, med, ts[0])
    if flops:
        msg += "  %7.1f TFLOP/s" % (flops / med / 1e6)
    if nbytes:
        msg += "  %7.1f GB/s" % (nbytes / med / 1e3)
    print(msg, flush=True)


xw = torch.randn(1, D, H, W, 64, device=dev).bfloat16()
gw = torch.randn(1, D, H, W, 64, device=dev).bfloat16()
w64 = torch.randn(64, 64, 3, 3, 3, device=dev) * 0.02
b64 = torch.zeros(64, device=dev)
xt = torch.randn(1, 3, D, H, W, device=dev)
gt = torch.randn(1, 3, D, H, W, device=dev)
w_head = torch.randn(64, 3, 3, 3, 3, device=dev) * 0.1
w_tail = torch.randn(3, 64, 3, 3, 3, device=dev) * 0.02
stats = torch.zeros(128, device=dev)
F64 = 2.0 * V * 64 * 64 * 27
FN = 2.0 * V * 64 * 3 * 27

if which in ("all", "tc"):
    packed = ops.pack_weights(w64, 64, 64, 27, False)
    def conv_tc_only():
        y = torch.empty_like(xw)
        lib.call("hpvg_conv_forward", xw.data_ptr(), 1, w64.data_ptr(), packed.data_ptr(), b64.data_ptr(), y.data_ptr(), 1, 1, 64, 64, D, H, W, 3, 1, 0,
                 0, 0.0, stats.data_ptr(), None, torch.cuda.current_stream().cuda_stream)
    timeit(conv_tc_only, "conv_tc 64->64 fprop+stats", F64)
    def conv_tc_nostats():
        y = torch.empty_like(xw)
        lib.call("hpvg_conv_forward", xw.data_ptr(), 1, w64.data_ptr(), packed.data_ptr(), b64.data_ptr(), y.data_ptr(), 1, 1, 64, 64, D, H, W, 3, 1, 0,
                 1, 0.2, None, None, torch.cuda.current_stream().cuda_stream)
    timeit(conv_tc_nostats, "conv_tc 64->64 fprop+lrelu", F64)
if which in ("all", "wtc"):
    timeit(lambda: ops.wgrad_raw(xw, gw, 1, (64, 64, 3, 3, 3)), "wgrad_tc 64->64 (+reduce)", F64)
if which in ("all", "narrow"):
    timeit(lambda: ops.conv_raw(xt, w_head, b64, 1, False, True, stats=stats), "head conv 3->64 (+stats)", FN, V * 128 + V * 12)
    timeit(lambda: ops.conv_raw(xw, w_tail, None, 1, False, False), "tail conv 64->3", FN, V * 128 + V * 12)
    timeit(lambda: ops.conv_raw(xw, w_head, None, 1, True, False), "head dgrad 64->3", FN, V * 128 + V * 12)
    timeit(lambda: ops.conv_raw(gt, w_tail, None, 1, True, True), "tail dgrad 3->64", FN, V * 128 + V * 12)
    timeit(lambda: ops.wgrad_raw(xt, gw, 1, (64, 3, 3, 3, 3), want_bias=True), "head wgrad (x thin, gy wide)", FN, V * 128 + V * 12)
    timeit(lambda: ops.wgrad_raw(xw, gt, 1, (3, 64, 3, 3, 3), want_bias=True), "tail wgrad (x wide, gy thin)", FN, V * 128 + V * 12)
if which in ("all", "ew"):
    ss = torch.randn(128, device=dev)
    out = torch.empty_like(xw)
    timeit(lambda: lib.call("hpvg_bn_apply_lrelu", xw.data_ptr(), ss.data_ptr(), out.data_ptr(), V, 64, 0.2, torch.cuda.current_stream().cuda_stream),
           "bn_apply_lrelu", None, V * 256)
    timeit(lambda: lib.call("hpvg_lrelu_bwd", gw.data_ptr(), xw.data_ptr(), out.data_ptr(), V * 64, 0.2, 64, None, torch.cuda.current_stream().cuda_stream),
           "lrelu_bwd", None, V * 384)
    chs = torch.zeros(64, device=dev)
    timeit(lambda: lib.call("hpvg_lrelu_bwd", gw.data_ptr(), xw.data_ptr(), out.data_ptr(), V * 64, 0.2, 64, chs.data_ptr(), torch.cuda.current_stream().cuda_stream),
           "lrelu_bwd + channel sum", None, V * 384)
    timeit(lambda: ops.channel_sum(xw), "channel_sum wide", None, V * 128)
if which == "clk":
    packed = ops.pack_weights(w64, 64, 64, 27, False)
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    lib.call("hpvg_debug_set_clock_buffer", dbg.data_ptr())
    for rep in range(3):
        flush.zero_()
        y = torch.empty_like(xw)
        lib.call("hpvg_conv_forward", xw.data_ptr(), 1, w64.data_ptr(), packed.data_ptr(), b64.data_ptr(), y.data_ptr(), 1, 1, 64, 64, D, H, W, 3, 1, 0,
                 1, 0.2, stats.data_ptr() if rep == 2 else None, None, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        d = dbg.view(148, 8).cpu()
        print("rep", rep, "per-CTA clocks [mma loop, wait weights, wait slabs, epilogue total, epilogue wait acc]")
        for b in (0, 1, 64, 127):
            print("  cta", b, d[b, :5].tolist())
        print("  mean", [int(v) for v in d[:128, :5].float().mean(0).tolist()])
        g0, g1 = d[:128, 5], d[:128, 6]
        print("  globaltimer (ns): kernel span first CTA start -> last CTA end %d, CTA start spread %d, per-CTA duration mean %d max %d"
              % (int(g1.max() - g0.min()), int(g0.max() - g0.min()), int((g1 - g0).float().mean()), int((g1 - g0).max())))
    print("thin 64->3")
    packed_t = ops.pack_weights(w_tail, 3, 64, 27, False, rows=16)
    b3 = torch.zeros(3, device=dev)
    for rep in range(3):
        flush.zero_()
        y = torch.empty(1, 3, D, H, W, device=dev)
        lib.call("hpvg_conv_forward", xw.data_ptr(), 1, w_tail.data_ptr(), packed_t.data_ptr(), b3.data_ptr(), y.data_ptr(), 0, 1, 64, 3, D, H, W, 3, 1, 0,
                 0, 0.0, None, None, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        d = dbg.view(148, 8).cpu()
        print("  rep", rep, "mean", d[:128, :5].float().mean(0).tolist())
    lib.call("hpvg_debug_set_clock_buffer", None)

if which == "clkw":
    dbg = torch.zeros(160 * 8, dtype=torch.int64, device=dev)
    lib.call("hpvg_debug_set_clock_buffer", dbg.data_ptr())
    for rep in range(3):
        flush.zero_() if rep == 0 else None
        ops.wgrad_raw(xw, gw, 1, (64, 64, 3, 3, 3))
        torch.cuda.synchronize()
        d = dbg.view(160, 8).cpu()
        print("rep", rep, "wgrad_tc per-CTA mean [mma loop, wait tma, drain wait acc, drain]", [int(v) for v in d[:147, :4].float().mean(0).tolist()])
    lib.call("hpvg_debug_set_clock_buffer", None)
