#!/usr/bin/env python
"""bench.py — finest-scale HP-VAE-GAN training throughput (and generation frames/s) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl hpvg|reference] [--draws D]

Workload (BASELINE.json configs[1]): train_video.py's 3-D HP-VAE-GAN on one synthetic 16-frame 64x64 clip, nfc 64,
latent 128, vae-levels 3, sampling rates 5 3 1 -> 5 pyramid levels, finest level = 16 x 64 x 64 (a GAN level).
A step = ONE iteration of the reference's loop at that level (train_video.py:126-202): generator 'rec' and 'rand'
passes over the whole pyramid, four critic passes, the WGAN-GP double backward, both backward sweeps, gradient
clipping and both Adam steps.  1 720.43 GFLOP of convolution work per step.

  value : iterations/s with the clip resident in HBM, CUDA-event timed, max over ranks.
  e2e   : the same loop fed from pinned HOST buffers every step (H2D of the clip inside the timed region) with a
          device->host read of the step's reconstruction loss.
  N > 1 : batched-noise data-parallel training — one clip + its noise per GPU, replicated weights, the gradients of each
          backward averaged by one libhpvg kernel over NVLink peer memory (hpvg_peer_allreduce_avg_tensors; NCCL all-reduce
          where the GPUs cannot map each other's memory; `allreduce` in the line says which ran)
          (weak scaling: value = N clip-iterations per iteration time).
  roofline : the dominant kernel's per-launch time (CUDA events on its stream) against the MEASURED burst bf16 peak.
  gpu_eager_baseline : the reference's PyTorch path (oracle/port.py restates it op for op) in torch eager + cuDNN TF32 on the
          same GPU — what the unmodified reference resolves to on this hardware.
  parity : one recorded-graph replay of the benched iteration against the CPU oracle on identical weights and draws.
  --impl reference : the CPU arm — the oracle's restatement of the same iteration (oracle/train_ref.py, pinned to the
          reference by tests/golden/train_*.pt) in PyTorch-CPU fp32 on all host cores.  The reference itself is a
          Python package under /root/reference, which does not exist on the GPU box.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "hp-vae-gan_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

CONV_GFLOP_PER_ITER = 1720.43      # SURVEY.md App. B, config 2 finest level (config 5: 12 650.99): what the REFERENCE executes
# the generator step's -D(fake) backward also produces critic weight gradients that D.zero_grad() discards unread
# (train_video.py:168); ScaleTrainer(skip_critic_grads=True, the default) does not launch them: 3 -> 64, 5 x 64 -> 64, 64 -> 1
# weight-gradient passes at the finest level
def skipped_gflop(voxels):
    return 2.0 * voxels * 27 * (3 * 64 + 5 * 64 * 64 + 64 * 1) / 1e9
METRIC = "train_iters_per_s_finest_scale"
UNIT = "iter/s"


# ---------------------------------------------------------------------------------------------------------------
# workload
# ---------------------------------------------------------------------------------------------------------------
WORKLOAD = {"name": "cfg2"}


def make_opt():
    from hpvg.options import Options   # the argparse-namespace stand-in with the reference's defaults
    if WORKLOAD["name"] == "cfg5":     # BASELINE configs[4]: 32-frame 128x128 clip (--img-size 128 --sampling-rates 31 1)
        o = Options(img_size=128, sampling_rates=[31, 1], vae_levels=3, nfc=64, latent_dim=128, num_layer=5, batch_size=1)
    else:
        o = Options(img_size=64, sampling_rates=[5, 3, 1], vae_levels=3, nfc=64, latent_dim=128, num_layer=5, batch_size=1)
    o.scale_idx = o.stop_scale
    o.Noise_Amps = [1.0] + [0.07] * (o.stop_scale - 1)      # survey-observed amplitudes; the finest one is computed at iteration 0
    t0, h0, w0 = o.level_size(0)
    o.Z_init_size = [1, o.latent_dim, t0, h0, w0]
    return o


def level_shape(o, idx):
    t, h, w = o.level_size(idx)
    return (1, 3, t, h, w)


def workload_name(o):
    if WORKLOAD["name"] == "cfg5":
        return ("configs[4]: 3D HP-VAE-GAN, synthetic 32-frame 128x128 clip, vae-levels 3, nfc 64, rates 31 1, "
                "finest level %d of %d (GAN), batch 1 per GPU" % (o.scale_idx, o.stop_scale))
    return ("configs[1]: train_video.py 3D HP-VAE-GAN, synthetic 16-frame 64x64 clip, vae-levels 3, nfc 64, rates 5 3 1, "
            "finest level %d of %d (GAN), batch 1 per GPU" % (o.scale_idx, o.stop_scale))


def synthetic_clip(o, seed):
    g = torch.Generator().manual_seed(seed)
    real = torch.rand(level_shape(o, o.scale_idx), generator=g) * 2 - 1
    real_zero = torch.rand(level_shape(o, 0), generator=g) * 2 - 1
    return real, real_zero


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port on the host cores)
# ---------------------------------------------------------------------------------------------------------------
def cpu_iteration_timer(o, state_g, state_d, budget_s, steps, warmup):
    """time `steps` iterations (after `warmup`) of the oracle's restatement of the same iteration; stops early when the
    budget is exhausted.  Returns (seconds per iteration, iterations timed, cores)."""
    from oracle import train_ref
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(cores)
    sd_g = {k: v.detach().clone().float().cpu() for k, v in state_g.items()}
    sd_d = {k: v.detach().clone().float().cpu() for k, v in state_d.items()}
    oc = make_opt()
    tr = train_ref.ScaleTrainer(oc, sd_g, sd_d)
    real, real_zero = synthetic_clip(oc, 0)
    torch.manual_seed(0)
    t_start = time.perf_counter()
    for _ in range(warmup):
        tr.iteration(real, real_zero)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        tr.iteration(real, real_zero)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > budget_s:
            break
    return sum(times) / len(times), len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    o = make_opt()
    sg, sd = fresh_states(o)
    warm = max(1, min(args.warmup, 1))
    sec, done, cores = cpu_iteration_timer(o, sg, sd, budget_s=240.0, steps=args.steps, warmup=warm)
    val = 1.0 / sec
    sample = "%d full iteration(s) of the same workload after %d warm-up (oracle/train_ref.py, PyTorch-CPU fp32, %d threads)" % (done, warm, cores)
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": done, "warmup": warm,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": workload_name(o)},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def fresh_states(o):
    """default-initialised generator / critic state_dicts for the workload (seed 0), built on the CPU with torch modules
    of the drop-in package (construction only: no kernel runs here)"""
    from modules import networks_3d
    torch.manual_seed(0)
    g = networks_3d.GeneratorHPVAEGAN(o)
    for _ in range(o.scale_idx):
        g.init_next_stage()
    d = networks_3d.WDiscriminator3D(o)
    return g.state_dict(), d.state_dict()


# ---------------------------------------------------------------------------------------------------------------
# torch-eager arm on the GPU: the reference's own PyTorch path (oracle/port.py restates its modules op for op) with cuDNN
# TF32 convolutions — "the real kernel to beat on the same box" 
# ---------------------------------------------------------------------------------------------------------------
def gpu_eager_timer(o_name, state_g, state_d, dev, steps, warmup):
    from oracle import train_ref
    prev = WORKLOAD["name"]
    WORKLOAD["name"] = o_name
    try:
        oc = make_opt()
        sd_g = {k: v.detach().clone().float().to(dev) for k, v in state_g.items()}
        sd_d = {k: v.detach().clone().float().to(dev) for k, v in state_d.items()}
        tr = train_ref.ScaleTrainer(oc, sd_g, sd_d)
        real, real_zero = synthetic_clip(oc, 0)
        real, real_zero = real.to(dev), real_zero.to(dev)
        tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = True          # PyTorch's default for convolutions: what the unmodified scripts get
        for _ in range(warmup):
            tr.iteration(real, real_zero)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            tr.iteration(real, real_zero)
        e1.record()
        torch.cuda.synchronize()
        torch.backends.cudnn.allow_tf32 = tf32[0]
        return e0.elapsed_time(e1) / steps
    finally:
        WORKLOAD["name"] = prev


# ---------------------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------------------
class TrainLeg:
    """one workload's training measurement: networks, trainer, recorded iteration, timed steps"""

    def __init__(self, name, args, rank, world, dev, distributed):
        from hpvg import lib, train
        from modules import networks_3d
        self.name, self.args, self.rank, self.world, self.dev, self.distributed = name, args, rank, world, dev, distributed
        WORKLOAD["name"] = name
        self.o = o = make_opt()
        self.sg, self.sd = fresh_states(o)             # identical on every rank (seed 0): replicated weights
        self.G = networks_3d.GeneratorHPVAEGAN(o)
        for _ in range(o.scale_idx):
            self.G.init_next_stage()
        self.G.load_state_dict(self.sg)
        self.D = networks_3d.WDiscriminator3D(o)
        self.D.load_state_dict(self.sd)
        self.G.to(dev)
        self.D.to(dev)
        self.use_graph = not args.no_graph
        self.trainer = train.ScaleTrainer(o, self.G, self.D, distributed=distributed, capturable=self.use_graph)
        real_h, real_zero_h = synthetic_clip(o, rank)  # one clip per rank
        self.real_h, self.real_zero_h = real_h.pin_memory(), real_zero_h.pin_memory()
        self.real, self.real_zero = self.real_h.to(dev), self.real_zero_h.to(dev)
        self.last = {}
        self.W = max(3, args.warmup)
        self.gflop = CONV_GFLOP_PER_ITER if name == "cfg2" else 12650.99
        t, h, w = o.level_size(o.scale_idx)
        self.gflop_executed = self.gflop - (skipped_gflop(t * h * w) if self.trainer.skip_critic_grads else 0.0)

    def barrier(self):
        import torch.distributed as dist
        if self.distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps):
        import torch.distributed as dist
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    def prepare(self):
        from hpvg import lib
        args, trainer, last = self.args, self.trainer, self.last
        if self.use_graph:
            # the whole iteration (every libhpvg kernel, autograd glue, clipping, both Adam steps, the all-reduces) is
            # recorded once into a CUDA graph after W eager warm-up iterations and replayed per step
            c0 = lib.launch_count()
            trainer.capture(self.real, self.real_zero, warmup=self.W, candidates=args.graph_candidates)
            torch.cuda.synchronize()
            last["pool_mb"] = getattr(trainer, "pool_bytes", 0) / 2**20
            # launches counted per recorded iteration: W eager warm-up iterations + the recording pass of the first candidate, then
            # 1 warm-up + 1 recording pass per further candidate (replays launch nothing through the library's host side)
            self.launches_per_iter = (lib.launch_count() - c0) // (self.W + 1 + 2 * (max(1, args.graph_candidates) - 1))

            def step_resident():
                trainer.replay()

            def step_e2e():
                out = trainer.replay(self.real_h, self.real_zero_h)          # H2D of the clip into the graph's input buffers
                last["rec_loss"] = out["rec_loss"].item()                    # device -> host read of the step's result
            for _ in range(2 + args.settle_steps):
                step_resident()
            last["extra_warmup"] = 2 + args.settle_steps
        else:
            self.launches_per_iter = None

            def step_resident():
                trainer.iteration(self.real, self.real_zero)

            def step_e2e():
                r = self.real_h.to(self.dev, non_blocking=True)
                rz = self.real_zero_h.to(self.dev, non_blocking=True)
                out = trainer.iteration(r, rz)
                last["rec_loss"] = out["rec_loss"].item()
            for _ in range(self.W):
                step_resident()
        self.step_resident, self.step_e2e = step_resident, step_e2e

    def measure(self):
        from hpvg import lib
        args = self.args
        n0 = lib.launch_count()
        self.ms = self.timed(self.step_resident, args.steps)
        self.launches = lib.launch_count() - n0 if not self.use_graph else self.launches_per_iter * args.steps
        self.ms_e2e = self.timed(self.step_e2e, args.steps)
        self.value = self.world * args.steps / (self.ms * 1e-3)
        self.e2e = self.world * args.steps / (self.ms_e2e * 1e-3)
        self.h2d = (self.real_h.numel() + self.real_zero_h.numel()) * 4

    def profile_kernels(self, prof_steps):
        """the same steps again with CUDA events around every convolution launch (on the launching stream): one kernel at a
        time — the side streams are joined into the main stream and the programmatic launch overlap is off for this leg, so an
        event pair brackets exactly one kernel running alone"""
        from hpvg import lib
        trainer = self.trainer

        def step_eager():
            # keep the GPU busy while the host enqueues the iteration, so that the event pairs bracket back-to-back kernels
            # and not host launch latency (eager launching is CPU-bound on this workload)
            torch.cuda._sleep(int(0.12 * 1.9e9))
            trainer.iteration(self.real, self.real_zero)
        saved = (trainer.overlap, trainer._side, trainer._wside)
        trainer.overlap, trainer._side, trainer._wside = False, None, None
        pdl_was = lib.set_pdl(False)
        step_eager()
        torch.cuda.synchronize()
        lib.profile_enable(True)
        self.timed(step_eager, prof_steps)
        lib.profile_enable(False)
        lib.set_pdl(pdl_was)
        trainer.overlap, trainer._side, trainer._wside = saved
        return lib.profile_dump()


KERNEL_NAMES = {"conv_tc": "conv_tc_kernel / conv_col_kernel (tcgen05 implicit-GEMM conv: forward of the critic's layers, every data gradient)",
                "conv_bn_fused": "conv_tc_kernel<FUSE> (tcgen05 conv + BatchNorm batch statistics + LeakyReLU in one launch, grid barrier)",
                "wgrad_tc": "wgrad_tc_kdstack_kernel + wgrad_reduce_kernel (tcgen05 weight gradient, split-K over the SMs)",
                "conv_expand": "expand_conv_mma_kernel (3 -> 64 heads, TF32 mma.sync)", "wgrad_narrow": "outer_corr_mma_kernel (3 / 1-channel ends)"}


def ncu_traffic(kind):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the kernel's 64 -> 64 @ 16 x 64 x 64 shape from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json is written from the capture's csv by experiments/ncu_traffic.py)"""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(kind)
        return (t["dram_read_bytes"] + t["dram_write_bytes"], t["source"]) if t else (None, None)
    except (OSError, KeyError, ValueError):
        return None, None


def roofline_from(rows, prof_steps, ms_per_step, workload):
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except OSError:
        pass
    # per-launch CUDA events around a kernel running alone inside a ~10 ms eager step at full clocks: the BURST figure applies
    peak_tf = peaks.get("bf16_tflops")
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone, per launch)"
    if not peak_tf:
        peak_tf, peak_src = 1600.0, "fallback (B200_PROFILING.md burst bf16)"
    by_kind = {}
    for r in rows:
        k = by_kind.setdefault(r["kind"], {"launches": 0, "ms": 0.0, "flops": 0.0})
        k["launches"] += r["launches"]; k["ms"] += r["ms"]; k["flops"] += r["work"] * r["launches"]
    tensor_kinds = ("conv_tc", "conv_bn_fused", "wgrad_tc")
    cand = [r for r in rows if r["kind"] in tensor_kinds]
    if not cand:
        return None
    dominant_kind = max(tensor_kinds, key=lambda k: by_kind.get(k, {"ms": 0.0})["ms"])
    top = max((r for r in cand if r["kind"] == dominant_kind), key=lambda r: r["ms"])

    def describe(r):
        us = r["ms"] / r["launches"] * 1e3
        tf = r["work"] / (us * 1e-6) / 1e12
        return {"kind": r["kind"], "gflop_per_launch": r["work"] / 1e9, "launches_timed": r["launches"], "us_per_launch": us,
                "tflops": tf, "frac_of_peak": tf / peak_tf}
    per_launch_ms = top["ms"] / top["launches"]
    achieved = top["work"] / (per_launch_ms * 1e-3) / 1e12
    traffic, traffic_src = ncu_traffic(top["kind"]) if workload == "cfg2" else (None, None)
    return {"bound": "tensor", "kernel": KERNEL_NAMES.get(top["kind"], top["kind"]) + ", its heaviest shape in the step (%.2f GFLOP per launch)" % (top["work"] / 1e9),
            "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
            "traffic": traffic, "traffic_unit": "bytes per launch", "traffic_source": traffic_src,
            "flops_per_launch": top["work"], "us_per_launch": per_launch_ms * 1e3, "launches_timed": top["launches"],
            "peak_source": peak_src,
            "share_of_step": (by_kind[dominant_kind]["ms"] / prof_steps) / ms_per_step,
            "by_kernel_ms_per_step": {k: v["ms"] / prof_steps for k, v in by_kind.items()},
            "by_kernel_tflops": {k: (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None) for k, v in by_kind.items()},
            "heaviest_shape_per_kernel": [describe(max((r for r in cand if r["kind"] == k), key=lambda r: r["ms"]))
                                          for k in tensor_kinds if any(r["kind"] == k for r in cand)],
            "note": "per-launch CUDA events need eager launches: this leg runs the same iteration un-graphed on ONE stream "
                    "(no concurrent side-stream kernels, no programmatic launch overlap) behind a GPU-side delay so "
                    "that launches are queued ahead of the GPU; share_of_step = the dominant kernel's ms per iteration / "
                    "graph-replay ms per iteration (the replay overlaps streams, so shares can sum above 1)"}


def dependent_chain_leg(dev):
    """The dominant kernel as it runs inside the recorded iteration: TEN dependent launches of the 64 -> 64 tcgen05 convolution at
    the finest level's shape (16 x 64 x 64, output of one launch = input of the next, two 8.4 MB tensors ping-pong), recorded into a
    CUDA graph and timed with ONE CUDA-event pair around the replay.  No per-launch event records between the kernels: the figure
    is kernel + the dependency gap to the next kernel, the cost a layer has in the benched iteration.  (Per-launch event pairs, the
    `roofline` object's primary figure, add the front-end latency of an isolated launch: ncu's sm__cycles_active for the same
    launch is 32.4 k cycles = 16.5 us against 24 us of gpu__time_duration, profiles/r02k_ncu_full_top_kernels.txt.)"""
    from hpvg import lib, ops
    d, h, w = 16, 64, 64
    x = torch.randn(1, d, h, w, 64, device=dev).bfloat16()
    y = torch.empty_like(x)
    wt = torch.randn(64, 64, 3, 3, 3, device=dev) * 0.03
    bias = torch.zeros(64, device=dev)
    packed = ops.pack_weights(wt, 64, 64, 27, False)

    def chain():
        for i in range(10):
            src, dst = (x, y) if i % 2 == 0 else (y, x)
            lib.call("hpvg_conv_forward", src.data_ptr(), 1, wt.data_ptr(), packed.data_ptr(), bias.data_ptr(), dst.data_ptr(), 1, 1, 64, 64, d, h, w,
                     3, 1, 0, 1, 0.2, None, None, torch.cuda.current_stream().cuda_stream)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        chain()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain()
    for _ in range(3):
        gr.replay()
    ts = []
    for _ in range(15):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 100.0)      # ms per 10 launches -> us per launch
    ts.sort()
    flops = 2.0 * d * h * w * 64 * 64 * 27
    us = ts[len(ts) // 2]
    return {"what": "10 dependent launches of conv_tc_kernel (64 -> 64, 16 x 64 x 64) in one CUDA graph, one event pair around the replay "
                    "(median of 15); inputs ping-pong between two 8.4 MB tensors, i.e. they come from L2 as inside the iteration",
            "us_per_launch": us, "tflops": flops / us / 1e6, "flops_per_launch": flops}


def parity_leg(leg):
    """The benched path against the CPU oracle: fresh copies of the workload's networks, iteration 0 eagerly (it computes the
    noise amplitude on the host), iteration 1 as the warm-up step of the recording, iteration 2 as ONE REPLAY of the recorded CUDA
    graph — all on injected draws that the oracle (oracle/train_ref.py, fp32) steps too.  Raises if the replayed
    reconstruction loss is off by more than 1 % (north_star's end-of-scale criterion)."""
    from hpvg import train
    from modules import networks_3d
    from oracle import train_ref
    WORKLOAD["name"] = leg.name
    o_g, o_c = make_opt(), make_opt()
    G = networks_3d.GeneratorHPVAEGAN(o_g)
    for _ in range(o_g.scale_idx):
        G.init_next_stage()
    G.load_state_dict(leg.sg)
    D = networks_3d.WDiscriminator3D(o_g)
    D.load_state_dict(leg.sd)
    G.to(leg.dev); D.to(leg.dev)
    sd_g = {k: v.detach().clone().float() for k, v in leg.sg.items()}
    sd_d = {k: v.detach().clone().float() for k, v in leg.sd.items()}
    real, real_zero = synthetic_clip(o_g, 0)
    gen = torch.Generator().manual_seed(123)
    z = tuple(o_g.Z_init_size)
    gan_levels = list(range(o_g.vae_levels, o_g.stop_scale + 1))
    draws = []
    for it in range(3):
        dr = {"noise_init": torch.randn(z, generator=gen)}
        if it == 0:
            dr["eps_amp"] = torch.randn(z, generator=gen)
        dr["eps"] = torch.randn(z, generator=gen)
        dr["noises"] = {lvl: torch.randn(level_shape(o_g, lvl), generator=gen) for lvl in gan_levels}
        dr["alpha"] = 0.25 + 0.25 * it
        draws.append(dr)

    def flat(dr):
        return [dr["noise_init"]] + ([dr["eps_amp"]] if "eps_amp" in dr else []) + [dr["eps"]] + [dr["noises"][l] for l in gan_levels]
    tr = train.ScaleTrainer(o_g, G, D, capturable=True)
    feed = train.NoiseFeed(leg.dev)
    rd, rzd = real.to(leg.dev), real_zero.to(leg.dev)
    with feed:
        feed.load(flat(draws[0]), draws[0]["alpha"])
        tr.iteration(rd, rzd)                          # iteration 0, eager: computes the level's noise amplitude on the host
        # iteration 1 is the eager warm-up step capture() needs; the recording pass that follows executes nothing
        feed.load(flat(draws[1]), draws[1]["alpha"])
        stock = tr.iteration

        def rewinding(a, b):
            feed.rewind()
            return stock(a, b)
        tr.iteration = rewinding
        try:
            tr.capture(rd, rzd, warmup=1)
        finally:
            tr.iteration = stock
        torch.cuda.synchronize()
        feed.load(flat(draws[2]), draws[2]["alpha"])
        out = tr.replay()                              # iteration 2: one replay of the recorded iteration
        got = {k: v.item() for k, v in out.items()}
    oracle = train_ref.ScaleTrainer(o_c, sd_g, sd_d)
    for dr in draws:
        ref = oracle.iteration(real, real_zero, noise_init=dr["noise_init"], eps=dr["eps"], noises=dr["noises"], alpha=dr["alpha"],
                               eps_amp=dr.get("eps_amp"))
    ref = {k: v.item() for k, v in ref.items()}
    rel = {k: abs(got[k] - ref[k]) / abs(ref[k]) for k in ("rec_loss", "gradient_penalty")}
    res = {"what": "third iteration of the level, replayed from the recorded CUDA graph, vs oracle/train_ref.py (CPU fp32) stepping the "
                   "same three iterations on identical weights and draws",
           "rec_loss_gpu": got["rec_loss"], "rec_loss_cpu": ref["rec_loss"], "gradient_penalty_gpu": got["gradient_penalty"],
           "gradient_penalty_cpu": ref["gradient_penalty"], "rel_err": rel, "tolerance": {"rec_loss": 1e-2, "gradient_penalty": 2e-2},
           "noise_amp_gpu": o_g.Noise_Amps[-1], "noise_amp_cpu": o_c.Noise_Amps[-1]}
    # north_star: reconstruction loss within 1 %.  The penalty is lambda * mean((|grad| - 1)^2) of a critic two Adam steps away
    # from its initialisation, a far more sensitive quantity (measured: 0.2 - 0.5 % between eager runs of this build): 2 %
    if rel["rec_loss"] > 1e-2 or rel["gradient_penalty"] > 2e-2:
        raise AssertionError("bench parity leg: the recorded iteration disagrees with the CPU oracle: %s" % json.dumps(res))
    return res


def generation_leg(leg, args):
    """BASELINE configs[3]: fresh z per draw through the whole pyramid, draws split over ranks, no collective.
    value: frames stay in HBM.  e2e: every batch is converted to uint8 HWC frames on the device (hpvg_frames_to_uint8_batched,
    what utils/saver.py::write_video does on the host) and copied to pinned host memory inside the timed region."""
    from hpvg import data, train
    _, draws_rank = train.draws_for_rank(args.draws, leg.world, leg.rank)
    draws_rank = max(1, draws_rank)
    gen_batch = max(1, min(args.gen_batch, draws_rank))
    gen_calls = max(1, draws_rank // gen_batch)
    draws_rank = gen_calls * gen_batch
    # batch > 1: every draw is normalised with its own BatchNorm statistics (ops.bn_per_sample), i.e. the reference's batch-1 draws
    sampler = train.Sampler(leg.G, leg.o, leg.dev, batch=gen_batch, graph=leg.use_graph, streams=args.gen_streams, static_weights=True)
    frames = [0]

    def gen_all():
        n = 0
        sampler.begin()
        for _ in range(gen_calls):
            n += sampler.frames_per_call(sampler.sample())
        sampler.wait()
        frames[0] = n
    gen_all()
    ms_gen = leg.timed(gen_all, 1)
    shape = level_shape(leg.o, leg.o.scale_idx)
    t, h, w = shape[2:]
    nslots = max(1, sampler.nstreams)
    dev_u8 = [torch.empty((gen_batch, t, h, w, 3), dtype=torch.uint8, device=leg.dev) for _ in range(nslots)]
    host_u8 = [torch.empty((gen_batch, t, h, w, 3), dtype=torch.uint8).pin_memory() for _ in range(nslots)]
    d2h = [0]

    def gen_all_e2e():
        n, bytes_out = 0, 0
        sampler.begin()
        for c in range(gen_calls):
            k = sampler.next
            fake = sampler.sample()
            st = sampler.streams[k] if sampler.nstreams > 1 else torch.cuda.current_stream()
            with torch.cuda.stream(st):
                data.to_uint8_frames(fake, out=dev_u8[k])
                host_u8[k].copy_(dev_u8[k], non_blocking=True)
            n += sampler.frames_per_call(fake)
            bytes_out += host_u8[k].numel()
        sampler.wait()
        torch.cuda.current_stream().synchronize()
        frames[0], d2h[0] = n, bytes_out
    gen_all_e2e()
    ms_gen_e2e = leg.timed(gen_all_e2e, 1)
    return {"metric": "generated_frames_per_s", "value": leg.world * frames[0] / (ms_gen * 1e-3), "unit": "frames/s",
            "draws": draws_rank * leg.world, "frames_per_draw": t, "batch": gen_batch,
            "batchnorm": "per-draw statistics (each draw normalised as in a batch-1 forward)",
            "ms_per_draw": ms_gen / draws_rank, "streams": sampler.nstreams,
            "e2e": {"value": leg.world * frames[0] / (ms_gen_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": d2h[0], "what": "uint8 [T,H,W,3] frames of every draw (device-side conversion) copied to "
                    "pinned host memory inside the timed region; a step = this rank's share of the draws"},
            "note": "draws split over ranks, no collective; rank 0's share timed x N (equal shares)"}


def run_hpvg(args):
    import torch.distributed as dist
    from hpvg import lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl hpvg needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    leg = TrainLeg(args.workload, args, rank, world, dev, distributed)
    torch.manual_seed(1 + rank)                    # per-rank noise; the GP alpha comes from the CPU generator below
    cpu_gen_state = torch.random.get_rng_state()
    leg.prepare()
    sys.stderr.write("graph recordings, ms per replay: %s\n" % ", ".join("%.3f" % t for t in getattr(leg.trainer, "capture_trials", [])))
    use_graph, trainer, o = leg.use_graph, leg.trainer, leg.o
    if use_graph and args.recapture < 0:
        # diagnostic: per-replay GPU time (events) and host enqueue time of consecutive replays of one recording
        import time as _time
        nrep = -args.recapture if args.recapture < -1 else 48
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(nrep + 1)]
        host = []
        torch.cuda.synchronize()
        evs[0].record()
        for k in range(nrep):
            t0 = _time.perf_counter()
            leg.step_resident()
            host.append((_time.perf_counter() - t0) * 1e3)
            evs[k + 1].record()
        torch.cuda.synchronize()
        sys.stderr.write("per-replay ms (gpu): %s\n" % " ".join("%.2f" % evs[k].elapsed_time(evs[k + 1]) for k in range(nrep)))
        sys.stderr.write("per-replay ms (host enqueue): %s\n" % " ".join("%.2f" % h for h in host))
    if use_graph and args.recapture > 0:
        # diagnostic: replay time of several recordings of the same iteration in ONE process
        for k in range(args.recapture + 1):
            for _ in range(3):
                leg.step_resident()
            sys.stderr.write("recording %d: %.3f ms per replay\n" % (k, leg.timed(leg.step_resident, 10) / 10))
            if k < args.recapture:
                trainer.capture(leg.real, leg.real_zero, warmup=1)
    if args.profile_gen:
        from hpvg import train
        # for `ncu --profile-from-start off`: one forward of the generation leg (gen-batch draws, one stream) between Start/Stop
        ps = train.Sampler(leg.G, o, dev, batch=max(1, args.gen_batch), graph=use_graph, streams=1, static_weights=True)
        ps.sample()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        ps.sample()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        _emit({"profiled": "one generation forward", "batch": args.gen_batch, "graph": use_graph})
        return
    if args.profile_one:
        # for `ncu --profile-from-start off`: exactly one replayed (or eager) iteration between cudaProfilerStart/Stop
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        leg.step_resident()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        _emit({"profiled": "one iteration", "graph": use_graph})
        return
    torch.cuda.reset_peak_memory_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    leg.measure()
    clocks = sampler.stop() if rank == 0 else None
    peak_mb = torch.cuda.max_memory_allocated() / 2**20
    # working set of one iteration: the memory pool of the recorded graph (every activation saved for the backward passes,
    # every gradient) when graphed, else the allocator's peak during the timed steps
    work_mb = leg.last.get("pool_mb") or peak_mb
    prof_steps = min(args.steps, 3)
    rows = leg.profile_kernels(prof_steps)
    generation = generation_leg(leg, args)

    # multi-GPU runs of the metric's workload also measure BASELINE configs[4], the configuration SURVEY.md §8e names for the
    # batched-noise data-parallel mode (32 x 128 x 128, one clip per GPU): same loop, same all-reduces, 7.4x the work per step
    dp_named = None
    if distributed and args.workload == "cfg2" and not args.no_cfg5:
        big = TrainLeg("cfg5", args, rank, world, dev, distributed)
        big.prepare()
        big.measure()
        if rank == 0:
            dp_named = {"workload": workload_name(big.o), "value": big.value, "unit": UNIT, "ms_per_step": big.ms / args.steps,
                        "e2e": {"value": big.e2e, "unit": UNIT, "h2d_bytes_per_step": big.h2d, "d2h_bytes_per_step": 4},
                        "allreduce_bytes_per_step": big.trainer.allreduce_bytes_per_iter,
                        "model_tflops_per_gpu": big.value * big.gflop_executed / 1e3 / world, "conv_gflop_per_iter_executed": big.gflop_executed}
        WORKLOAD["name"] = args.workload
        del big
        torch.cuda.empty_cache()

    if rank == 0:
        W = leg.W
        roofline = roofline_from(rows, prof_steps, leg.ms / args.steps, args.workload)
        if roofline is not None and args.workload == "cfg2":
            chain = dependent_chain_leg(dev)
            chain["frac_of_peak"] = chain["tflops"] / roofline["peak"]
            roofline["dependent_chain"] = chain
        line = {"metric": METRIC, "value": leg.value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W + leg.last.get("extra_warmup", 0),
                "ms_per_step": leg.ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload_name(o), "parallelism": "dp%d (one clip per GPU, flat gradient bucket averaged %s)" % (world, ALLREDUCE_KINDS.get(trainer.bucketG.kind, str(trainer.bucketG.kind))) if distributed else "single GPU",
                           "l2": "no explicit flush: one iteration writes and re-reads a %.0f MB working set (activations saved for the "
                                 "backward passes, gradients), %s the 126 MB L2" % (work_mb, "above" if work_mb > 126 else "NOT above"),
                           "conv_gflop_per_iter": leg.gflop, "conv_gflop_per_iter_executed": leg.gflop_executed,
                           "warmup": "%d eager iterations + %d replays of the recorded iteration" % (W, leg.last.get("extra_warmup", 0))
                                     if use_graph else "%d eager iterations" % W,
                           "launch": ("one CUDA graph replay per iteration (%d libhpvg kernels recorded)" % leg.launches_per_iter) if use_graph
                           else "eager launches"},
                "e2e": {"value": leg.e2e, "unit": UNIT, "h2d_bytes_per_step": leg.h2d, "d2h_bytes_per_step": 4, "ms_per_step": leg.ms_e2e / args.steps},
                "gpu_launches": leg.launches, "clocks": clocks, "roofline": roofline,
                # FLOPs that actually run: the reference's count minus the critic weight gradients nobody reads (skip_critic_grads)
                "model_tflops": leg.value * leg.gflop_executed / 1e3 / world,
                "generation": generation}
        if distributed:
            line["allreduce_bytes_per_step"] = trainer.allreduce_bytes_per_iter
            line["allreduce"] = {"critic": trainer.bucketD.kind, "generator": trainer.bucketG.kind}
            if dp_named is not None:
                line["dp_named_config"] = dp_named
        if world == 1 and not args.no_cpu_baseline:
            torch.random.set_rng_state(cpu_gen_state)
            sec, done, cores = cpu_iteration_timer(o, leg.sg, leg.sd, budget_s=30.0, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d full iteration(s) of the same workload after 1 warm-up (oracle/train_ref.py, PyTorch-CPU fp32)" % done}
            eager_steps = 10 if args.workload == "cfg2" else 3
            ms_eager = gpu_eager_timer(args.workload, leg.sg, leg.sd, dev, steps=eager_steps, warmup=3)
            line["gpu_eager_baseline"] = {"value": 1e3 / ms_eager, "unit": UNIT, "ms_per_step": ms_eager, "steps": eager_steps,
                                          "kind": "port", "what": "the reference's PyTorch path (oracle/port.py, op for op) in torch eager on this "
                                          "GPU: cuDNN convolutions with allow_tf32 (PyTorch's default), fp32 storage, ATen BatchNorm / "
                                          "LeakyReLU / upsample kernels, torch.optim.Adam",
                                          "speedup_of_value": leg.value / (1e3 / ms_eager)}
            line["parity"] = parity_leg(leg)
        _emit(line)
    if distributed:
        # all ranks are done once rank 0 has printed; leave without tearing NCCL down: destroy_process_group() was seen
        # to hang while CUDA graphs holding captured all-reduces are alive
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ---------------------------------------------------------------------------------------------------------------
# the other single-GPU BASELINE configurations: configs[0] (train_image.py, 2-D, 128 px) and configs[2] (train_video_baselines.py,
# GeneratorSG).  Same metric (finest-level training iterations/s), value + e2e only.
# ---------------------------------------------------------------------------------------------------------------
ALLREDUCE_KINDS = {"peer": "by one libhpvg kernel over NVLink peer memory (hpvg_peer_allreduce_avg)", "nccl": "by one NCCL all-reduce",
                   "nccl-coalesced": "by a coalesced NCCL group call", "sum": "by all-reduce SUM + divide"}


def run_other(args):
    from hpvg import lib, train
    from hpvg.options import Options
    from modules import networks_2d, networks_3d
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback")
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        raise RuntimeError("--workload %s is a single-GPU configuration" % args.workload)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    lib.load()
    torch.manual_seed(0)
    W = max(3, args.warmup)
    if args.workload == "cfg1":
        o = Options(img_size=128, vae_levels=3, nfc=64, latent_dim=128, num_layer=5, batch_size=1)
        o.scale_idx = o.stop_scale
        o.Noise_Amps = [1.0] + [0.07] * (o.stop_scale - 1)
        _, h0, w0 = o.level_size(0)
        _, h, w = o.level_size(o.scale_idx)
        o.Z_init_size = [1, o.latent_dim, h0, w0]
        G = networks_2d.GeneratorHPVAEGAN(o)
        for _ in range(o.scale_idx):
            G.init_next_stage()
        D = networks_2d.WDiscriminator2D(o)
        G.to(dev); D.to(dev)
        gen = torch.Generator().manual_seed(0)
        real_h = (torch.rand((1, 3, h, w), generator=gen) * 2 - 1).pin_memory()
        second_h = (torch.rand((1, 3, h0, w0), generator=gen) * 2 - 1).pin_memory()
        tr = train.ScaleTrainer(o, G, D, capturable=True, dims=2)
        c0 = lib.launch_count()
        tr.capture(real_h.to(dev), second_h.to(dev), warmup=W)
        per_iter = (lib.launch_count() - c0) // (W + 1)
        for _ in range(2):
            tr.replay()
        last = {}

        def step():
            tr.replay()

        def step_e2e():
            last["rec_loss"] = tr.replay(real_h, second_h)["rec_loss"].item()
        name = ("configs[0]: train_image.py 2D HP-VAE-GAN, synthetic 3-channel 128px image, vae-levels 3, nfc 64, finest level %d of %d "
                "(GAN), batch 1" % (o.scale_idx, o.stop_scale))
        gflop, launch = 181.05, "one CUDA graph replay per iteration (%d libhpvg kernels recorded)" % per_iter
    else:
        o = Options(img_size=64, sampling_rates=[5, 3, 1], nfc=64, num_layer=5, batch_size=1, train_depth=1)
        o.scale_idx = o.stop_scale
        o.Noise_Amps = [1.0] + [0.07] * (o.stop_scale - 1)
        t0, h0, w0 = o.level_size(0)
        t, h, w = o.level_size(o.scale_idx)
        G = networks_3d.GeneratorSG(o)
        for _ in range(o.scale_idx):
            G.init_next_stage()
        D = networks_3d.WDiscriminator3D(o)
        G.to(dev); D.to(dev)
        gen = torch.Generator().manual_seed(0)
        real_h = (torch.rand((1, 3, t, h, w), generator=gen) * 2 - 1).pin_memory()
        second_h = torch.randn((1, 3, t0, h0, w0), generator=gen).pin_memory()        # opt.Z_init (train_video_baselines.py:38-43)
        tr = train.BaselineTrainer(o, G, D)
        real, z_init = real_h.to(dev), second_h.to(dev)
        for _ in range(W):
            tr.iteration(real, z_init)
        last = {}
        c0 = lib.launch_count()
        tr.iteration(real, z_init)
        per_iter = lib.launch_count() - c0

        def step():
            tr.iteration(real, z_init)

        def step_e2e():
            last["rec_loss"] = tr.iteration(real_h.to(dev, non_blocking=True), z_init)["rec_loss"].item()
        name = ("configs[2]: train_video_baselines.py GeneratorSG (SinGAN-3D), train-depth 1, synthetic 16-frame 64x64 clip, rates 5 3 1, "
                "finest level %d of %d, batch 1" % (o.scale_idx, o.stop_scale))
        gflop, launch = 2090.59, "eager launches (%d libhpvg kernels per iteration; this loop has no recorded-graph form)" % per_iter

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    sampler = ClockSampler(0)
    sampler.start()
    ms = timed(step, args.steps)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()
    value = args.steps / (ms * 1e-3)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": W, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": name, "parallelism": "single GPU", "conv_gflop_per_iter": gflop, "launch": launch,
                       "l2": "no explicit flush: the iteration's working set exceeds the 126 MB L2" if args.workload == "cfg3" else
                             "no explicit flush (2-D working set: tens of MB)"},
            "e2e": {"value": args.steps / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": real_h.numel() * 4 + (second_h.numel() * 4 if args.workload == "cfg1" else 0),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": per_iter * args.steps, "clocks": clocks, "roofline": None, "model_tflops": value * gflop / 1e3}
    _emit(line)


_REAL_STDOUT = [None]


def _emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    fd = _REAL_STDOUT[0]
    if fd is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(fd, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="hpvg", choices=["hpvg", "reference"])
    ap.add_argument("--draws", type=int, default=4096, help="noise draws of the generation leg, all ranks together (BASELINE config 4: 4096)")
    ap.add_argument("--gen-batch", type=int, default=32, help="draws per forward of the generation leg; BatchNorm statistics stay per draw")
    ap.add_argument("--gen-streams", type=int, default=2, help="independent draws in flight on separate CUDA streams (generation leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg, the torch-eager GPU leg and the parity leg")
    ap.add_argument("--no-cfg5", action="store_true", help="multi-GPU runs: skip the extra measurement of BASELINE configs[4]")
    ap.add_argument("--profile-one", action="store_true", help="run one iteration between cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--profile-gen", action="store_true", help="run one generation forward between cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--settle-steps", type=int, default=0, help="extra untimed replays of the recorded iteration before timing")
    ap.add_argument("--graph-candidates", type=int, default=1,
                    help="record the iteration this many times and keep the recording that replays fastest (ScaleTrainer.capture)")
    ap.add_argument("--recapture", type=int, default=0, help="diagnostic: re-record the iteration this many times and time each recording")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying the recorded iteration")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5", "cfg1", "cfg3"],
                    help="cfg2 = BASELINE configs[1] (16 x 64 x 64, the metric's configuration); cfg5 = configs[4] (32 x 128 x 128); "
                         "cfg1 = configs[0] (2-D, 128 px); cfg3 = configs[2] (GeneratorSG baseline): value + e2e only")
    args = ap.parse_args()
    WORKLOAD["name"] = args.workload
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its version
    # there when NCCL_DEBUG is VERSION or WARN) are sent to stderr, and only _emit() writes to the real stdout
    sys.stdout.flush()
    _REAL_STDOUT[0] = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("cfg1", "cfg3"):
        run_other(args)
    else:
        run_hpvg(args)


if __name__ == "__main__":
    main()
