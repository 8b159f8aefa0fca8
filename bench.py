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
  N > 1 : batched-noise data-parallel training — one clip + its noise per GPU, replicated weights, one flat NCCL
          all-reduce of the gradients per backward (weak scaling: value = N clip-iterations per iteration time).
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

CONV_GFLOP_PER_ITER = 1720.43      # SURVEY.md App. B, config 2 finest level (config 5: 12 650.99)
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
# product arm
# ---------------------------------------------------------------------------------------------------------------
def run_hpvg(args):
    import torch.distributed as dist
    from hpvg import lib, train
    from modules import networks_3d

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

    o = make_opt()
    sg, sd = fresh_states(o)                       # identical on every rank (seed 0): replicated weights
    G = networks_3d.GeneratorHPVAEGAN(o)
    for _ in range(o.scale_idx):
        G.init_next_stage()
    G.load_state_dict(sg)
    D = networks_3d.WDiscriminator3D(o)
    D.load_state_dict(sd)
    G.to(dev)
    D.to(dev)
    use_graph = not args.no_graph
    trainer = train.ScaleTrainer(o, G, D, distributed=distributed, capturable=use_graph)
    real_h, real_zero_h = synthetic_clip(o, rank)  # one clip per rank
    real_h, real_zero_h = real_h.pin_memory(), real_zero_h.pin_memory()
    real, real_zero = real_h.to(dev), real_zero_h.to(dev)
    torch.manual_seed(1 + rank)                    # per-rank noise; the GP alpha comes from the CPU generator below
    cpu_gen_state = torch.random.get_rng_state()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    last = {}
    W = max(3, args.warmup)
    if use_graph:
        # the whole iteration (every libhpvg kernel, autograd glue, clipping, both Adam steps, the all-reduces) is
        # recorded once into a CUDA graph after W eager warm-up iterations and replayed per step
        c0 = lib.launch_count()
        trainer.capture(real, real_zero, warmup=W, candidates=args.graph_candidates)
        torch.cuda.synchronize()
        last["pool_mb"] = getattr(trainer, "pool_bytes", 0) / 2**20
        sys.stderr.write("graph recordings, ms per replay: %s\n" % ", ".join("%.3f" % t for t in getattr(trainer, "capture_trials", [])))
        # launches counted per recorded iteration: W eager warm-up iterations + the recording pass of the first candidate, then
        # 1 warm-up + 1 recording pass per further candidate (replays launch nothing through the library's host side)
        launches_per_iter = (lib.launch_count() - c0) // (W + 1 + 2 * (max(1, args.graph_candidates) - 1))

        def step_resident():
            trainer.replay()

        if args.recapture < 0:
            # diagnostic: per-replay GPU time (events) and host enqueue time of 48 consecutive replays of one recording
            import time as _time
            nrep = -args.recapture if args.recapture < -1 else 48
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(nrep + 1)]
            host = []
            torch.cuda.synchronize()
            evs[0].record()
            for k in range(nrep):
                t0 = _time.perf_counter()
                step_resident()
                host.append((_time.perf_counter() - t0) * 1e3)
                evs[k + 1].record()
            torch.cuda.synchronize()
            sys.stderr.write("per-replay ms (gpu): %s\n" % " ".join("%.2f" % evs[k].elapsed_time(evs[k + 1]) for k in range(nrep)))
            sys.stderr.write("per-replay ms (host enqueue): %s\n" % " ".join("%.2f" % h for h in host))
        if args.recapture > 0:
            # diagnostic: replay time of several recordings of the same iteration in ONE process (how much of the run-to-run
            # spread of ms_per_step is the schedule the graph instantiation happens to pick)
            for k in range(args.recapture + 1):
                for _ in range(3):
                    step_resident()
                sys.stderr.write("recording %d: %.3f ms per replay\n" % (k, timed(step_resident, 10) / 10))
                if k < args.recapture:
                    trainer.capture(real, real_zero, warmup=1)

        def step_e2e():
            out = trainer.replay(real_h, real_zero_h)          # H2D of the clip into the graph's input buffers
            last["rec_loss"] = out["rec_loss"].item()          # device -> host read of the step's result
        # The replay time of this iteration is bimodal on the pool's B200s (5.27 / 5.49 ms at config 2): per-replay CUDA events
        # over 600 consecutive replays showed 125 replays at 5.55 ms, one 7.2 ms hiccup, then 475 at 5.33 ms, with the reported
        # SM / memory clocks unchanged; other runs stayed in one mode for their whole length, and neither extra warm-up nor
        # re-recording the graph selects the mode.  --settle-steps adds untimed replays for experiments with it (default 0).
        settle = args.settle_steps
        for _ in range(2 + settle):
            step_resident()
        last["extra_warmup"] = 2 + settle
    else:
        launches_per_iter = None

        def step_resident():
            trainer.iteration(real, real_zero)

        def step_e2e():
            r = real_h.to(dev, non_blocking=True)
            rz = real_zero_h.to(dev, non_blocking=True)
            out = trainer.iteration(r, rz)
            last["rec_loss"] = out["rec_loss"].item()
        for _ in range(W):
            step_resident()
    if args.profile_gen:
        # for `ncu --profile-from-start off`: one forward of the generation leg (gen-batch draws, one stream) between Start/Stop
        ps = train.Sampler(G, o, dev, batch=max(1, args.gen_batch), graph=use_graph, streams=1, static_weights=True)
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
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        _emit({"profiled": "one iteration", "graph": use_graph})
        return
    torch.cuda.reset_peak_memory_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = lib.launch_count() - n0 if not use_graph else launches_per_iter * args.steps
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    peak_mb = torch.cuda.max_memory_allocated() / 2**20
    # working set of one iteration: the memory pool of the recorded graph (every activation saved for the backward passes,
    # every gradient) when graphed, else the allocator's peak during the timed steps
    work_mb = last.get("pool_mb") or peak_mb

    # roofline leg: the same steps again with CUDA events around every convolution launch (on the launching stream)
    prof_steps = min(args.steps, 3)

    def step_eager():
        # keep the GPU busy while the host enqueues the iteration, so that the event pairs bracket back-to-back kernels
        # and not host launch latency (eager launching is CPU-bound on this workload)
        torch.cuda._sleep(int(0.12 * 1.9e9))
        trainer.iteration(real, real_zero)
    # one kernel at a time: the reconstruction / weight-gradient side streams are joined into the main stream and the
    # programmatic launch overlap is off for this leg, so an event pair brackets exactly one kernel running alone — its
    # duration, not its share of a GPU that another stream's kernel is also using
    saved = (trainer.overlap, trainer._side, trainer._wside)
    trainer.overlap, trainer._side, trainer._wside = False, None, None
    pdl_was = lib.set_pdl(False)
    step_eager()
    torch.cuda.synchronize()
    lib.profile_enable(True)
    ms_prof = timed(step_eager, prof_steps)
    lib.profile_enable(False)
    lib.set_pdl(pdl_was)
    trainer.overlap, trainer._side, trainer._wside = saved
    rows = lib.profile_dump()

    # generation (BASELINE config 4): fresh z per draw through the whole pyramid, batch 1, draws split over ranks
    _, draws_rank = train.draws_for_rank(args.draws, world, rank)
    draws_rank = max(1, draws_rank)
    gen_batch = max(1, min(args.gen_batch, draws_rank))
    gen_calls = max(1, draws_rank // gen_batch)
    draws_rank = gen_calls * gen_batch
    # batch > 1: every draw is normalised with its own BatchNorm statistics (ops.bn_per_sample), i.e. the reference's batch-1 draws
    sampler = train.Sampler(G, o, dev, batch=gen_batch, graph=use_graph, streams=args.gen_streams, static_weights=True)   # G is not trained during this leg
    frames = [0]

    def gen_all():
        n = 0
        sampler.begin()
        for _ in range(gen_calls):
            n += sampler.frames_per_call(sampler.sample())
        sampler.wait()
        frames[0] = n
    gen_all()
    ms_gen = timed(gen_all, 1)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except OSError:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)"
        if not peak_tf:
            peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md: ~1.4 PFLOP/s sustained)"
        by_kind = {}
        for r in rows:
            k = by_kind.setdefault(r["kind"], {"launches": 0, "ms": 0.0, "flops": 0.0})
            k["launches"] += r["launches"]; k["ms"] += r["ms"]; k["flops"] += r["work"] * r["launches"]
        # dominant kernel = the tcgen05 implicit-GEMM convolution (fprop / dgrad form) at the finest level's 64->64 shape
        top = max((r for r in rows if r["kind"] == "conv_tc"), key=lambda r: r["ms"], default=None)
        roofline = None
        if top is not None:
            per_launch_ms = top["ms"] / top["launches"]
            achieved = top["work"] / (per_launch_ms * 1e-3) / 1e12
            roofline = {"bound": "tensor", "kernel": "conv_tc_kernel (tcgen05 implicit-GEMM conv, 64->64 @ 16x64x64)",
                        "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                        # dram__bytes_read.sum + dram__bytes_write.sum of one 64->64 launch at 16x64x64 from the committed ncu --set full
                        # capture (profiles/r01g_conv_tc_ncu.txt): 8.71 MB read (algorithmic 8.39 MB input + 0.22 MB weights), 0 written
                        # before the kernel ends (the 8.39 MB output is still in L2); None for other workloads
                        "traffic": 8705024 if WORKLOAD["name"] == "cfg2" else None, "traffic_unit": "bytes per launch",
                        "flops_per_launch": top["work"], "us_per_launch": per_launch_ms * 1e3, "launches_timed": top["launches"],
                        "peak_source": peak_src,
                        "share_of_step": (top["ms"] / prof_steps) / (ms / args.steps),
                        "by_kernel_ms_per_step": {k: v["ms"] / prof_steps for k, v in by_kind.items()},
                        "by_kernel_tflops": {k: (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else None) for k, v in by_kind.items()},
                        "note": "per-launch CUDA events need eager launches: this leg runs the same iteration un-graphed on ONE stream "
                                "(no concurrent side-stream kernels, no programmatic launch overlap) behind a GPU-side delay so "
                                "that launches are queued ahead of the GPU; share_of_step = kernel ms per iteration / "
                                "graph-replay ms per iteration (the replay overlaps streams, so shares can sum above 1)"}
        value = world * args.steps / (ms * 1e-3)
        e2e = world * args.steps / (ms_e2e * 1e-3)
        bi = (real_h.numel() + real_zero_h.numel()) * 4
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": W + last.get("extra_warmup", 0),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16", "data": "synthetic",
                "config": {"workload": workload_name(o), "parallelism": "dp%d (one clip per GPU, flat NCCL grad all-reduce)" % world if distributed else "single GPU",
                           "l2": "no explicit flush: one iteration writes and re-reads a %.0f MB working set (activations saved for the "
                                 "backward passes, gradients), %s the 126 MB L2" % (work_mb, "above" if work_mb > 126 else "NOT above"),
                           "conv_gflop_per_iter": CONV_GFLOP_PER_ITER if WORKLOAD["name"] == "cfg2" else 12650.99,
                           "warmup": "%d eager iterations + %d replays of the recorded iteration" % (W, last.get("extra_warmup", 0))
                                     if use_graph else "%d eager iterations" % W,
                           "launch": ("one CUDA graph replay per iteration (%d libhpvg kernels recorded)" % launches_per_iter) if use_graph
                           else "eager launches"},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": bi, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "model_tflops": value * (CONV_GFLOP_PER_ITER if WORKLOAD["name"] == "cfg2" else 12650.99) / 1e3 / world,
                "generation": {"metric": "generated_frames_per_s", "value": world * frames[0] / (ms_gen * 1e-3), "unit": "frames/s",
                               "draws": draws_rank * world, "frames_per_draw": level_shape(o, o.scale_idx)[2], "batch": gen_batch,
                               "batchnorm": "per-draw statistics (each draw normalised as in a batch-1 forward)",
                               "ms_per_draw": ms_gen / draws_rank, "streams": sampler.nstreams,
                               "note": "draws split over ranks, no collective; rank 0's share timed x N (equal shares)"}}
        if distributed:
            line["allreduce_bytes_per_step"] = trainer.allreduce_bytes // max(1, trainer.iterations)
        if world == 1 and not args.no_cpu_baseline:
            torch.random.set_rng_state(cpu_gen_state)
            sec, done, cores = cpu_iteration_timer(o, sg, sd, budget_s=30.0, steps=2, warmup=1)
            line["cpu_baseline"] = {"value": 1.0 / sec, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d full iteration(s) of the same workload after 1 warm-up (oracle/train_ref.py, PyTorch-CPU fp32)" % done}
        _emit(line)
    if distributed:
        # all ranks are done once rank 0 has printed; leave without tearing NCCL down: destroy_process_group() was seen
        # to hang while CUDA graphs holding captured all-reduces are alive
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-one", action="store_true", help="run one iteration between cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--profile-gen", action="store_true", help="run one generation forward between cudaProfilerStart/Stop and exit (for ncu)")
    ap.add_argument("--settle-steps", type=int, default=0, help="extra untimed replays of the recorded iteration before timing")
    ap.add_argument("--graph-candidates", type=int, default=1,
                    help="record the iteration this many times and keep the recording that replays fastest (ScaleTrainer.capture)")
    ap.add_argument("--recapture", type=int, default=0, help="diagnostic: re-record the iteration this many times and time each recording")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly instead of replaying the recorded iteration")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 = BASELINE configs[1] (16 x 64 x 64, the metric's configuration); cfg5 = configs[4] (32 x 128 x 128)")
    args = ap.parse_args()
    WORKLOAD["name"] = args.workload
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its version
    # there when NCCL_DEBUG is VERSION or WARN) are sent to stderr, and only _emit() writes to the real stdout
    sys.stdout.flush()
    _REAL_STDOUT[0] = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_hpvg(args)


if __name__ == "__main__":
    main()
