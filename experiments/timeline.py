"""Kernel timeline of one replayed iteration (CUPTI through torch.profiler): per-stream busy time, idle gaps, top kernels.
   python experiments/timeline.py [cfg2|cfg5] > gpurun_out/timeline.txt   (also writes gpurun_out/timeline_<cfg>.json)"""
import json
import os
import sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [os.path.join(ROOT, "hp-vae-gan_b200"), ROOT]
import torch
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"


class A:
    no_graph = False; warmup = 3; graph_candidates = 1; settle_steps = 0; steps = 5


dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
leg = bench.TrainLeg(name, A, 0, 1, dev, False)
leg.prepare()
for _ in range(3):
    leg.step_resident()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    leg.step_resident()
    torch.cuda.synchronize()
    leg.step_resident()
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
        ev.append((e.time_range.start, e.time_range.end, e.name, getattr(e, 'device_resource_id', getattr(e, 'thread', 0))))
ev.sort()
# the second replay: events after the largest gap in the middle
mid = (ev[0][0] + ev[-1][1]) / 2
second = [x for x in ev if x[0] >= mid]
first_start = second[0][0]
t0, t1 = second[0][0], max(x[1] for x in second)
print("replay span: %.1f us, %d kernels" % (t1 - t0, len(second)))
streams = {}
for s, e, n, st in second:
    streams.setdefault(st, []).append((s, e, n))
for st, xs in sorted(streams.items()):
    print("stream %s: %d kernels, busy %.1f us, first %.1f last %.1f" % (st, len(xs), sum(e - s for s, e, _ in xs), xs[0][0] - t0, xs[-1][1] - t0))
# union busy time / idle gaps
ivs = sorted((s, e) for s, e, _, _ in second)
busy, cur_s, cur_e, gaps = 0.0, ivs[0][0], ivs[0][1], []
for s, e in ivs[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e - t0))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("union busy %.1f us, idle %.1f us in %d gaps (largest: %s)" % (busy, (t1 - t0) - busy, len(gaps), ", ".join("%.1f@%.0f" % g for g in sorted(gaps, reverse=True)[:8])))
agg = {}
for s, e, n, st in second:
    k = n.split('(')[0][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += e - s
tot = sum(v[1] for v in agg.values())
print("sum of kernel durations %.1f us" % tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%8.1f us %5d x %6.2f us  %s" % (v[1], v[0], v[1] / v[0], k))
with open(os.path.join(ROOT, "gpurun_out", "timeline_%s.json" % name), "w") as f:
    json.dump([dict(start=s - t0, dur=e - s, name=n[:120], stream=st) for s, e, n, st in second], f)
