#!/bin/bash
# release check on one GPU (final build): what the driver runs at round end (tests, smoke, bench both arms) + cfg5 / cfg1 / cfg3 lines
set -u
mkdir -p gpurun_out
O=gpurun_out/r03c
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_tests.txt 2>&1; tail -4 ${O}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; tail -2 ${O}_smoke.txt
timeout 300 python experiments/gen_stress.py 40 2 2>&1 | grep gen_stress
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -2 ${O}_bench.err | cut -c1-200
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; cut -c1-300 ${O}_bench_reference.json
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > ${O}_bench_cfg5.json 2> ${O}_bench_cfg5.err; echo "cfg5 rc=$?"
timeout 600 python bench.py --workload cfg1 --no-cpu-baseline > ${O}_bench_cfg1.json 2> ${O}_bench_cfg1.err; echo "cfg1 rc=$?"
timeout 600 python bench.py --workload cfg3 --no-cpu-baseline > ${O}_bench_cfg3.json 2> ${O}_bench_cfg3.err; echo "cfg3 rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r03c_bench.json"))
print("cfg2: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f) model_tflops %.0f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"], d["model_tflops"], d["gpu_launches"]))
r = d["roofline"]
print("roofline: %.1f us/launch frac %.3f; chain %.1f us frac %.3f; traffic %s share %.3f" % (r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"], r["traffic"], r["share_of_step"]))
print("cpu", d["cpu_baseline"]["value"], "eager gpu", d["gpu_eager_baseline"]["value"], "parity", d["parity"]["rel_err"], "clocks", d["clocks"])
for k in ("cfg5", "cfg1", "cfg3"):
    c = json.load(open("gpurun_out/r03c_bench_%s.json" % k))
    rr = c.get("roofline") or {}
    print("%s: %.1f iter/s  %.3f ms  e2e %.1f  model_tflops %.0f roofline %s (%s us)" % (k, c["value"], c["ms_per_step"], c["e2e"]["value"], c["model_tflops"], rr.get("frac"), rr.get("us_per_launch")))
PY
