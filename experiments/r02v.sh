#!/bin/bash
# release check on one GPU: what the driver runs at round end (tests, smoke, bench both arms) + the configs[4] single-GPU line
set -u
mkdir -p gpurun_out
O=gpurun_out/r02v
timeout 1200 python -m pytest tests -x -q -m gpu > ${O}_tests.txt 2>&1; tail -5 ${O}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; tail -2 ${O}_smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; cat ${O}_bench_reference.json | cut -c1-400
timeout 600 python bench.py --workload cfg5 --no-cpu-baseline > ${O}_bench_cfg5.json 2> ${O}_bench_cfg5.err; tail -2 ${O}_bench_cfg5.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02v_bench.json"))
print("cfg2: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"]))
r = d["roofline"]
print("roofline: %.1f us/launch frac %.3f; chain %.1f us frac %.3f; traffic %s" % (r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"], r["traffic"]))
print("cpu", d["cpu_baseline"]["value"], "eager gpu", d["gpu_eager_baseline"]["value"], "parity", d["parity"]["rel_err"], "clocks", d["clocks"])
c = json.load(open("gpurun_out/r02v_bench_cfg5.json"))
print("cfg5: %.1f iter/s  %.3f ms  e2e %.1f  model_tflops %.0f roofline frac %.3f (%.1f us)" % (c["value"], c["ms_per_step"], c["e2e"]["value"], c["model_tflops"], c["roofline"]["frac"], c["roofline"]["us_per_launch"]))
PY
