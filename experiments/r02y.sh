#!/bin/bash
# rolled (kh, kw) issue loop of the tcgen05 convolutions: parity, isolated / graphed timings, bench line
set -u
mkdir -p gpurun_out
O=gpurun_out/r02y
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_layers.py -x -q -m gpu > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
timeout 200 python experiments/bench_kernels.py tc 20 > ${O}_tc.txt 2>&1; cat ${O}_tc.txt
BENCH_GRAPHED=1 timeout 200 python experiments/bench_kernels.py tc 15 > ${O}_tc_graphed.txt 2>&1; cat ${O}_tc_graphed.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02y_bench.json"))
print("cfg2: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"]))
r = d["roofline"]
print("roofline: %.1f us/launch frac %.3f; chain %.1f us frac %.3f" % (r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"]))
print(json.dumps(r["by_kernel_ms_per_step"]), json.dumps(r["heaviest_shape_per_kernel"]))
PY
