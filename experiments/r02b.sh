#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02b
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q -s > ${O}_fullsize.txt 2>&1; tail -25 ${O}_fullsize.txt
timeout 120 python experiments/check_fused_bn.py > ${O}_fused_timing.txt 2>&1; tail -8 ${O}_fused_timing.txt
timeout 300 python -m pytest tests -m gpu -q > ${O}_tests.txt 2>&1; tail -30 ${O}_tests.txt
timeout 200 python bench.py --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02b_bench.json"))
print("%.1f iter/s  %.3f ms  e2e %.1f  gen %.0f frames/s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"]), d["roofline"]["by_kernel_ms_per_step"])
PY
