#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02q
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scripts.py > ${O}_tests.txt 2>&1; tail -4 ${O}_tests.txt
HPVG_THIN_GS=0 timeout 200 python experiments/thin_bench.py > ${O}_thin_old.txt 2>&1; tail -6 ${O}_thin_old.txt
timeout 200 python experiments/thin_bench.py > ${O}_thin_new.txt 2>&1; tail -6 ${O}_thin_new.txt
timeout 300 python bench.py --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02q_bench.json"))
print("%.1f iter/s  %.3f ms  e2e %.1f  gen %.0f frames/s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"]), d["roofline"]["by_kernel_ms_per_step"])
PY
