#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02m
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scripts.py > ${O}_tests.txt 2>&1; tail -4 ${O}_tests.txt
timeout 200 python experiments/seq_bench.py > ${O}_seq.txt 2>&1; cat ${O}_seq.txt | tail -9
timeout 300 python bench.py --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02m_bench.json"))
print("%.1f iter/s  %.3f ms  e2e %.1f  gen %.0f frames/s (e2e %.0f)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"]), d["roofline"]["by_kernel_ms_per_step"])
PY
timeout 300 python experiments/timeline.py cfg2 > ${O}_timeline_cfg2.txt 2>&1; sed -n 1,45p ${O}_timeline_cfg2.txt | grep -v "^stream"
