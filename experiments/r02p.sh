#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02p
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scripts.py > ${O}_tests.txt 2>&1; tail -6 ${O}_tests.txt
HPVG_THIN_GS=0 timeout 300 python bench.py --no-cpu-baseline > ${O}_bench_nogs.json 2> ${O}_bench_nogs.err
timeout 300 python bench.py --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err
python - <<'PY'
import json
for tag in ("nogs", ""):
    try:
        d = json.load(open("gpurun_out/r02p_bench%s.json" % ("_" + tag if tag else "")))
        print("[%s] %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f frames/s" % (tag or "thin gs", d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"]), d["roofline"]["by_kernel_ms_per_step"])
    except Exception as e:
        print(tag, "unreadable", e)
PY
timeout 300 python experiments/timeline.py cfg2 > ${O}_timeline_cfg2.txt 2>&1; sed -n 1,40p ${O}_timeline_cfg2.txt | grep -v "^stream"
