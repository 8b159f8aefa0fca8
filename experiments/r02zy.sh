#!/bin/bash
# early backward of the reconstruction path: A/B + training parity tests; new narrow tests
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zy
timeout 900 python -m pytest tests/test_gpu_narrow.py tests/test_gpu_training.py tests/test_gpu_modules.py -x -q -m gpu > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
for k in 0 1 0 1; do
HPVG_EARLY_REC_BWD=$k timeout 300 python bench.py --no-cpu-baseline --draws 512 > ${O}_bench_$k.json 2> ${O}_bench_$k.err; echo -n "early_rec_bwd=$k rc=$? "
python - <<PY
import json
try:
    d = json.load(open("${O}_bench_$k.json"))
    print("%.1f iter/s  %.3f ms  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("no line", e)
PY
done
