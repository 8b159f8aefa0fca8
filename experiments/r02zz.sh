#!/bin/bash
# stream priorities: critical chains high, weight gradients default: A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zz
for k in 0 1 0 1; do
HPVG_CHAIN_PRIORITY=$k timeout 300 python bench.py --no-cpu-baseline --draws 512 > ${O}_bench_$k.json 2> ${O}_bench_$k.err; echo -n "chain_priority=$k rc=$? "
python - <<PY
import json
try:
    d = json.load(open("${O}_bench_$k.json"))
    print("%.1f iter/s  %.3f ms  e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print("no line", e)
PY
done
