#!/bin/bash
# early issue of the first unit's loads + CTA exit after the staging reads: parity, stress, bench
set -u
mkdir -p gpurun_out
O=gpurun_out/r03a
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
timeout 300 python experiments/gen_stress.py 40 2 2>&1 | grep gen_stress
for k in a b c; do
timeout 300 python bench.py --no-cpu-baseline --draws 512 > ${O}_bench_$k.json 2> ${O}_bench_$k.err; echo -n "$k rc=$? "
python - <<PY
import json
try:
    d = json.load(open("${O}_bench_$k.json"))
    r = d["roofline"]
    print("%.1f iter/s  %.3f ms  e2e %.1f | conv_tc %.2f us (%.3f) chain %.2f (%.3f)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"]))
except Exception as e:
    print("no line", e)
PY
done
