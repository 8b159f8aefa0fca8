#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r03b
L=$PWD/hp-vae-gan_b200/lib
for v in e0r1 e1r0 e2r0; do
HPVG_LIB=$L/libhpvg_$v.so timeout 300 python bench.py --no-cpu-baseline --draws 512 > ${O}_bench_$v.json 2> ${O}_bench_$v.err; echo -n "$v rc=$? "
python - <<PY
import json
try:
    d = json.load(open("${O}_bench_$v.json"))
    r = d["roofline"]
    print("%.1f iter/s  %.3f ms  e2e %.1f | conv_tc %.2f us (%.3f) chain %.2f (%.3f)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"]))
except Exception as e:
    print("no line", e)
PY
done
