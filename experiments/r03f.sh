#!/bin/bash
set -u
echo -n "== early loads in an elect.sync region, all instantiations : "; HPVG_LIB=$PWD/hp-vae-gan_b200/lib/libhpvg_el15.so timeout 200 python bench.py --no-cpu-baseline --draws 512 --steps 20 2> gpurun_out/r03f.err > gpurun_out/r03f.json; grep -c "illegal memory" gpurun_out/r03f.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r03f.json")); r = d["roofline"]
    print("%.1f iter/s  %.3f ms | conv_tc %.2f us (%.3f) chain %.2f (%.3f)" % (d["value"], d["ms_per_step"], r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"]))
except Exception as e:
    print("no line", e)
PY
