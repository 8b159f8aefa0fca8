#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zc
L=$PWD/hp-vae-gan_b200/lib
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_layers.py tests/test_gpu_modules.py tests/test_gpu_training.py -x -q -m gpu > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
for v in default s0p0 old default s0p0 old; do
  if [ $v == default ]; then unset HPVG_LIB; else export HPVG_LIB=$L/libhpvg_$v.so; fi
  timeout 300 python bench.py --no-cpu-baseline --draws 1024 > ${O}_bench_$v.json 2> ${O}_bench_$v.err; echo "$v rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("${O}_bench_$v.json"))
    r = d["roofline"]
    print("$v: %.1f iter/s  %.3f ms  gen %.0f | conv_tc %.2f us/launch chain %.2f | %s" % (d["value"], d["ms_per_step"], d["generation"]["value"], r["us_per_launch"], r["dependent_chain"]["us_per_launch"], json.dumps(r["by_kernel_ms_per_step"])))
except Exception as e:
    print("$v: no line", e)
PY
done
