#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02j
for knobs in "" "HPVG_FUSED_COOP=0" "HPVG_FUSED_COOP=0 HPVG_FUSED_BN_BWD=0" "HPVG_FUSED_BN=0 HPVG_FUSED_BN_BWD=0" ""; do
  tag=$(echo "$knobs" | tr -c 'A-Za-z0-9\n' '_'); [ -z "$tag" ] && tag=default_$RANDOM
  env $knobs timeout 300 python bench.py --no-cpu-baseline > ${O}_bench_${tag}.json 2> ${O}_bench_${tag}.err
  python - "$knobs" ${O}_bench_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("[%s] %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 300 python -m pytest tests/test_gpu_training.py tests/test_gpu_modules.py -m gpu -q -x > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
