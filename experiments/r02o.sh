#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02o
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scripts.py > ${O}_tests.txt 2>&1; tail -4 ${O}_tests.txt
for knobs in "" "HPVG_SN_PREFETCH=0" ""; do
  tag=$(echo "$knobs" | tr -c 'A-Za-z0-9\n' '_'); [ -z "$tag" ] && tag=default_$RANDOM
  env $knobs timeout 300 python bench.py --no-cpu-baseline > ${O}_bench_${tag}.json 2> ${O}_bench_${tag}.err
  python - "$knobs" ${O}_bench_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("[%s] %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f  chain %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], json.dumps(d["roofline"].get("dependent_chain", {}))[-160:]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 900 python -m pytest tests/test_gpu_scripts.py -m gpu -q -s > ${O}_scripts.txt 2>&1; tail -12 ${O}_scripts.txt | cut -c1-900
