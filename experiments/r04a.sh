#!/bin/bash
# 2-GPU A/B: does the early backward of the reconstruction path (HPVG_EARLY_REC_BWD=1) fill the bubble of the critic's all-reduce?
set -u
mkdir -p gpurun_out
O=gpurun_out/r04a
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 --no-cpu-baseline --no-cfg5 --draws 512 > ${O}_${name}.json 2> ${O}_${name}.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r04a_%s.json" % sys.argv[1]))
    print("%-12s N=2: %.1f iter/s  %.3f ms  e2e %.1f" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
}
run default HPVG_X=0
run early HPVG_EARLY_REC_BWD=1
run default2 HPVG_X=0
run early2 HPVG_EARLY_REC_BWD=1
