#!/bin/bash
# 8-GPU call: gradient buckets in peer memory at the full node: micro-benchmark against NCCL, bench (driver's command line) A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r04c
N=${1:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 experiments/peer_bench.py > ${O}_peer_bench.txt 2>&1; grep -v Warning ${O}_peer_bench.txt | tail -6
run() {  # name, extra bench args, env...
  local name=$1; shift
  local extra=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 30 --warmup 3 $extra > ${O}_${name}.json 2> ${O}_${name}.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r04c_%s.json" % sys.argv[1]))
    print("%-12s N=%d: %.1f iter/s  %.3f ms  e2e %.1f  %s  cfg5 %s" % (sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("allreduce"),
          (d.get("dp_named_config") or {}).get("ms_per_step")))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
  grep -v -i "warn\|graph recordings\|^\*\|OMP_NUM" ${O}_${name}.err | tail -3 | cut -c1-300
}
run default "--draws 512" HPVG_X=0
run nccl "--draws 512 --no-cfg5" HPVG_PEER_ALLREDUCE=0
run peer_late "--draws 512 --no-cfg5" HPVG_EARLY_REC_BWD=0
run nccl_late "--draws 512 --no-cfg5" HPVG_PEER_ALLREDUCE=0 HPVG_EARLY_REC_BWD=0
