#!/bin/bash
# 2-GPU call: gradient buckets in peer memory (csrc/peer.cu): tests, micro-benchmark against NCCL, bench A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r04b
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q -s -x > ${O}_dist.txt 2>&1; tail -6 ${O}_dist.txt | cut -c1-600
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 experiments/peer_bench.py > ${O}_peer_bench.txt 2>&1; grep -v Warning ${O}_peer_bench.txt | tail -8
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 --no-cpu-baseline --no-cfg5 --draws 512 > ${O}_${name}.json 2> ${O}_${name}.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r04b_%s.json" % sys.argv[1]))
    print("%-12s N=2: %.1f iter/s  %.3f ms  e2e %.1f  %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("allreduce")))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
  tail -2 ${O}_${name}.err | cut -c1-300
}
run peer HPVG_X=0
run nccl HPVG_PEER_ALLREDUCE=0
run peer_early HPVG_EARLY_REC_BWD=1
run nccl_early HPVG_PEER_ALLREDUCE=0 HPVG_EARLY_REC_BWD=1
