#!/bin/bash
# 2-GPU call: the peer-memory kernel gathers / scatters the gradient tensors itself (no pack / unpack launches): tests, bench A/B
set -u
mkdir -p gpurun_out
O=gpurun_out/r04d
timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -q -s -x > ${O}_dist.txt 2>&1; tail -6 ${O}_dist.txt | cut -c1-900
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 40 --warmup 3 --no-cpu-baseline --no-cfg5 --draws 512 > ${O}_${name}.json 2> ${O}_${name}.err
  python - "$name" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/r04d_%s.json" % sys.argv[1]))
    print("%-12s N=2: %.1f iter/s  %.3f ms  e2e %.1f  %s launches %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("allreduce"), d["config"]["launch"]))
except Exception as e:
    print(sys.argv[1], "no line", e)
PY
  grep -v -i "warn\|graph recordings\|^\*\|OMP_NUM\|NCCL version" ${O}_${name}.err | tail -3 | cut -c1-300
}
run gather HPVG_X=0
run packed HPVG_PEER_FUSED_PACK=0
run gather2 HPVG_X=0
