#!/bin/bash
set -u
mkdir -p gpurun_out
N=${NGPU:-8}
O=gpurun_out/r02t_${N}gpu
for knobs in "HPVG_FLAT_BUCKET=1" "HPVG_FLAT_BUCKET=0" "HPVG_FLAT_BUCKET=0 HPVG_FUSED_BN=0"; do
  tag=$(echo "$knobs" | tr -c 'A-Za-z0-9\n' '_')
  env $knobs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 30 --warmup 3 --no-cfg5 --draws 256 > ${O}_${tag}.json 2> ${O}_${tag}.err
  python - "$knobs" ${O}_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("[%s] N=%d: %.1f iter/s  %.3f ms  e2e %.1f" % (sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
