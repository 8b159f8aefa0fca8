#!/bin/bash
set -u
mkdir -p gpurun_out
N=${NGPU:-4}
O=gpurun_out/r02u_${N}gpu
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; tail -2 ${O}_bench.err
python - ${O}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("N=%d: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f)  allreduce %s B/step" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"], d.get("allreduce_bytes_per_step")))
print("configs[4]:", json.dumps(d.get("dp_named_config"))[:400])
PY
for knobs in "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=2 NCCL_MIN_CTAS=1" "HPVG_FUSED_COOP=0"; do
  tag=$(echo "$knobs" | tr -c 'A-Za-z0-9\n' '_')
  env $knobs timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 30 --warmup 3 --no-cfg5 --draws 256 > ${O}_${tag}.json 2> ${O}_${tag}.err
  python - "$knobs" ${O}_${tag}.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    print("[%s] N=%d: %.1f iter/s  %.3f ms  e2e %.1f" % (sys.argv[1], d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
