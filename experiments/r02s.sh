#!/bin/bash
# N-GPU bench: configs[1] + configs[4] (dp_named_config); NGPU from the environment
set -u
mkdir -p gpurun_out
N=${NGPU:-8}
O=gpurun_out/r02zw_${N}gpu
nvidia-smi --query-gpu=index,name --format=csv | head -10
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err
python - ${O}_bench.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("N=%d: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f)  allreduce %s B/step" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"], d.get("allreduce_bytes_per_step")))
print("configs[4]:", json.dumps(d.get("dp_named_config"))[:700])
PY
