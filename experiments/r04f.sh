#!/bin/bash
# N-GPU bench line of the last build, the driver's command line (configs[1] + dp_named_config = configs[4]):  bash experiments/r04f.sh N
set -u
mkdir -p gpurun_out
N=${1:-2}
O=gpurun_out/r04f_bench_${N}gpu
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > ${O}.json 2> ${O}.err; echo "rc=$?"
python - "$O.json" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print("N=%d: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f) %s %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"], d.get("allreduce"), d["config"]["parallelism"]))
print("configs[4]:", json.dumps(d.get("dp_named_config"))[:500])
PY
