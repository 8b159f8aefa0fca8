#!/bin/bash
# 2-GPU call (final build): DataParallel script test, distributed gradient test, bench at N=2 (cfg2 + configs[4])
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zu
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests/test_gpu_dist.py -m gpu -q -s > ${O}_dist.txt 2>&1; tail -4 ${O}_dist.txt | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_scripts.py -m gpu -q -s -k dataparallel > ${O}_dp.txt 2>&1; tail -4 ${O}_dp.txt | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 3 > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err; tail -3 ${O}_bench_2gpu.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02zu_bench_2gpu.json"))
print("N=2: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f) allreduce %s B/step" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"], d.get("allreduce_bytes_per_step")))
print("configs[4]:", json.dumps(d.get("dp_named_config"))[:600])
PY
