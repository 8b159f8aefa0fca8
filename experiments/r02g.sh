#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02g
timeout 120 python experiments/coop_concurrency.py 300 > ${O}_coop1.txt 2>&1; echo "coop=1 rc=$?"; tail -4 ${O}_coop1.txt
HPVG_FUSED_COOP=0 timeout 120 python experiments/coop_concurrency.py 300 > ${O}_coop0.txt 2>&1; echo "coop=0 rc=$?"; tail -4 ${O}_coop0.txt
nvidia-smi --query-gpu=name,memory.used --format=csv
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_scripts.py > ${O}_tests.txt 2>&1; tail -5 ${O}_tests.txt
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err; head -c 7000 ${O}_bench.json
