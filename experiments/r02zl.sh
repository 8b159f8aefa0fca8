#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zl
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_layers.py -x -q -m gpu > ${O}_tests1.txt 2>&1; tail -3 ${O}_tests1.txt
python experiments/head_clk.py 2>&1 | tail -5
timeout 200 python experiments/head_prof.py 2>&1 | grep " us " | awk '{print $1, $3, $4}' | sort -k2 | uniq -c -f1 | head -20
timeout 200 python experiments/bench_kernels.py tc 20 2>&1 | tail -2
BENCH_GRAPHED=1 timeout 200 python experiments/bench_kernels.py tc 15 2>&1 | tail -2
