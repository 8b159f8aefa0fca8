#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02f
timeout 120 python experiments/bench_kernels.py clk > ${O}_clk.txt 2>&1; cat ${O}_clk.txt | tail -30
timeout 120 python experiments/bench_kernels.py tc > ${O}_tc.txt 2>&1; tail -3 ${O}_tc.txt
BENCH_GRAPHED=1 timeout 120 python experiments/bench_kernels.py tc > ${O}_tc_graphed.txt 2>&1; tail -3 ${O}_tc_graphed.txt
timeout 600 python -m pytest tests/test_gpu_scripts.py -m gpu -q -s -k "train_video_runs or train_image" > ${O}_scripts.txt 2>&1; tail -30 ${O}_scripts.txt | cut -c1-1500
