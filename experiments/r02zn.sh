#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zn
timeout 900 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_layers.py tests/test_gpu_modules.py -x -q -m gpu > ${O}_tests1.txt 2>&1; tail -3 ${O}_tests1.txt
python experiments/head_clk.py 2>&1 | tail -5
timeout 200 python experiments/head_prof.py 2>&1 | grep " us " | tail -12
