#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02x
timeout 300 python experiments/gen_sweep.py > ${O}_gen_sweep.txt 2>&1; tail -7 ${O}_gen_sweep.txt
for i in 1 2 3 4; do
  timeout 600 python -m pytest tests/test_gpu_training.py tests/test_gpu_modules.py tests/test_gpu_layers.py -q -m gpu > ${O}_tests_$i.txt 2>&1; tail -2 ${O}_tests_$i.txt | cut -c1-300
done
