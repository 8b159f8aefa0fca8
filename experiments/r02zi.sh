#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zi
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "head or tail or thin or narrow or conv3d" > ${O}_tests1.txt 2>&1; tail -12 ${O}_tests1.txt
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_modules.py -x -q -m gpu > ${O}_tests2.txt 2>&1; tail -12 ${O}_tests2.txt
timeout 200 python experiments/thin_bench.py > ${O}_thin_new.txt 2>&1; cat ${O}_thin_new.txt
timeout 200 python experiments/head_prof.py 2>&1 | grep " us " | tail -3
