#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02d
timeout 400 python -m pytest tests/test_gpu_training.py tests/test_gpu_data.py tests/test_gpu_modules.py -m gpu -q -s > ${O}_training.txt 2>&1; tail -40 ${O}_training.txt
timeout 900 python -m pytest tests/test_gpu_scripts.py -m gpu -q -s > ${O}_scripts.txt 2>&1; tail -60 ${O}_scripts.txt
timeout 300 python bench.py > ${O}_bench.json 2> ${O}_bench.err; tail -5 ${O}_bench.err; cat ${O}_bench.json | head -c 6000
