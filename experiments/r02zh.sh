#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zh
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "head or tail or thin or narrow or conv3d" > ${O}_tests1.txt 2>&1; tail -3 ${O}_tests1.txt
timeout 200 python experiments/thin_bench.py 2>&1 | grep "3->64"
timeout 200 python experiments/head_prof.py 2>&1 | grep " us " 
HPVG_EXPAND_TC=0 timeout 200 python experiments/head_prof.py 2>&1 | grep " us " | tail -6
