#!/bin/bash
# tcgen05 head convolution (narrow_tc.cu): parity + timing vs the mma.sync kernel
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zg
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "head or tail or thin or narrow or conv3d" > ${O}_tests1.txt 2>&1; tail -15 ${O}_tests1.txt
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_modules.py -x -q -m gpu > ${O}_tests2.txt 2>&1; tail -15 ${O}_tests2.txt
timeout 200 python experiments/thin_bench.py > ${O}_thin_new.txt 2>&1; cat ${O}_thin_new.txt
HPVG_EXPAND_TC=0 timeout 200 python experiments/thin_bench.py > ${O}_thin_old.txt 2>&1; grep "3->64" ${O}_thin_old.txt
