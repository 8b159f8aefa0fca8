#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python experiments/bn_stats_diag.py > gpurun_out/r02r_bn_diag.txt 2>&1; tail -5 gpurun_out/r02r_bn_diag.txt | cut -c1-700
timeout 300 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -s -k convblock3d_layer > gpurun_out/r02r_layers.txt 2>&1; grep -i "ConvBlock3D\|passed\|failed" gpurun_out/r02r_layers.txt | head
