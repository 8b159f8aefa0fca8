#!/bin/bash
# sanitizer evidence (one tool per gpurun call): memcheck over the per-layer kernel tests
set -u
mkdir -p gpurun_out
O=gpurun_out/r02n
timeout 120 python -m pytest tests/test_gpu_layers.py -m gpu -q -x -k "convblock3d_layer or tail_conv or resize or vae_head or gradient_penalty_reduction" > ${O}_plain.txt 2>&1; tail -2 ${O}_plain.txt
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 77 --log-file ${O}_memcheck.log python -m pytest tests/test_gpu_layers.py -m gpu -q -x -k "convblock3d_layer or tail_conv or resize or vae_head or gradient_penalty_reduction" > ${O}_memcheck_pytest.txt 2>&1
echo "memcheck rc=$?"; tail -3 ${O}_memcheck_pytest.txt; tail -5 ${O}_memcheck.log
