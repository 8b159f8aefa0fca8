#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out/r02e
timeout 300 python experiments/parity_diag.py > ${O}_parity_diag.txt 2>&1; tail -12 ${O}_parity_diag.txt
timeout 300 python experiments/timeline.py cfg2 > ${O}_timeline_cfg2.txt 2>&1; head -70 ${O}_timeline_cfg2.txt
timeout 300 python bench.py --no-cpu-baseline > ${O}_bench.json 2> ${O}_bench.err; tail -3 ${O}_bench.err; head -c 5000 ${O}_bench.json
