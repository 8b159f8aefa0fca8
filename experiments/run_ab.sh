#!/bin/bash
# A/B of the round's launch-level changes on one B200 (development aid): tests first, then bench.py with each switch off
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() { name=$1; shift; ( time env "$@" timeout 600 python bench.py --no-cpu-baseline --steps 20 --draws 64 ) > gpurun_out/ab_$name.log 2>&1; echo "$name exit $?"; grep -o '"ms_per_step": [0-9.]*' gpurun_out/ab_$name.log | head -1; grep -o '"us_per_launch": [0-9.]*' gpurun_out/ab_$name.log; grep -o '"generation": {"metric": "generated_frames_per_s", "value": [0-9.]*' gpurun_out/ab_$name.log; }
run all X=1
run nopdl HPVG_PDL=0
run noprefetch HPVG_TC_WPREFETCH=0
run nofuse HPVG_FUSE_MASK=0
run noskip HPVG_SKIP_CRITIC_GRADS=0
timeout 300 python experiments/bench_kernels.py clk > gpurun_out/clk.log 2>&1; tail -12 gpurun_out/clk.log
