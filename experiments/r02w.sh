#!/bin/bash
# stability: the full GPU suite three times; then the configs[0] / configs[2] bench lines
set -u
mkdir -p gpurun_out
O=gpurun_out/r02w
for i in 1 2 3; do
  timeout 1200 python -m pytest tests -q -m gpu > ${O}_tests_$i.txt 2>&1; tail -3 ${O}_tests_$i.txt | cut -c1-300
done
timeout 300 python bench.py --workload cfg1 > ${O}_bench_cfg1.json 2> ${O}_bench_cfg1.err; tail -2 ${O}_bench_cfg1.err; cut -c1-700 ${O}_bench_cfg1.json
timeout 300 python bench.py --workload cfg3 > ${O}_bench_cfg3.json 2> ${O}_bench_cfg3.err; tail -2 ${O}_bench_cfg3.err; cut -c1-700 ${O}_bench_cfg3.json
