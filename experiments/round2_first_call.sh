#!/bin/bash
# First GPU call of the next round: everything that was written after round 1's GPU budget ran out, measured in one go.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash experiments/round2_first_call.sh'
# Results land in gpurun_out/r02a_*.  Every step runs under its own timeout; a failing step does not stop the others.
set -u
mkdir -p gpurun_out
O=gpurun_out/r02a
# 1. default build still green (tests + smoke)
timeout 120 python -m pytest tests -m gpu -x -q > ${O}_tests.txt 2>&1; tail -2 ${O}_tests.txt
# 2. the kd-stacked weight gradient: parity against the CUDA-core kernel and the measured default, then timing
timeout 90 python experiments/check_wgrad_stack.py > ${O}_wgrad_stack.txt 2>&1; tail -6 ${O}_wgrad_stack.txt
# 3. bench lines: default, maximum shared-memory carve-out, stacked weight gradient (only meaningful if step 2 said ALL OK)
timeout 200 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
HPVG_CARVEOUT=1 timeout 200 python bench.py --no-cpu-baseline > ${O}_bench_carveout.json 2> ${O}_bench_carveout.err
# critic / encoder weight gradients on the side stream (hpvg.ops.deferred_weight): parity tests first, then the bench line
HPVG_CRITIC_WSIDE=1 timeout 150 python -m pytest tests/test_gpu_modules.py tests/test_gpu_training.py -m gpu -q > ${O}_tests_critic_wside.txt 2>&1
tail -2 ${O}_tests_critic_wside.txt
if tail -1 ${O}_tests_critic_wside.txt | grep -q "passed" && ! tail -1 ${O}_tests_critic_wside.txt | grep -q "failed"; then
  HPVG_CRITIC_WSIDE=1 timeout 200 python bench.py --no-cpu-baseline > ${O}_bench_critic_wside.json 2> ${O}_bench_critic_wside.err
fi
if grep -q "ALL OK" ${O}_wgrad_stack.txt; then
  HPVG_WGRAD_STACK=2 timeout 200 python bench.py --no-cpu-baseline > ${O}_bench_wgrad_staged.json 2> ${O}_bench_wgrad_staged.err
  HPVG_WGRAD_STACK=1 timeout 200 python bench.py --no-cpu-baseline > ${O}_bench_wgrad_stack.json 2> ${O}_bench_wgrad_stack.err
  HPVG_WGRAD_STACK=1 timeout 120 python -m pytest tests/test_gpu_layers.py tests/test_gpu_training.py -m gpu -q > ${O}_tests_wgrad_stack.txt 2>&1
  tail -2 ${O}_tests_wgrad_stack.txt
fi
for f in ${O}_bench_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], "%.1f iter/s  %.3f ms  e2e %.1f  gen %.0f frames/s  conv_tc frac %.3f" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["roofline"]["frac"]),
        {k: round(v, 3) for k, v in d["roofline"]["by_kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
# 4. launch list of the final build (serialised, cold cache: shares only)
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_train.csv \
  python bench.py --profile-one > ${O}_ncu.log 2>&1
