#!/bin/bash
# release check on one GPU (final build of the round): what the driver runs at round end (tests, smoke, bench both arms), then the ncu
# evidence of the same build: launch list of one recorded iteration + full capture of the dominant kernels
set -u
mkdir -p gpurun_out
O=gpurun_out/r04e
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_tests.txt 2>&1; tail -4 ${O}_tests.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.txt 2>&1; tail -2 ${O}_smoke.txt
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "bench rc=$?"; tail -2 ${O}_bench.err | cut -c1-200
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; cut -c1-300 ${O}_bench_reference.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r04e_bench.json"))
print("cfg2: %.1f iter/s  %.3f ms  e2e %.1f  gen %.0f (e2e %.0f) model_tflops %.0f launches %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], d["generation"]["e2e"]["value"], d["model_tflops"], d["gpu_launches"]))
r = d["roofline"]
print("roofline: %.1f us/launch frac %.3f; chain %.1f us frac %.3f; traffic %s share %.3f" % (r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"], r["traffic"], r["share_of_step"]))
print("cpu", d["cpu_baseline"]["value"], "eager gpu", d["gpu_eager_baseline"]["value"], "parity", d["parity"]["rel_err"], "clocks", d["clocks"])
PY
python bench.py --no-cpu-baseline --profile-one > ${O}_plain_train.log 2>&1 &&
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_train.csv \
  python bench.py --no-cpu-baseline --profile-one > ${O}_ncu_train.log 2>&1
echo "launch list rc=$?"; wc -l ${O}_launches_train.csv
python experiments/ncu_targets.py 2 > ${O}_plain_targets.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kdstack|bn_lrelu_bwd_fused|wgrad_reduce|expand_tc_kernel|narrow_wgrad_tc" -s 16 -c 16 \
  -o ${O}_targets python experiments/ncu_targets.py 2 > ${O}_ncu_targets.log 2>&1
echo "full capture rc=$?"; ls -la ${O}_targets.ncu-rep; tail -2 ${O}_plain_targets.log
