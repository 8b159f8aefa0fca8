#!/bin/bash
# ncu evidence (round 2, final build): launch list of one recorded iteration + full capture of the dominant kernels
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zs
python bench.py --no-cpu-baseline --profile-one > ${O}_plain_train.log 2>&1 &&
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ${O}_launches_train.csv \
  python bench.py --no-cpu-baseline --profile-one > ${O}_ncu_train.log 2>&1
echo "launch list rc=$?"; wc -l ${O}_launches_train.csv
python experiments/ncu_targets.py 2 > ${O}_plain_targets.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_kernel|wgrad_tc_kdstack|bn_lrelu_bwd_fused|wgrad_reduce|expand_tc_kernel|narrow_wgrad_tc" -s 16 -c 16 \
  -o ${O}_targets python experiments/ncu_targets.py 2 > ${O}_ncu_targets.log 2>&1
echo "full capture rc=$?"; ls -la ${O}_targets.ncu-rep; tail -2 ${O}_plain_targets.log
