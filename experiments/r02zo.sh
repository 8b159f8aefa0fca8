#!/bin/bash
# tcgen05 narrow kernels (8 + 8 warps, tile iterator) + 32-bit unit decode: full suite, stress, bench, cfg5 kernel timings
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zo
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
timeout 300 python experiments/gen_stress.py 40 2 2>&1 | grep gen_stress
for v in a b; do
timeout 300 python bench.py --no-cpu-baseline > ${O}_bench_$v.json 2> ${O}_bench_$v.err; echo "rc=$?"
done
python - <<'PY'
import json
for v in ("a", "b"):
    try:
        d = json.load(open("gpurun_out/r02zo_bench_%s.json" % v))
        r = d["roofline"]
        print("%s: %.1f iter/s  %.3f ms  e2e %.1f gen %.0f | conv_tc %.2f us/launch (%.3f) chain %.2f (%.3f) | %s" % (v, d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"], json.dumps(r["by_kernel_ms_per_step"])))
    except Exception as e:
        print(v, "no line", e)
PY
BENCH_BIG=1 timeout 300 python experiments/bench_kernels.py narrow 10 2>&1 | tail -8
BENCH_BIG=1 HPVG_EXPAND_TC=0 HPVG_NARROW_WGRAD_TC=0 timeout 300 python experiments/bench_kernels.py narrow 10 2>&1 | tail -8
