#!/bin/bash
# rolled issue loop (wide kernels) + mask prefetch: full GPU suite, generation stress, bench lines
set -u
mkdir -p gpurun_out
O=gpurun_out/r02zf
timeout 1500 python -m pytest tests -x -q -m gpu > ${O}_tests.txt 2>&1; tail -3 ${O}_tests.txt
for i in 1 2; do timeout 300 python experiments/gen_stress.py 40 2 2>&1 | grep gen_stress; done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; echo "rc=$?"; tail -2 ${O}_bench.err | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > ${O}_bench2.json 2> ${O}_bench2.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r02zf_bench.json", "gpurun_out/r02zf_bench2.json"):
    try:
        d = json.load(open(f))
        r = d["roofline"]
        print("%s: %.1f iter/s  %.3f ms  e2e %.1f gen %.0f | conv_tc %.2f us/launch (%.3f) chain %.2f (%.3f) | %s" % (f[-12:], d["value"], d["ms_per_step"], d["e2e"]["value"], d["generation"]["value"], r["us_per_launch"], r["frac"], r["dependent_chain"]["us_per_launch"], r["dependent_chain"]["frac_of_peak"], json.dumps(r["by_kernel_ms_per_step"])))
        if "parity" in d: print(d["parity"]["rel_err"], d.get("gpu_eager_baseline", {}).get("value"), d["cpu_baseline"]["value"])
    except Exception as e:
        print(f, "no line", e)
PY
