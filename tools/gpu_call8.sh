#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s8c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s8c.log
tail -3 gpurun_out/pytest_s8c.log
bash tools/gpu_variants.sh 3000 2>&1 | grep -v "^$"
for wl in hopper box_packing; do
  python bench.py --workload $wl --steps 300 --warmup 10 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_${wl}_s8c.json 2> gpurun_out/bench_${wl}_s8c.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_${wl}_s8c.json")); r=j["roofline"]
    print("$wl", j["config"]["particles"], "value %.4g"%j["value"], "ms/step %.4f"%j["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], "C %.2f T %.2f"%(r["C_half"], r["T_half"]), "rebuilds", j["config"]["rebuilds_in_timed_region"])
except Exception as e:
    print("$wl FAILED", e)
PY
done
