#!/bin/bash
mkdir -p gpurun_out
for mode in 0 1 2; do
  LETHE_DEM_L2_PERSIST=$mode python bench.py --no-cpu-baseline --e2e-steps 2 --steps 300 > gpurun_out/bench_l2p$mode.json 2>/dev/null
  python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/bench_l2p$mode.json") if l.startswith("{")][-1]); r=j["roofline"]
print("L2_PERSIST=$mode value %.4g"%j["value"], "ms/step %.4f"%j["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms"])
PY
done
