#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_warpnb.json 2> gpurun_out/bench_warpnb.err
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/bench_warpnb.json") if l.startswith("{")][-1]); r=j["roofline"]
print("value %.4g"%j["value"], "ms/step %.4f"%j["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms"], "rebuild ms %.2f/%d"%(r["rebuild_ms_total"], r["rebuilds"]))
PY
python bench.py --no-cpu-baseline --e2e-steps 3 --workload periodic_box --n-per-gpu 1000000 --steps 200 > gpurun_out/bench_warpnb_p.json 2>/dev/null
python - <<PY
import json
j=json.loads([l for l in open("gpurun_out/bench_warpnb_p.json") if l.startswith("{")][-1]); r=j["roofline"]
print("periodic value %.4g"%j["value"], "ms/step %.4f"%j["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms"], "rebuild ms %.2f/%d"%(r["rebuild_ms_total"], r["rebuilds"]))
PY
