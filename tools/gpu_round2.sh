#!/bin/bash
# Single-GPU round: parity tests, bench (both arms), kernel variants, ncu launch list + full capture.
# usage (on the GPU box, from the repo root): bash tools/gpu_round2.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
cat gpurun_out/bench_ref_$TAG.json
for wl in hopper cohesive_jkr cohesive_dmt periodic_box box_packing; do
  python bench.py --workload $wl --steps 200 --warmup 10 --settle 1000 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_${wl}_$TAG.json 2> gpurun_out/bench_${wl}_$TAG.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/bench_${wl}_$TAG.json")); r=j["roofline"]
    print("$wl", j["config"]["particles"], "value %.4g"%j["value"], "ms/step %.4f"%j["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms"], "frac %.3f"%r["frac"], "C %.2f T %.2f"%(r["C_half"], r["T_half"]))
except Exception as e:
    print("$wl FAILED", e)
PY
done
bash tools/gpu_variants.sh 3000 2>&1 | grep -v "^$"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 250 -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --settle 300 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 3010 -c 2 -o gpurun_out/kstep_$TAG -f \
  python bench.py --settle 3000 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/ | tail -20
