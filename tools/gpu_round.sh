#!/bin/bash
# Standard GPU round: parity tests, bench, ncu full capture of k_step, launch list.
# usage (on the GPU box, from the repo root): bash tools/gpu_round.sh <tag> [settle]
TAG=${1:-run}
SETTLE=${2:-3000}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
python bench.py --steps 300 --warmup 20 --settle $SETTLE > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench rc=$?" >> gpurun_out/bench_$TAG.err
SKIP=$((SETTLE + 10))
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s $SKIP -c 1 -o gpurun_out/kstep_$TAG -f \
  python bench.py --settle $SETTLE --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/pytest_$TAG.log
cat gpurun_out/bench_$TAG.json
