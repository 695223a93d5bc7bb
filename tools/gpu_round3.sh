#!/bin/bash
# Final single-GPU round: tests, both bench arms, settled launch list, full ncu capture.
TAG=${1:-final}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -3 gpurun_out/pytest_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3650 -c 500 --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --no-cpu-baseline --e2e-steps 2 > gpurun_out/ncu_list_$TAG.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_step -s 3040 -c 2 -o gpurun_out/kstep_$TAG -f \
  python bench.py --settle 3000 --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/ | grep $TAG
