#!/bin/bash
# round 2 session 4: main (full-list k_step) on the disordered workloads: parity at the tightened
# JKR/linear bar, drum / 1M / 64M periodic timings, ncu --set full capture of k_step on the drum
mkdir -p gpurun_out
S=${1:-s4}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$S.log
timeout 300 python bench.py --workload drum --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_drum_$S.json 2> gpurun_out/bench_drum_$S.err; echo "drum rc=$?"
timeout 300 python bench.py --particles 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_per1M_$S.json 2> gpurun_out/bench_per1M_$S.err; echo "per1M rc=$?"
if [ "$2" != "no64" ]; then
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_per64M_$S.json 2> gpurun_out/bench_per64M_$S.err; echo "per64M rc=$?"
tail -5 gpurun_out/bench_per64M_$S.err
fi
for f in gpurun_out/bench_*_$S.json; do echo $f; python tools/bench_line.py "$f"; done
if [ "$3" != "noncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip 3200 --launch-count 1 -f -o gpurun_out/kstep_$S \
  python bench.py --workload drum --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$S.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$S.log
ls -la gpurun_out/kstep_$S.ncu-rep
fi
