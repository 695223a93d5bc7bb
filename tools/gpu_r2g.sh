#!/bin/bash
# round 2 session 23: closing verification on one GPU — full GPU suite, default bench (64 M) + reference arm, ncu launch list of the
# default command, ncu --set full of one k_step launch on the 64 M box
mkdir -p gpurun_out
S=${1:-s23v}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$S.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_default_$S.json 2> gpurun_out/bench_default_$S.err; echo -n "default rc=$? "; python tools/bench_line.py gpurun_out/bench_default_$S.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$S.json 2> gpurun_out/bench_ref_$S.err; echo "reference rc=$?"; cut -c1-300 gpurun_out/bench_ref_$S.json
timeout 600 python bench.py --workload drum > gpurun_out/bench_drum_$S.json 2> gpurun_out/bench_drum_$S.err; echo -n "drum rc=$? "; python tools/bench_line.py gpurun_out/bench_drum_$S.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$S.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list_$S.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip 230 --launch-count 1 -f -o gpurun_out/kstep64M_$S \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu64M_$S.log 2>&1; echo "ncu 64M rc=$?"
ls -la gpurun_out/kstep64M_$S.ncu-rep
