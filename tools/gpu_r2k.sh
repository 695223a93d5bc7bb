#!/bin/bash
# round 2 session 23: closing verification with sweep depth 3 — full GPU suite, smoke, default bench, drum; QUEUE 256 beside it
mkdir -p gpurun_out
S=${1:-s23z}
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$S.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/bench_default_$S.json 2> gpurun_out/bench_default_$S.err; echo -n "default rc=$? "; python tools/bench_line.py gpurun_out/bench_default_$S.json
timeout 300 python bench.py --workload drum --no-cpu-baseline > gpurun_out/bench_drum_$S.json 2> gpurun_out/bench_drum_$S.err; echo -n "drum rc=$? "; python tools/bench_line.py gpurun_out/bench_drum_$S.json
export LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/lib_sw3q256.so
timeout 300 python bench.py --workload drum --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_drum_q256_$S.json 2>/dev/null; echo -n "drum q256 "; python tools/bench_line.py gpurun_out/bench_drum_q256_$S.json
timeout 300 python bench.py --particles 1000000 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_per1M_q256_$S.json 2>/dev/null; echo -n "per1M q256 "; python tools/bench_line.py gpurun_out/bench_per1M_q256_$S.json
