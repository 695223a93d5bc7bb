#!/bin/bash
# round 2 session 23: 256-bit row accesses + 32-byte history rows + pipelined img in k_step; variants timed on the same box
mkdir -p gpurun_out
S=${1:-s23}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$S.log
LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/lib_stage.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "(packing_parity_stepwise or periodic_parity_stepwise) and hertz_mindlin_limit_overlap" > gpurun_out/pytest_stage_$S.log 2>&1; echo "pytest stage rc=$?"; tail -3 gpurun_out/pytest_stage_$S.log
bench() { # name workload-args
  timeout 300 python bench.py $2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
for name in base new ld128 q256 stage; do
  export LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/lib_$name.so
  bench drum_${name} "--workload drum"
  bench per1M_${name} "--particles 1000000"
done
