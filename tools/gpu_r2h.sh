#!/bin/bash
# round 2 session 23: neighbour search as one flat candidate walk (DEM_NB_FLAT) vs the nested walk — parity of the default, rebuild time
mkdir -p gpurun_out
S=${1:-s23s}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$S.log
bench() { # name workload-args
  timeout 600 python bench.py $2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
for name in nest flat; do
  export LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/lib_$name.so
  bench drum_${name} "--workload drum"
  bench per1M_${name} "--particles 1000000"
done
bench per64M_flat ""
