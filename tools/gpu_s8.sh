#!/bin/bash
# round 2 session 8: mixed-precision pair model: documented bound (lock-step vs the FP64 oracle), then timings of
# f64 / mixed on the default build and of mixed on every variant library under lethe_b200/csrc/variants
mkdir -p gpurun_out
S=${1:-s8}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mixed or mobility or bitwise or packing_parity" > gpurun_out/pytest_mixed_$S.log 2>&1; echo "mixed rc=$?"
grep -E "mixed precision|mobility_status golden|passed|failed|Error" gpurun_out/pytest_mixed_$S.log | tail -20
bench() { # name workload-args precision
  timeout 300 python bench.py $2 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --precision $3 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo "$1 rc=$?"
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
unset LETHE_DEM_B200_LIB
bench drum_f64 "--workload drum" f64
bench drum_mixed "--workload drum" mixed
bench per1M_f64 "--particles 1000000" f64
bench per1M_mixed "--particles 1000000" mixed
for lib in lethe_b200/csrc/variants/lib_*.so; do
  export LETHE_DEM_B200_LIB=$PWD/$lib; name=$(basename $lib .so)
  bench drum_mixed_$name "--workload drum" mixed
  bench per1M_mixed_$name "--particles 1000000" mixed
done
unset LETHE_DEM_B200_LIB
