#!/bin/bash
# round 2 session 10: sweep-ahead prefetch variants of k_step + the mixed-precision bounds
mkdir -p gpurun_out
S=${1:-s10}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mixed" > gpurun_out/pytest_mixed_$S.log 2>&1; echo "mixed rc=$?"
grep -E "mixed precision [a-zA-Z_]+/|passed|failed" gpurun_out/pytest_mixed_$S.log | tail -12
bench() { # name workload-args
  timeout 300 python bench.py $2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
for lib in default lethe_b200/csrc/variants/lib_*.so; do
  if [ $lib = default ]; then unset LETHE_DEM_B200_LIB; name=default; else export LETHE_DEM_B200_LIB=$PWD/$lib; name=$(basename $lib .so); fi
  bench drum_${name} "--workload drum"
  bench per1M_${name} "--particles 1000000"
done
