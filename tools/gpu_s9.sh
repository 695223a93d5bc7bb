#!/bin/bash
# round 2 session 9: L1 / shared-memory split of k_step: queue size (variants) x carve-out, on the disordered 1 M drum and 1 M periodic box
mkdir -p gpurun_out
S=${1:-s9}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "mixed or mobility" > gpurun_out/pytest_mixed_$S.log 2>&1; echo "mixed rc=$?"
grep -E "mixed precision|mobility_status golden|passed|failed|Error" gpurun_out/pytest_mixed_$S.log | tail -12
bench() { # name workload-args
  timeout 300 python bench.py $2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
for lib in default lethe_b200/csrc/variants/lib_*.so; do
  if [ $lib = default ]; then unset LETHE_DEM_B200_LIB; name=default; else export LETHE_DEM_B200_LIB=$PWD/$lib; name=$(basename $lib .so); fi
  for co in -1 50 62 75 88; do
    if [ $co = -1 ]; then unset LETHE_DEM_CARVEOUT; else export LETHE_DEM_CARVEOUT=$co; fi
    bench drum_${name}_co$co "--workload drum"
    bench per1M_${name}_co$co "--particles 1000000"
  done
done
