#!/bin/bash
# round 2 session 23: last k_step variants on the 256-bit kernel (round-operand prefetch to L1 / off, sweep depth 2, five blocks per SM)
mkdir -p gpurun_out
S=${1:-s23x}
bench() { # name workload-args
  timeout 300 python bench.py $2 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
for name in sw2 sw1 sw3 sw2q256 sw2pf2; do
  export LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/lib_$name.so
  bench drum_${name} "--workload drum"
  bench per1M_${name} "--particles 1000000"
done
