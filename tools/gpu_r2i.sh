#!/bin/bash
# round 2 session 23: streamed host step in transfer order at 64 M — segment size / stage count
mkdir -p gpurun_out
S=${1:-s23t}
bench() { # name workload-args
  timeout 600 python bench.py $2 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 8 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
export LETHE_DEM_HOST_MAX_SEGS=16384
for cfg in "32 4096" "32 65536" "48 16384" "16 16384"; do
  set -- $cfg
  export LETHE_DEM_HOST_STAGES=$1 LETHE_DEM_HOST_SEG_ROWS=$2
  bench per64M_k$1_seg$2 ""
done
