#!/bin/bash
# round 2 session 23: streamed host step with automatic stage count; full default bench (64 M) and the drum
mkdir -p gpurun_out
S=${1:-s23i}
bench() { # name workload-args
  timeout 900 python bench.py $2 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
bench drum "--workload drum --no-cpu-baseline"
bench per1M "--particles 1000000 --no-cpu-baseline"
bench hopper "--workload hopper --n-per-gpu 4000000 --no-cpu-baseline --steps 2"
bench per64M ""
