#!/bin/bash
# every BASELINE.json config at its named size on one GPU (config 5 is the default bench): one bench line each
mkdir -p gpurun_out
S=${1:-s18}
run() { # name args
  timeout 900 python bench.py $2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
run config1_box100k "--workload box_packing --n-per-gpu 100000"
run config2_drum1M "--workload drum"
run config3_hopper4M "--workload hopper --n-per-gpu 4000000"
run config4_jkr8M "--workload cohesive_jkr --n-per-gpu 8000000"
run config4_dmt8M "--workload cohesive_dmt --n-per-gpu 8000000"
