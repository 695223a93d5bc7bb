#!/bin/bash
# round 2 session 23: streamed host step, segment size / stage count tuning
mkdir -p gpurun_out
S=${1:-s23g}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed or step_host" > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$S.log
bench() { # name workload-args
  timeout 600 python bench.py $2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
export LETHE_DEM_HOST_MAX_SEGS=16384
for cfg in "4 2048" "4 8192" "8 2048" "2 4096"; do
  set -- $cfg
  export LETHE_DEM_HOST_STAGES=$1 LETHE_DEM_HOST_SEG_ROWS=$2
  bench drum_k$1_seg$2 "--workload drum"
  bench per1M_k$1_seg$2 "--particles 1000000"
done
for cfg in "16 4096" "16 8192" "8 16384" "32 8192"; do
  set -- $cfg
  export LETHE_DEM_HOST_STAGES=$1 LETHE_DEM_HOST_SEG_ROWS=$2
  bench per64M_k$1_seg$2 ""
done
