#!/bin/bash
# round 2 session 23: streamed host step (e2e path): bitwise test, then e2e of drum / 1 M / 64 M periodic with the pipeline off and on
mkdir -p gpurun_out
S=${1:-s23h}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "streamed or step_host" > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$S.log
bench() { # name workload-args
  timeout 600 python bench.py $2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
export LETHE_DEM_HOST_PIPELINE=0
bench drum_pipe0 "--workload drum"
export LETHE_DEM_HOST_PIPELINE=1
for st in 4 8 16; do
  export LETHE_DEM_HOST_STAGES=$st
  bench drum_stages$st "--workload drum"
  bench per1M_stages$st "--particles 1000000"
done
unset LETHE_DEM_HOST_STAGES
if [ "$2" != "no64" ]; then
bench per64M_pipe1 ""
tail -3 gpurun_out/bench_per64M_pipe1_$S.err
export LETHE_DEM_HOST_STAGES=32
bench per64M_stages32 ""
fi
