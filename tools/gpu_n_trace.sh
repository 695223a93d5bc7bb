#!/bin/bash
# N-GPU drum with LETHE_DEM_TRACE=1: per-phase wall-clock of every rebuild with exchange
N=${1:-4}; S=${2:-trace}
mkdir -p gpurun_out
export LETHE_DEM_TRACE=1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 3 --workload drum --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_n${N}_drum_$S.json 2> gpurun_out/bench_n${N}_drum_$S.err; echo "drum rc=$?"
python tools/bench_line.py gpurun_out/bench_n${N}_drum_$S.json
grep "lethe_dem" gpurun_out/bench_n${N}_drum_$S.err | tail -24 | cut -c1-260
