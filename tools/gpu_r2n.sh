#!/bin/bash
# round 2 session 23: N-GPU check (incl. the streamed host step across slabs) and the default bench at N GPUs; usage gpu_r2n.sh <N> <tag>
N=${1:-2}; S=${2:-s23n}
mkdir -p gpurun_out
run() { timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_n${N}_$S.log 2>&1; echo "multi check rc=$?"
grep -E "^\[|MULTI_GPU_CHECK|lethe_dem\]|Error|error" gpurun_out/multi_check_n${N}_$S.log | cut -c1-250 | tail -14
for mode in 0 1; do
  export LETHE_DEM_HOST_PIPELINE=$mode
  run 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_per64M_pipe${mode}_$S.json 2> gpurun_out/bench_n${N}_per64M_pipe${mode}_$S.err; echo -n "N=$N 64M pipe$mode rc=$? "
  python tools/bench_line.py gpurun_out/bench_n${N}_per64M_pipe${mode}_$S.json
done
