#!/bin/bash
# N-GPU correctness + bench (fused halo, peer-memory agreement); usage: gpu_call4.sh <ngpus>
N=${1:-2}
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1; echo "multi check rc=$?"
grep -E "^\[|MULTI_GPU_CHECK|lethe_dem\]|Error|error" gpurun_out/multi_check_n$N.log | tail -12
run 29512 bench.py --gpus $N --steps 400 --warmup 20 > gpurun_out/bench_n${N}_mailbox.json 2> gpurun_out/bench_n${N}_mailbox.err; echo "rc=$?"
cat gpurun_out/bench_n${N}_mailbox.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_mailbox.err | tail -5
LETHE_DEM_AGREE=nccl run 29513 bench.py --gpus $N --steps 400 --warmup 20 > gpurun_out/bench_n${N}_ncclagree.json 2> /dev/null; echo "rc=$?"
cat gpurun_out/bench_n${N}_ncclagree.json
run 29515 bench.py --gpus $N --steps 200 --warmup 20 --workload periodic_box --n-per-gpu 2000000 --settle 500 > gpurun_out/bench_n${N}_periodic.json 2> gpurun_out/bench_n${N}_periodic.err; echo "rc=$?"
cat gpurun_out/bench_n${N}_periodic.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_periodic.err | tail -5
