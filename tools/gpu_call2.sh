#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s8b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s8b.log
tail -3 gpurun_out/pytest_s8b.log
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run2 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_fused.log 2>&1; echo "multi fused rc=$?"
grep -E "^\[|MULTI_GPU_CHECK|lethe_dem\]|Error|error" gpurun_out/multi_check_fused.log | tail -12
LETHE_DEM_HALO=nccl run2 29513 tests/multi_gpu_check.py > gpurun_out/multi_check_nccl.log 2>&1; echo "multi nccl rc=$?"
grep -E "^\[|MULTI_GPU_CHECK|lethe_dem\]|Error|error" gpurun_out/multi_check_nccl.log | tail -12
run2 29512 bench.py --gpus 2 --steps 400 --warmup 20 > gpurun_out/bench_n2_fused.json 2> gpurun_out/bench_n2_fused.err; echo "rc=$?"
cat gpurun_out/bench_n2_fused.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n2_fused.err | tail -5
LETHE_DEM_HALO=nccl run2 29514 bench.py --gpus 2 --steps 400 --warmup 20 > gpurun_out/bench_n2_nccl.json 2> gpurun_out/bench_n2_nccl.err; echo "rc=$?"
cat gpurun_out/bench_n2_nccl.json
run2 29515 bench.py --gpus 2 --steps 200 --warmup 20 --workload periodic_box --n-per-gpu 2000000 --settle 500 > gpurun_out/bench_n2_periodic.json 2> gpurun_out/bench_n2_periodic.err; echo "rc=$?"
cat gpurun_out/bench_n2_periodic.json; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n2_periodic.err | tail -5
