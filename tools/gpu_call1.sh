#!/bin/bash
# round-1 session-8 call: parity, pipelined vs synchronous bench, 2-GPU correctness + bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_s8.log
tail -3 gpurun_out/pytest_s8.log
python bench.py --steps 400 --warmup 20 --no-cpu-baseline > gpurun_out/bench_s8_pipe.json 2> gpurun_out/bench_s8_pipe.err
LETHE_DEM_NO_PIPELINE=1 python bench.py --steps 400 --warmup 20 --no-cpu-baseline --e2e-steps 3 > gpurun_out/bench_s8_nopipe.json 2> gpurun_out/bench_s8_nopipe.err
python bench.py --steps 400 --warmup 20 --no-cpu-baseline --e2e-steps 3 --n-per-gpu 100000 > gpurun_out/bench_s8_100k.json 2> gpurun_out/bench_s8_100k.err
cat gpurun_out/bench_s8_pipe.json gpurun_out/bench_s8_nopipe.json gpurun_out/bench_s8_100k.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_s8.log 2>&1; echo "multi rc=$?" >> gpurun_out/multi_check_s8.log
grep -E "^\[|MULTI_GPU_CHECK|rc=|Error|error" gpurun_out/multi_check_s8.log | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 400 --warmup 20 > gpurun_out/bench_s8_n2.json 2> gpurun_out/bench_s8_n2.err; echo "rc=$?"
cat gpurun_out/bench_s8_n2.json; tail -3 gpurun_out/bench_s8_n2.err
