#!/bin/bash
# 4 GPUs: slab-vs-oracle check (incl. load balancing) and the default bench (64 M periodic box, strong scaling)
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_n4_s19.log 2>&1; echo "multi check rc=$?"
grep -E "^\[|MULTI_GPU_CHECK" gpurun_out/multi_check_n4_s19.log | cut -c1-250 | tail -14
run 29512 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n4_per64M_s19.json 2> gpurun_out/bench_n4_per64M_s19.err; echo "per64M rc=$?"
python tools/bench_line.py gpurun_out/bench_n4_per64M_s19.json
