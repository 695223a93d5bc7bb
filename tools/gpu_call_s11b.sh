#!/bin/bash
# Session-11 closing verification on 2 GPUs: full GPU suite, smoke, N=1 and N=2 bench stdout (exactly one JSON line each).
TAG=${1:-r01_s11b}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_$TAG.log
tail -4 gpurun_out/pytest_$TAG.log
python __graft_entry__.py smoke 2>&1 | tail -1
python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err; echo "n1 rc=$? lines=$(wc -l < gpurun_out/bench_n1_$TAG.json)"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2_$TAG.json 2> gpurun_out/bench_n2_$TAG.err; echo "n2 rc=$? lines=$(wc -l < gpurun_out/bench_n2_$TAG.json)"
head -c 200 gpurun_out/bench_n2_$TAG.json; echo
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2_$TAG.json 2> gpurun_out/bench_ref_n2_$TAG.err; echo "ref n2 rc=$? lines=$(wc -l < gpurun_out/bench_ref_n2_$TAG.json)"
