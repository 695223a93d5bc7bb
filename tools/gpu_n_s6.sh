#!/bin/bash
# N-GPU round: pytest -m gpu multi test (slab vs oracle), default bench (64M periodic, strong scaling), drum (weak)
N=${1:-2}; S=${2:-s6}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "$3" = "check" ]; then
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi_$S.log 2>&1; echo "pytest multi rc=$?"
grep -E "^\[|MULTI_GPU_CHECK|passed|failed|Error" gpurun_out/pytest_multi_$S.log | cut -c1-260 | tail -14
fi
run 29512 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}_per64M_$S.json 2> gpurun_out/bench_n${N}_per64M_$S.err; echo "per64M rc=$?"
python tools/bench_line.py gpurun_out/bench_n${N}_per64M_$S.json
run 29513 bench.py --gpus $N --steps 3 --warmup 3 --workload drum --no-cpu-baseline > gpurun_out/bench_n${N}_drum_$S.json 2> gpurun_out/bench_n${N}_drum_$S.err; echo "drum rc=$?"
python tools/bench_line.py gpurun_out/bench_n${N}_drum_$S.json
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_per64M_$S.err gpurun_out/bench_n${N}_drum_$S.err | tail -8
