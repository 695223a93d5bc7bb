#!/bin/bash
# N-GPU correctness + weak-scaling bench lines; usage: gpu_calln.sh <ngpus> [check] [periodic n-per-gpu]
N=${1:-2}; CHECK=${2:-1}; NPG=${3:-2000000}
mkdir -p gpurun_out
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
if [ "$CHECK" = 1 ]; then
  run 29511 tests/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1; echo "multi check rc=$?"
  grep -E "^\[|MULTI_GPU_CHECK|lethe_dem\]|Error|error" gpurun_out/multi_check_n$N.log | cut -c1-250 | tail -12
fi
run 29512 bench.py --gpus $N --steps 400 --warmup 20 > gpurun_out/bench_n${N}_drum.json 2> gpurun_out/bench_n${N}_drum.err; echo "rc=$?"
run 29515 bench.py --gpus $N --steps 200 --warmup 20 --workload periodic_box --n-per-gpu $NPG > gpurun_out/bench_n${N}_periodic.json 2> gpurun_out/bench_n${N}_periodic.err; echo "rc=$?"
python - <<PY
import json
for f in ["bench_n${N}_drum","bench_n${N}_periodic"]:
    try:
        j=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]); r=j["roofline"]
        print(f, j["config"]["particles"], "value %.4g"%j["value"], "ms/step %.4f"%j["ms_per_step"], "kernel_ms %.4f"%r["kernel_ms"], "share %.3f"%r["kernel_share_of_step"], "rebuild ms %.2f/%d"%(r["rebuild_ms_total"], r["rebuilds"]))
    except Exception as e: print(f, "FAILED", e)
PY
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_drum.err gpurun_out/bench_n${N}_periodic.err | tail -6
