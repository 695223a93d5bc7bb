#!/bin/bash
# bench every variant library (kernel time only) ; usage: bash tools/gpu_variants.sh [settle]
SETTLE=${1:-3000}
mkdir -p gpurun_out
for lib in default lethe_b200/csrc/variants/lib_*.so; do
  if [ $lib = default ]; then unset LETHE_DEM_B200_LIB; name=default; else export LETHE_DEM_B200_LIB=$PWD/$lib; name=$(basename $lib .so); fi
  python bench.py --steps 200 --warmup 10 --settle $SETTLE --no-cpu-baseline --e2e-steps 1 > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    j=json.load(open("gpurun_out/var_$name.json"))
    print("$name", "value %.4g"%j["value"], "kernel_ms %.4f"%j["roofline"]["kernel_ms"], "ms/step %.4f"%j["ms_per_step"])
except Exception as e:
    print("$name", "FAILED", e)
PY
done
