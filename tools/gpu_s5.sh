#!/bin/bash
# round 2 session 5: FastMath pair model + cached-candidate rebuild: self-test + parity, then drum / 1M periodic
# timings of the default build and of every variant library under lethe_b200/csrc/variants
mkdir -p gpurun_out
S=${1:-s5}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$S.log
for lib in default lethe_b200/csrc/variants/lib_*.so; do
  if [ $lib = default ]; then unset LETHE_DEM_B200_LIB; name=default; else export LETHE_DEM_B200_LIB=$PWD/$lib; name=$(basename $lib .so); fi
  timeout 300 python bench.py --workload drum --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_drum_${S}_$name.json 2> gpurun_out/bench_drum_${S}_$name.err; echo "$name drum rc=$?"
  python tools/bench_line.py gpurun_out/bench_drum_${S}_$name.json
  timeout 300 python bench.py --particles 1000000 --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_per1M_${S}_$name.json 2> gpurun_out/bench_per1M_${S}_$name.err; echo "$name per1M rc=$?"
  python tools/bench_line.py gpurun_out/bench_per1M_${S}_$name.json
done
unset LETHE_DEM_B200_LIB
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip 3200 --launch-count 1 -f -o gpurun_out/kstep_$S \
  python bench.py --workload drum --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$S.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$S.log
fi
