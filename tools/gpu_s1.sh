#!/bin/bash
# session 1 of round 2: baseline numbers on the disordered workloads + the "evaluate once" timing experiment
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_s1.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_s1.log
python bench.py --workload drum --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_drum_s1.json 2> gpurun_out/bench_drum_s1.err; echo "drum rc=$?"
LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/liblethe_dem_b200_half.so python bench.py --workload drum --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_drum_half_s1.json 2> gpurun_out/bench_drum_half_s1.err; echo "drum half rc=$?"
python bench.py --particles 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_per1M_s1.json 2> gpurun_out/bench_per1M_s1.err; echo "per1M rc=$?"
LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/liblethe_dem_b200_half.so python bench.py --particles 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_per1M_half_s1.json 2> gpurun_out/bench_per1M_half_s1.err; echo "per1M half rc=$?"
/usr/bin/time -v python bench.py --steps 3 --warmup 3 > gpurun_out/bench_per64M_s1.json 2> gpurun_out/bench_per64M_s1.err; echo "per64M rc=$?"
tail -5 gpurun_out/bench_per64M_s1.err
for f in gpurun_out/bench_*_s1.json; do echo $f; python - "$f" <<'P'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l['roofline']; c=l['config']
    print(' value %.3e ms/dem %.4f k_ms %.4f frac %.3f C %.2f T %.2f rebuilds %d (each %.2f ms, share %.3f) e2e %.3e setup %.0fs'%(l['value'],c['ms_per_dem_step'],r['kernel_ms'],r['frac'],r['C_half'],r['T_half'],c['rebuilds_in_timed_region'],r['rebuild_ms_each'],r['rebuild_share_of_step'],l['e2e']['value'],c['setup_s']))
except Exception as e: print(' failed',e)
P
done
