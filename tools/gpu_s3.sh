#!/bin/bash
# parity + timings + one ncu --set full capture of k_step on the settled disordered drum
mkdir -p gpurun_out
S=${1:-s3}
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_$S.log
timeout 300 python bench.py --workload drum --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_drum_$S.json 2> gpurun_out/bench_drum_$S.err; echo "drum rc=$?"
timeout 300 python bench.py --particles 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_per1M_$S.json 2> gpurun_out/bench_per1M_$S.err; echo "per1M rc=$?"
for f in gpurun_out/bench_*_$S.json; do echo $f; python - "$f" <<'P'
import json,sys
try:
    l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=l['roofline']; c=l['config']
    print(' value %.3e ms/dem %.4f k_ms %.4f frac %.3f C %.2f T %.2f rebuilds %d (each %.2f ms, share %.3f) e2e %.3e setup %.0fs'%(l['value'],c['ms_per_dem_step'],r['kernel_ms'],r['frac'],r['C_half'],r['T_half'],c['rebuilds_in_timed_region'],r['rebuild_ms_each'],r['rebuild_share_of_step'],l['e2e']['value'],c['setup_s']))
except Exception as e: print(' failed',e)
P
done
if [ "$2" != "noncu" ]; then
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip 3200 --launch-count 1 -f -o gpurun_out/kstep_$S \
  python bench.py --workload drum --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$S.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$S.log
ls -la gpurun_out/kstep_$S.ncu-rep
fi
