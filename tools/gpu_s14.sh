#!/bin/bash
# round 2 session 14: full GPU suite, ncu launch list of the default bench command (64 M periodic box), one ncu --set full
# capture of k_step on that workload (roofline.traffic of this round) and one on the 1 M drum
mkdir -p gpurun_out
S=${1:-s14}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_$S.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$S.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$S.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_list_$S.log 2>&1; echo "launch list rc=$?"
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/launches_s14.csv")) if len(r) > 10 and r[0].isdigit()]
t = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][-60:]
    t[name][0] += 1
    t[name][1] += float(r[-1].replace(",", "")) / 1e6
tot = sum(v[1] for v in t.values())
for k, v in sorted(t.items(), key=lambda kv: -kv[1][1])[:16]:
    print("%-62s n=%4d  %9.3f ms  %5.1f %%" % (k, v[0], v[1], 100 * v[1] / tot))
P
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip 230 --launch-count 1 -f -o gpurun_out/kstep64M_$S \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu64M_$S.log 2>&1; echo "ncu 64M rc=$?"
ls -la gpurun_out/kstep64M_$S.ncu-rep
