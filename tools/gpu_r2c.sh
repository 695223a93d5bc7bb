#!/bin/bash
# round 2 session 23b: PCIe duplex capability of the box, streaming cache hints in k_step, ncu --set full of the new k_step (1 M drum)
mkdir -p gpurun_out
S=${1:-s23b}
python tools/pcie_duplex.py 2e9 2>&1 | tee gpurun_out/pcie_duplex_$S.log
bench() { # name workload-args
  timeout 300 python bench.py $2 --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/bench_${1}_$S.json 2> gpurun_out/bench_${1}_$S.err; echo -n "$1 rc=$? "
  python tools/bench_line.py gpurun_out/bench_${1}_$S.json
}
for name in new cs; do
  export LETHE_DEM_B200_LIB=$PWD/lethe_b200/csrc/variants/lib_$name.so
  bench drum_${name} "--workload drum"
  bench per1M_${name} "--particles 1000000"
done
unset LETHE_DEM_B200_LIB
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step --launch-skip 3200 --launch-count 1 -f -o gpurun_out/kstep_$S \
  python bench.py --workload drum --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_$S.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$S.log
ls -la gpurun_out/kstep_$S.ncu-rep
