#!/bin/bash
# usage: tools/ncu_report.sh <tag> <kernel-mangled-substring> [top]
TAG=$1; K=${2:-k_stepILi2ELi1ELb0}; TOP=${3:-25}
python tools/ncu_key.py gpurun_out/kstep_$TAG.ncu-rep | head -24
ncu -i gpurun_out/kstep_$TAG.ncu-rep --page source --csv > gpurun_out/kstep_${TAG}_src.csv 2>/dev/null
python tools/ncu_mix.py gpurun_out/kstep_${TAG}_src.csv 12
python tools/ncu_lines.py gpurun_out/kstep_${TAG}_src.csv lethe_b200/csrc/liblethe_dem_b200.so $K $TOP
