#!/bin/bash
# usage: tools/build_variant.sh <name> "<-D flags>"  -> lethe_b200/csrc/variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../lethe_b200/csrc"
mkdir -p variants/$1
for f in dem_step dem_kernels dem_engine dem_multi dem_solid; do
  if [ $f = dem_step ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a $2 -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -c $f.cu -o variants/$1/$f.o
  else
    cp $f.o variants/$1/$f.o
  fi
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$1.so variants/$1/*.o -lcudart
rm -rf variants/$1
