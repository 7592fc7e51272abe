#!/bin/bash
# usage: tools/devbuild.sh <name> [extra nvcc flags...] — single-instantiation (C172S/WA unless -DFB_ONLY_X1_WA is given)
# development build into devlibs/<name>.so; prints the spill statistics of the integrator kernels
name=$1; shift
mkdir -p devlibs build/dev_$name
ONLY=-DFB_ONLY_S_WA
for a in "$@"; do [ "$a" = "-DFB_ONLY_X1_WA" ] && ONLY=; done
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $ONLY -Xptxas -v"
for tu in flightbatch flightbatch_complete; do
  EXTRA=; [ $tu = flightbatch ] && EXTRA="-Xptxas -dlcm=cg"
  nvcc $FLAGS $EXTRA "$@" -c -o build/dev_$name/$tu.o flight.jl_b200/csrc/$tu.cu > build/dev_$name/$tu.log 2>&1 &
done
wait
grep -h -A1 "Function properties for _Z11step_kernel" build/dev_$name/*.log | grep -v "^--" | sed 's/ptxas info    : Function properties for //'
grep -h -i "error" build/dev_$name/*.log | head -5
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o devlibs/$name.so build/dev_$name/flightbatch.o build/dev_$name/flightbatch_complete.o
