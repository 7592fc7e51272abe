#!/bin/bash
# usage: tools/devbuild.sh <name> [extra nvcc flags...] — single-instantiation (C172S/WA) development build into devlibs/<name>.so
name=$1; shift
mkdir -p devlibs
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DFB_ONLY_S_WA -Xptxas -v "$@" \
  -o devlibs/$name.so flight.jl_b200/csrc/flightbatch.cu 2>&1 | grep -A2 "Compiling entry function '_Z11step_kernel" | grep -v "^--" | grep -v Compiling
