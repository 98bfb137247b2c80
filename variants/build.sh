#!/bin/bash
# build a library variant: variants/build.sh <name> [-DMACRO=VALUE ...]
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o variants/$name.so active_gs_b200/csrc/*.cu
