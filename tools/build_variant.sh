#!/bin/bash
# Builds a kernel variant of the library into build/variants/<name>/ for same-box A/B timing.
# usage: tools/build_variant.sh <name> [-DMACRO ...]
set -e
name=$1; shift
out=variants/$name
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -std=c++17 "$@" -c -o $out/crb_device.o clownresampler_b200/csrc/crb_device.cu
gcc -O2 -fPIC -std=gnu99 "$@" -c -o $out/crb_plan.o clownresampler_b200/csrc/crb_plan.c
gcc -O2 -fPIC -std=gnu99 "$@" -c -o $out/crb_api.o clownresampler_b200/csrc/crb_api.c
gcc -O2 -fPIC -std=gnu99 "$@" -c -o $out/crb_voices.o clownresampler_b200/csrc/crb_voices.c
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libclownresampler_b200.so $out/*.o -lpthread -lm
echo built $out
