#!/bin/bash
# Builds a kernel variant of the library into variants/<name>/ for same-box A/B timing (tools/ab.py, CRB200_LIB).
# usage: tools/build_variant.sh <name> [-DMACRO ...]     (VARIANT_KINDS="1" limits the recompiled kernel kinds; others are
# taken from build/obj, so run `make` first)
set -e
name=$1; shift
out=variants/$name
kinds=${VARIANT_KINDS:-"0 1 6 8 10 12"}
mkdir -p $out
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -Xcompiler -fPIC -std=c++17"
cp build/obj/crb_inst_k*.o $out/
for k in $kinds; do for p in 0 1; do
  $NV "$@" -DCRB_INST_K=$k -DCRB_INST_PART=$p -c -o $out/crb_inst_k${k}_p${p}.o clownresampler_b200/csrc/crb_inst.cu &
done; done
$NV "$@" -c -o $out/crb_device.o clownresampler_b200/csrc/crb_device.cu &
gcc -O2 -fPIC -std=gnu99 "$@" -c -o $out/crb_plan.o clownresampler_b200/csrc/crb_plan.c
gcc -O2 -fPIC -std=gnu99 "$@" -c -o $out/crb_api.o clownresampler_b200/csrc/crb_api.c
gcc -O2 -fPIC -std=gnu99 "$@" -c -o $out/crb_voices.o clownresampler_b200/csrc/crb_voices.c
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libclownresampler_b200.so $out/*.o -lpthread -lm
echo built $out
