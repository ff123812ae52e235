#!/bin/bash
# developer A/B builds of the library with extra -D switches:  tools/ab_build.sh NAME "-DVB_X=1 ..." [tu ...]
# -> build/ab_NAME/libvegas_b200.so  (select at run time with VB200_LIB=build/ab_NAME/libvegas_b200.so)
# With a list of translation units (e.g. "fused_ridge fused_light") only those are recompiled with the
# switches; the rest is linked from the main build's objects.
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2 || true
only="$@"
out=build/ab_$name
mkdir -p $out
FL="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags"
pids=()
for f in $(cd vegas_b200/csrc && ls *.cu | sed "s/.cu//"); do
  if [ -n "$only" ] && ! echo " $only " | grep -q " $f "; then
    cp vegas_b200/csrc/$f.o $out/$f.o
  else
    nvcc $FL -c vegas_b200/csrc/$f.cu -o $out/$f.o 2>/dev/null & pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $out/libvegas_b200.so $out/*.o 2>/dev/null
echo built $out/libvegas_b200.so
