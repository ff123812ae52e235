#!/bin/bash
# developer A/B builds of the library with extra -D switches:  tools/ab_build.sh NAME "-DVB_X=1 ..."
# -> build/ab_NAME/libvegas_b200.so  (select at run time with VB200_LIB=build/ab_NAME/libvegas_b200.so)
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=build/ab_$name
mkdir -p $out
FL="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $flags"
pids=()
for f in $(cd vegas_b200/csrc && ls *.cu | sed "s/.cu//"); do
  nvcc $FL -c vegas_b200/csrc/$f.cu -o $out/$f.o & pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o $out/libvegas_b200.so $out/*.o 2>/dev/null
echo built $out/libvegas_b200.so
