#!/bin/bash
# developer sweep on the GPU box: tools/sweep.sh OUT  (A/B libraries under build/ab_*/, env switches)
out=$1
run() { echo "## $*" >> $out; env "$@" timeout 120 python tools/quick_bench.py $N >> $out 2>&1; }
for N in 1 30; do
  for cap in 2048 3072 5120; do
    N=$N run VB200_LCAP=$cap
    N=$N run VB200_LCAP=$cap VB200_LIB=build/ab_c1/libvegas_b200.so
  done
  N=$N run VB200_LIB=build/ab_l3/libvegas_b200.so VB200_LCAP=2560
  N=$N run VB200_LIB=build/ab_l3/libvegas_b200.so VB200_LCAP=1536
  N=$N run VB200_LIB=build/ab_lw8/libvegas_b200.so VB200_LCAP=3072
done
N=30; run VB200_RIDGE_LIGHT_N=0 VB200_LIB=build/ab_h4/libvegas_b200.so
N=1000
run VB200_LIB=build/ab_h4/libvegas_b200.so
run VB200_LIB=build/ab_c1/libvegas_b200.so
run VB200_RIDGE_PAR=0
run VB200_CAP=1024 VB200_LIB=build/ab_h4/libvegas_b200.so
