#!/bin/bash
out=$1
run() { echo "## $*" >> $out; env "$@" timeout 40 python ${QB:-tools/quick_bench.py} $N >> $out 2>&1; }
N=1; run A=0
N=30
run A=0
for ns in 1000 2000 3000 5000; do run VB200_STAGGER_NS=$ns; done
N=1000; run A=0; run VB200_STAGGER_NS=100000
