#!/bin/bash
# ncu captures behind profiles/*_r02.summary.txt (run on the GPU box through gpurun; reports land in gpurun_out/):
#   one `--set full` capture per kernel DESIGN.md quotes, and the launch list of a short bench.py run.
# Each report is summarised on the box (tools/ncu_summary.py: the metrics profiles/ quotes; tools/ncu_lines.py: the
# per-source-line table) into gpurun_out/prof_NAME_r02.summary.txt.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
cap() {   # cap NAME KERNEL_REGEX SKIP command...
  name=$1; rx=$2; skip=$3; shift 3
  $NCU -k regex:"$rx" --launch-skip "$skip" -c 1 -o $O/prof_${name}_r02 "$@" > $O/prof_${name}_r02.log 2>&1
  # the reports are 15-25 MB each and gpurun_out/ travels back only below 64 MiB: summarise here, drop the report
  { python tools/ncu_summary.py $O/prof_${name}_r02.ncu-rep; python tools/ncu_lines.py $O/prof_${name}_r02.ncu-rep 60; } \
      > $O/prof_${name}_r02.summary.txt 2>&1
  rm -f $O/prof_${name}_r02.ncu-rep
}
cap ridge1000 'k_engine' 4 python tools/profile_run.py ridge1000 1e8
cap ridge30   'k_engine' 4 python tools/profile_run.py ridge30 1e8
cap ridge1    'k_engine' 4 python tools/profile_run.py ridge1 1e8
cap reduce1   'k_reduce' 3 python tools/unfused_micro.py gauss8
cap reduce7   'k_engine' 20 python tools/unfused_micro.py pathint
cap sampler   'k_sample_x' 16 python tools/unfused_micro.py pathint
cap peaks20   'k_engine' 5 python tools/cfg_bench.py peaks20
cap pathint   'k_engine' 5 python tools/cfg_bench.py pathint
cap genz10    'k_engine' 5 python tools/cfg_bench.py genz
cap pdfmap    'k_pdf_map' 3 python tools/profile_run.py pdf 1e7
cap dyprofile 'k_dy_profile' 1 python tools/profile_run.py restratify 1e6
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r02.csv \
    python bench.py --steps 2 --warmup 3 --no-variants --no-cpu > $O/launches_r02.log 2>&1
ls -la $O/prof_*_r02.summary.txt
