#!/bin/bash
# ncu --set full capture of k_picard_resident (B equilibria, default 148 = one per SM); run under gpurun.
B=${1:-148}
OUT=${2:-gpurun_out/picard_res}
python tools/prof_picard.py $B 1 > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_picard_resident -s 1 -c 1 -f -o $OUT \
    python tools/prof_picard.py $B 1 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
