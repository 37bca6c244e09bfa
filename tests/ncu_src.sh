#!/bin/bash
# ncu --set full + source page of one kernel of the current build (profiling helper):
#   tests/ncu_src.sh kernel_regex n_images cfg tag   -> gpurun_out/ncu_<tag>.ncu-rep, _raw.csv, _src.csv
k=$1; n=$2; cfg=$3; tag=$4
ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_$tag \
    python tests/prof_run.py $n 2 $cfg > gpurun_out/ncu_$tag.log 2>&1
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_src.csv 2>/dev/null
