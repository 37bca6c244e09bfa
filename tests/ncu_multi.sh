#!/bin/bash
# ncu --set full (+ source page) of several kernels x library variants on one box (profiling helper):
#   tests/ncu_multi.sh n_images cfg "name1 name2" "kernel_regex:skip ..."   -> gpurun_out/ncu_<name>_<k>.{ncu-rep,_raw.csv,_src.csv}
n=$1; cfg=$2; names=$3; kernels=$4
for v in $names; do
  if [ "$v" = "cur" ]; then lib=""; else lib="$PWD/build/var_$v/libb2j.so"; fi
  i=0
  for ks in $kernels; do
    k=${ks%%:*}; skip=${ks##*:}; i=$((i+1))
    B2J_LIBRARY=$lib ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o gpurun_out/ncu_${v}_$i \
        python tests/prof_run.py $n 2 $cfg > gpurun_out/ncu_${v}_$i.log 2>&1
    ncu -i gpurun_out/ncu_${v}_$i.ncu-rep --page raw --csv > gpurun_out/ncu_${v}_${i}_raw.csv 2>/dev/null
    ncu -i gpurun_out/ncu_${v}_$i.ncu-rep --page source --csv > gpurun_out/ncu_${v}_${i}_src.csv 2>/dev/null
    rm -f gpurun_out/ncu_${v}_$i.ncu-rep
  done
done
