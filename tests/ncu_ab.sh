#!/bin/bash
# ncu --set full of one kernel for several library variants (profiling helper):
#   tests/ncu_ab.sh kernel_regex n_images cfg name1 name2 ...   -> gpurun_out/ncu_<name>.ncu-rep + _raw.csv
k=$1; n=$2; cfg=$3; shift 3
for v in "$@"; do
  if [ "$v" = "cur" ]; then lib=""; else lib="$PWD/build/var_$v/libb2j.so"; fi
  B2J_LIBRARY=$lib ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -f -o gpurun_out/ncu_$v \
      python tests/prof_run.py $n 2 $cfg > gpurun_out/ncu_$v.log 2>&1
  ncu -i gpurun_out/ncu_$v.ncu-rep --page raw --csv > gpurun_out/ncu_${v}_raw.csv 2>/dev/null
done
