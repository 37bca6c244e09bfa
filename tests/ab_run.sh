#!/bin/bash
# A/B of library variants on one box: tests/ab_run.sh out.log n_images steps cfg name1 name2 ... (profiling helper)
out=$1; n=$2; steps=$3; cfg=$4; shift 4
for rep in 1 2; do
  for v in "$@"; do
    if [ "$v" = "cur" ]; then lib=""; else lib="$PWD/build/var_$v/libb2j.so"; fi
    echo -n "$v: " >> $out
    B2J_LIBRARY=$lib python tests/prof_run.py $n $steps $cfg 2>&1 | grep "steps\|rror" >> $out
  done
done
