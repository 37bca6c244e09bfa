#!/bin/bash
# End-of-round measurements on one B200 (run under gpurun): bench lines, ncu launch lists, ncu --set full of
# the heavy kernels, stage times of the other BASELINE configs. Output: gpurun_out/<tag>_*.
tag=${1:-r01e}
o=gpurun_out
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_n1.json 2> $o/${tag}_bench_reference_n1.err
python bench.py --steps 50 --warmup 5 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/${tag}_launches_bench256.csv \
    python tests/prof_run.py 256 3 > $o/${tag}_launches_bench256.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_unstuff_fused|k_huff_decode|k_idct_csc' -s 3 -c 3 -f -o $o/${tag}_full \
    python tests/prof_run.py 256 3 > $o/${tag}_full.log 2>&1
ncu -i $o/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>/dev/null
python tests/ncu_extract.py $o/${tag}_full_raw.csv > $o/${tag}_ncu_full_summary.txt
{
  echo "# stage times (CUDA events, last of 10 steps) of the other BASELINE configs at full size"
  echo "## configs[2]: 64 x 3840x2160 4:4:4 q95, no restart markers";  python tests/prof_run.py 64 10 2
  echo "## configs[3]: 1 x 7680x4320 4:2:2 q85, no restart markers";   python tests/prof_run.py 1 10 3
  echo "## configs[4]: 8192 x 500x375 4:2:0 q75, no restart markers";  python tests/prof_run.py 8192 10 4
} > $o/${tag}_other_configs.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg3_64x4k.csv \
    python tests/prof_run.py 64 1 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg5_8192.csv \
    python tests/prof_run.py 8192 1 4 > /dev/null 2>&1
