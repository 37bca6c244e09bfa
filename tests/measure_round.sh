#!/bin/bash
# End-of-round measurements on one B200 (run under gpurun): bench lines of every BASELINE config and of the reference
# arm, ncu launch lists, ncu --set full of the heavy kernels of the restart-interval and the self-synchronising path.
# Output: gpurun_out/<tag>_*.   usage: tests/measure_round.sh tag [quick]
tag=${1:-r02}
o=gpurun_out
mkdir -p $o
python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference_n1.json 2> $o/${tag}_bench_reference_n1.err
python bench.py --steps 50 --warmup 5 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
for c in 2 3 4; do
  python bench.py --config $c --steps 20 --warmup 3 > $o/${tag}_bench_cfg${c}_n1.json 2> $o/${tag}_bench_cfg${c}_n1.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $o/${tag}_launches_bench256.csv \
    python tests/prof_run.py 256 3 > $o/${tag}_launches_bench256.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg2_64x4k.csv python tests/prof_run.py 64 1 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg3_8k.csv python tests/prof_run.py 1 1 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg4_8192.csv python tests/prof_run.py 8192 1 4 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_unstuff_fused|k_huff_decode|k_idct_csc' -s 3 -c 3 -f -o $o/${tag}_full \
    python tests/prof_run.py 256 3 > $o/${tag}_full.log 2>&1
ncu -i $o/${tag}_full.ncu-rep --page raw --csv > $o/${tag}_full_raw.csv 2>/dev/null
python tests/ncu_extract.py $o/${tag}_full_raw.csv > $o/${tag}_ncu_full_summary.txt
ncu --set full --clock-control none --import-source on -k regex:'k_sync_chunks|k_sync_sweep|k_huff_decode' -s 3 -c 3 -f -o $o/${tag}_sync \
    python tests/prof_run.py 64 2 2 > $o/${tag}_sync.log 2>&1
ncu -i $o/${tag}_sync.ncu-rep --page raw --csv > $o/${tag}_sync_raw.csv 2>/dev/null
python tests/ncu_extract.py $o/${tag}_sync_raw.csv > $o/${tag}_ncu_sync_summary.txt
rm -f $o/${tag}_full.ncu-rep $o/${tag}_sync.ncu-rep
cat $o/${tag}_bench_n1.json; tail -2 $o/${tag}_bench_n1.err
