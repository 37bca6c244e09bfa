o=gpurun_out
python -m pytest tests -q -m gpu > $o/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $o/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" > $o/r02_smoke.log 2>&1
python bench.py --steps 50 --warmup 5 > $o/r02_bench_n1.json 2> $o/r02_bench_n1.err
for c in 2 3 4; do python bench.py --config $c --steps 20 --warmup 3 > $o/r02_bench_cfg${c}_n1.json 2> $o/r02_bench_cfg${c}_n1.err; done
tail -3 $o/r02_pytest_gpu.log; tail -2 $o/r02_smoke.log; cat $o/r02_bench_n1.json | cut -c1-300
