#!/bin/bash
# Probe (run under gpurun): the whole GPU test suite (no -x), stage times of all configs, variants, launch lists.
tag=${1:-probe}; vars=${2:-}
o=gpurun_out
mkdir -p $o
python -m pytest tests -q -m gpu > $o/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> $o/${tag}_pytest.log
{
  echo "## configs[2]: 64 x 3840x2160 4:4:4 q95";  python tests/prof_run.py 64 10 2
  echo "## configs[3]: 1 x 7680x4320 4:2:2 q85";   python tests/prof_run.py 1 10 3
  echo "## configs[4]: 8192 x 500x375 4:2:0 q75";  python tests/prof_run.py 8192 10 4
  echo "## configs[1]: 256 x 1080p";  python tests/prof_run.py 256 10 1
  for v in $vars; do
    echo "## variant $v configs[2] 64"; B2J_LIBRARY=$PWD/build/var_$v/libb2j.so python tests/prof_run.py 64 10 2
    echo "## variant $v configs[4] 8192"; B2J_LIBRARY=$PWD/build/var_$v/libb2j.so python tests/prof_run.py 8192 10 4
  done
} > $o/${tag}_configs.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg2.csv python tests/prof_run.py 64 1 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file $o/${tag}_launches_cfg4.csv python tests/prof_run.py 8192 1 4 > /dev/null 2>&1
tail -8 $o/${tag}_pytest.log; cat $o/${tag}_configs.txt
