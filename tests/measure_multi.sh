#!/bin/bash
# Multi-GPU measurements on one N-GPU box (run under gpurun --gpus N): bench lines of the configs BASELINE sweeps over
# GPUs, the pinned-copy probe with all ranks at once, the single-process C-ABI call. Output: gpurun_out/<tag>_*.
# usage: tests/measure_multi.sh tag "list of rank counts"      e.g.  r02 "1 8"
tag=${1:-r02}; NS=${2:-"1 2"}
o=gpurun_out
mkdir -p $o
NG=$(nvidia-smi -L | wc -l)
nvidia-smi topo -m > $o/${tag}_topo_${NG}gpu.txt 2>&1
run() { # config gpus steps
  if [ "$2" = "1" ]; then python bench.py --config $1 --steps $3 --warmup 3 --no-parity > $o/${tag}_bench_cfg$1_n$2.json 2> $o/${tag}_bench_cfg$1_n$2.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $2 --config $1 --steps $3 --warmup 3 --no-parity > $o/${tag}_bench_cfg$1_n$2.json 2> $o/${tag}_bench_cfg$1_n$2.err; fi
}
for g in $NS; do run 1 $g 30; done
for g in $NS; do if [ $g != 1 ]; then run 4 $g 10; run 3 $g 30; fi; done
{
  for g in $NS; do
    echo "## $g ranks"; if [ $g = 1 ]; then python tests/d2h_probe.py; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29519 tests/d2h_probe.py 2>/dev/null | grep rank; fi
  done
} > $o/${tag}_d2h_probe_${NG}gpu.txt 2>&1
python tests/multi_probe.py 1 $((64*NG)) 2 > $o/${tag}_multi_cabi_cfg1_${NG}gpu.txt 2>&1
python tests/multi_probe.py 4 4096 2 > $o/${tag}_multi_cabi_cfg4_${NG}gpu.txt 2>&1
python -m pytest tests/test_gpu_parity.py -q -k "multi or host_ex or decode_host_api" > $o/${tag}_pytest_multi_${NG}gpu.log 2>&1
grep -h '"metric"' $o/${tag}_bench_cfg*_n*.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['baseline_config'], d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac_pipeline'], d['e2e']['value'] if d.get('e2e') else None)"
cat $o/${tag}_d2h_probe_${NG}gpu.txt $o/${tag}_multi_cabi_cfg1_${NG}gpu.txt $o/${tag}_multi_cabi_cfg4_${NG}gpu.txt; tail -3 $o/${tag}_pytest_multi_${NG}gpu.log
