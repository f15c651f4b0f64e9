#!/bin/bash
mkdir -p gpurun_out
MSCS_GPU_RANDOM=10 timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x -s > gpurun_out/pytest_gpu_r2i.log 2>&1
echo "pytest exit $?"; grep -i "passed\|failed\|error\|cross entropy fwd\|co-losses\|: loss " gpurun_out/pytest_gpu_r2i.log | head -20
for i in 1 2; do
  for v in "0 0" "1 0" "0 1"; do
    set -- $v
    MSCS_DENSE=$1 MSCS_GATHER=$2 timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_d$1g$2_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_d$1g$2_$i.json')); print('MSCS_DENSE $1 MSCS_GATHER $2', $i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
  done
done
