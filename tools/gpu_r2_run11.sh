#!/bin/bash
mkdir -p gpurun_out
MSCS_GPU_RANDOM=20 timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2j.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2j.log
for i in 1 2; do
  for v in tma lanes; do
    MSCS_GATHER=$v timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_g${v}_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_g${v}_$i.json')); print('MSCS_GATHER $v', $i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
  done
done
