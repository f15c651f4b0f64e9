#!/bin/bash
mkdir -p gpurun_out
MSCS_GPU_RANDOM=80 timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2l.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2l.log
for i in 1 2; do
    timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_r2l_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_r2l_$i.json')); print($i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
done
MSCS_FWD_TIMELINE=1 timeout -s KILL 200 python tools/stage_times.py 2>&1 | tail -2
