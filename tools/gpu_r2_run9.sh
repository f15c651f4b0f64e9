#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_projector.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -x -s -k "projector or sampling_bit_exact_small or end_to_end_small or gather" > gpurun_out/pytest_gpu_r2h.log 2>&1
echo "pytest exit $?"; grep -i "projector tail\|scale \|cfg2 shape\|passed\|failed\|Error" gpurun_out/pytest_gpu_r2h.log | head -30
timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_r2h.json 2> /dev/null
python -c "
import json
d=json.load(open('gpurun_out/bench_r2h.json')); print(round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
