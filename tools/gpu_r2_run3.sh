#!/bin/bash
# Round 2, third GPU call: leak fix + packed forward epilogue: parity, mixed-format MMA probe, host segments,
# A/B of the polynomial share of the forward epilogue (MSCS_FWD_POLY).
mkdir -p gpurun_out
timeout 60 ./tools/mix_probe > gpurun_out/mix_probe.txt 2>&1; echo "mix probe exit $?"; cat gpurun_out/mix_probe.txt
MSCS_GPU_RANDOM=80 timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 900 -k "not cfg5" > gpurun_out/pytest_gpu_r2c.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu_r2c.log
timeout -s KILL 200 python tools/stage_times.py > gpurun_out/stage_times_r2c.txt 2>&1
echo "stage times exit $?"; tail -4 gpurun_out/stage_times_r2c.txt
for i in 1 2; do
  for v in 1 0 2 3; do
    MSCS_FWD_POLY=$v timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_poly${v}_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_poly${v}_$i.json')); print('poly $v', $i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
  done
done
