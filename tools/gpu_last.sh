#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 60 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench.json'));print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
timeout -s KILL 25 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --layout nhwc > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
echo "bench nhwc exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_nhwc.json'));print(d['ms_per_step'], d['roofline']['stage_ms'])"
