#!/bin/bash
# Quick check: parity tests, then bench with and without programmatic dependent launch.
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -2 gpurun_out/pytest_gpu.log
for p in 1 0; do
  MSCS_PDL=$p timeout -s KILL 100 python bench.py --steps 60 --warmup 10 --no-cpu-baseline > gpurun_out/bench_q$p.json 2> gpurun_out/bench_q$p.err
  echo "bench MSCS_PDL=$p exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_q$p.json'));print(d['ms_per_step'], {k: round(v,4) for k,v in d['roofline']['stage_ms'].items()}, d['e2e']['ms_per_step'])"
done
