#!/bin/bash
# Full GPU check used through gpurun: parity tests, bench, ncu launch list, one --set full capture.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_sim_|k_gather|k_scatter|k_fy_select|k_label_hist" -s 40 -c 14 \
  -o gpurun_out/prof_sim python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
