#!/bin/bash
# Channels-last (NHWC) row gather / scatter: parity tests, bench in both layouts, launch list + one full ncu capture.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --layout nhwc > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
echo "bench nhwc exit $?"; tail -c 2500 gpurun_out/bench_nhwc.json; tail -5 gpurun_out/bench_nhwc.err
timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_nchw.json 2> gpurun_out/bench_nchw.err
echo "bench nchw exit $?"; tail -c 2500 gpurun_out/bench_nchw.json; tail -5 gpurun_out/bench_nchw.err
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv \
  --log-file gpurun_out/launches_nhwc.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --layout nhwc > gpurun_out/ncu_bench_nhwc.log 2>&1
echo "ncu launches exit $?"
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"rows_nhwc" -s 4 -c 4 \
  -o gpurun_out/prof_nhwc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --layout nhwc > gpurun_out/ncu_full_nhwc.log 2>&1
echo "ncu full exit $?"
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
