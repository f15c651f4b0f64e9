#!/bin/bash
# What the driver runs at round end, plus the profiles that back the numbers in DESIGN.md (one B200, through gpurun):
# the -m gpu suite, smoke(), both bench arms, the ncu launch list of one step and one `--set full` capture of the hot
# kernels.  Outputs land in gpurun_out/; the summaries under profiles/ are made from them (tools/launch_table.py,
# tools/ncu_summary.py, tools/stall_summary.py).
mkdir -p gpurun_out
MSCS_GPU_RANDOM=80 timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm exit $?"; tail -c 1200 gpurun_out/bench_reference.json
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pooled > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout -s KILL 400 ncu --set full --clock-control none --import-source on \
  -k regex:"k_sim_|k_gather|k_dense|k_dx_rows|k_fy_select|k_label_hist" -s 42 -c 8 -o gpurun_out/prof_full -f \
  python bench.py --steps 2 --warmup 6 --no-cpu-baseline --no-pooled > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/prof_full.ncu-rep
ncu -i gpurun_out/prof_full.ncu-rep --page raw --csv > gpurun_out/prof_full_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_full.ncu-rep --page source --csv > gpurun_out/prof_full_source.csv 2>/dev/null
