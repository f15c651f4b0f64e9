#!/bin/bash
# Round-end style check through gpurun: parity tests (default = programmatic dependent launch on, and MSCS_PDL=0),
# smoke, the default bench line, the same with MSCS_PDL=0, the channels-last layout, launch list, memcheck of smoke.
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
MSCS_PDL=0 timeout -s KILL 300 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_pdl0.log 2>&1
echo "pytest MSCS_PDL=0 exit $?"; tail -1 gpurun_out/pytest_gpu_pdl0.log
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
MSCS_PDL=0 timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err
echo "bench MSCS_PDL=0 exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_pdl0.json'));print(d['ms_per_step'], d['roofline']['stage_ms'])"
timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err
echo "bench (default) 50 steps exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_pdl1.json'));print(d['ms_per_step'], d['roofline']['stage_ms'])"
timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline --layout nhwc > gpurun_out/bench_nhwc.json 2> gpurun_out/bench_nhwc.err
echo "bench nhwc exit $?"; python -c "import json;d=json.load(open('gpurun_out/bench_nhwc.json'));print(d['ms_per_step'], d['roofline']['stage_ms'])"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "ncu launches exit $?"
timeout -s KILL 150 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck.log 2>&1
echo "memcheck exit $?"; tail -2 gpurun_out/memcheck.log
