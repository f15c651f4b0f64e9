#!/bin/bash
# First GPU call of the next round: the parity suite with the random-case test widened from the 10 cases that ran green
# on a B200 to 80 (all pinned against the live reference on the CPU side), the default bench line, and the stale
# channels-last launch list refreshed (it predates the merged work-table launch: 16 kernels per step instead of 15).
mkdir -p gpurun_out
MSCS_GPU_RANDOM=80 timeout -s KILL 400 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu_random80.log 2>&1
echo "pytest (80 random cases) exit $?"; tail -3 gpurun_out/pytest_gpu_random80.log
timeout -s KILL 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -c 3800 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv \
  --log-file gpurun_out/launches_nhwc.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --layout nhwc > gpurun_out/ncu_bench_nhwc.log 2>&1
echo "ncu launches (nhwc) exit $?"
# A/B of the lean forward epilogue (csrc: `make lean` before the call; sim_fwd.cu MSCS_LEAN): parity, then 3 bench pairs
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
if [ -f $PKG/libmscs_lean.so ]; then
  MSCS_LIB=$PWD/$PKG/libmscs_lean.so timeout -s KILL 300 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_lean.log 2>&1
  echo "pytest (lean library) exit $?"; tail -1 gpurun_out/pytest_gpu_lean.log
  for i in 1 2 3; do
    timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_default_$i.json 2> /dev/null
    MSCS_LIB=$PWD/$PKG/libmscs_lean.so timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_lean_$i.json 2> /dev/null
    python -c "
import json
for n in ('default','lean'):
    d=json.load(open('gpurun_out/bench_%s_$i.json'%n)); print(n, d['ms_per_step'], d['roofline']['stage_ms']['sim_fwd'])"
  done
fi
