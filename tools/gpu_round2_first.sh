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
# A/B of the forward-epilogue variants (csrc: `make lean` before the call; sim_fwd.cu MSCS_LEAN = 1: smaller code,
# 2: + accumulator buffer released before the last chunk's math, 3: + polling waits in the MMA warp): parity under each, then 3 interleaved bench rounds
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
for v in 1 2 3; do
  [ -f $PKG/libmscs_lean$v.so ] || continue
  MSCS_LIB=$PWD/$PKG/libmscs_lean$v.so timeout -s KILL 300 python -m pytest tests -m gpu -q --timeout 200 > gpurun_out/pytest_gpu_lean$v.log 2>&1
  echo "pytest (lean$v library) exit $?"; tail -1 gpurun_out/pytest_gpu_lean$v.log
done
for i in 1 2 3; do
  for v in default lean1 lean2 lean3; do
    lib=$PWD/$PKG/libmscs_$v.so; [ $v = default ] && lib=$PWD/$PKG/libmscs.so
    [ -f $lib ] || continue
    MSCS_LIB=$lib timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${v}_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_${v}_$i.json')); print('$v', $i, d['ms_per_step'], d['roofline']['stage_ms']['sim_fwd'])"
  done
done
