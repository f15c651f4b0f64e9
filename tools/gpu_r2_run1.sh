#!/bin/bash
# Round 2, first GPU call: parity suite (80 random cases, LossWrapper + live-reference tests), default bench line,
# A/B of the MSCS_LEAN forward-epilogue variants.
mkdir -p gpurun_out
MSCS_GPU_RANDOM=80 timeout -s KILL 900 python -m pytest tests -m gpu -q --timeout 600 -k "not cfg5" > gpurun_out/pytest_gpu_r2a.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/pytest_gpu_r2a.log
timeout -s KILL 300 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err
echo "bench exit $?"; tail -c 3000 gpurun_out/bench_r2a.json; tail -3 gpurun_out/bench_r2a.err
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
for i in 1 2; do
  for v in default lean1 lean2 lean3; do
    lib=$PWD/$PKG/libmscs_$v.so; [ $v = default ] && lib=$PWD/$PKG/libmscs.so
    [ -f $lib ] || continue
    MSCS_LIB=$lib timeout -s KILL 200 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_${v}_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_${v}_$i.json')); print('$v', $i, d['ms_per_step'], d['roofline']['stage_ms'])"
  done
done
