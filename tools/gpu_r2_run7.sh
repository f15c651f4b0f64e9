#!/bin/bash
mkdir -p gpurun_out
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
MSCS_GPU_RANDOM=20 timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2g.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2g.log; grep "pooled cfg5" gpurun_out/pytest_gpu_r2g.log | head
MSCS_LIB=$PWD/$PKG/libmscs_trace.so timeout -s KILL 200 python tools/trace_fwd.py > gpurun_out/trace_fwd_f.txt 2>&1
echo "trace exit $?"; grep -v Warn gpurun_out/trace_fwd_f.txt | head -20
for i in 1 2; do
  for v in 0 1; do
    MSCS_FWD_POLY=$v timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_fpoly${v}_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_fpoly${v}_$i.json')); print('poly $v', $i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
  done
done
