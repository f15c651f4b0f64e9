#!/bin/bash
mkdir -p gpurun_out
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
MSCS_GPU_RANDOM=80 timeout -s KILL 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu_r2m.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu_r2m.log
MSCS_LIB=$PWD/$PKG/libmscs_trace1.so timeout -s KILL 200 python tools/trace_fwd1.py > gpurun_out/trace_fwd_sweep1_b.txt 2>&1
echo "trace exit $?"; grep -v Warn gpurun_out/trace_fwd_sweep1_b.txt | tail -12
for i in 1 2; do
    timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_r2m_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_r2m_$i.json')); print($i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
done
MSCS_FWD_TIMELINE=1 timeout -s KILL 200 python tools/stage_times.py 2>&1 | tail -1
