#!/bin/bash
mkdir -p gpurun_out
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
for sg in 600 900; do
MSCS_FWD_STAGGER=$sg MSCS_LIB=$PWD/$PKG/libmscs_trace.so timeout -s KILL 200 python tools/trace_fwd.py > gpurun_out/trace_fwd_g$sg.txt 2>&1
echo "trace stagger $sg exit $?"; grep -v Warn gpurun_out/trace_fwd_g$sg.txt | head -8
done
for i in 1 2; do
  for sg in 0 300 600 900 1200; do
    MSCS_FWD_STAGGER=$sg timeout -s KILL 200 python bench.py --steps 60 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_stag${sg}_$i.json 2> /dev/null
    python -c "
import json
d=json.load(open('gpurun_out/bench_stag${sg}_$i.json')); print('stagger $sg', $i, round(d['ms_per_step'],4), {k: round(x,4) for k,x in d['roofline']['stage_ms'].items()}, d['detail']['loss'])"
  done
done
