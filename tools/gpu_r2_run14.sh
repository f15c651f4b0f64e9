#!/bin/bash
mkdir -p gpurun_out
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
MSCS_LIB=$PWD/$PKG/libmscs_trace1.so timeout -s KILL 200 python tools/trace_fwd1.py > gpurun_out/trace_fwd_sweep1.txt 2>&1
echo "trace exit $?"; grep -v Warn gpurun_out/trace_fwd_sweep1.txt | tail -30
MSCS_GPU_RANDOM=0 timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "channels_last" -s 2>&1 | grep "nhwc vs\|passed\|failed"
