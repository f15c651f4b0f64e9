#!/bin/bash
mkdir -p gpurun_out
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
MSCS_FWD_POLY=0 MSCS_LIB=$PWD/$PKG/libmscs_trace.so timeout -s KILL 200 python tools/trace_fwd.py > gpurun_out/trace_fwd_poly0.txt 2>&1
echo "trace exit $?"; grep -v Warn gpurun_out/trace_fwd_poly0.txt
MSCS_FWD_POLY=1 MSCS_LIB=$PWD/$PKG/libmscs_trace.so timeout -s KILL 200 python tools/trace_fwd.py > gpurun_out/trace_fwd_poly1.txt 2>&1
echo "trace exit $?"; grep -v Warn gpurun_out/trace_fwd_poly1.txt | head -12
