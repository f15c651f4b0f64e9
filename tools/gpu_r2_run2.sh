#!/bin/bash
# Round 2, second GPU call: full parity suite (cfg-5 fixture pending), barrier wait profile, host profile.
mkdir -p gpurun_out
PKG=eccv2022-multi-scale-and-cross-scale-contrastive-segmentation_b200
MSCS_GPU_RANDOM=80 timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 900 -k "not cfg5" > gpurun_out/pytest_gpu_r2b.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_gpu_r2b.log
MSCS_LIB=$PWD/$PKG/libmscs_prof.so timeout -s KILL 200 python tools/wait_profile.py > gpurun_out/wait_profile_r2b.txt 2>&1
echo "wait profile exit $?"; cat gpurun_out/wait_profile_r2b.txt | grep -v Warning
timeout -s KILL 200 python tools/stage_times.py > gpurun_out/stage_times_r2b.txt 2>&1
echo "stage times exit $?"; tail -5 gpurun_out/stage_times_r2b.txt
timeout -s KILL 200 python tools/host_profile.py > gpurun_out/host_profile_r2b.txt 2>&1
echo "host profile exit $?"; head -50 gpurun_out/host_profile_r2b.txt
