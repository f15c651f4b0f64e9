#!/bin/bash
mkdir -p gpurun_out
MSCS_FWD_TIMELINE=1 timeout -s KILL 200 python tools/stage_times.py > gpurun_out/stage_times_r2k.txt 2>&1
echo "stage times exit $?"; tail -6 gpurun_out/stage_times_r2k.txt
