#!/bin/bash
# A/B experiments through MSCS_DEBUG_FLAGS (used through gpurun): per-stage device times of cfg2
mkdir -p gpurun_out
: > gpurun_out/exp.log
for f in "$@"; do
  MSCS_DEBUG_FLAGS=$f timeout -s KILL 120 python tools/stage_times.py cfg2 >> gpurun_out/exp.log 2>&1
done
cat gpurun_out/exp.log
