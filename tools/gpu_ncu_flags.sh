#!/bin/bash
# per-kernel durations of the tensor kernels under MSCS_DEBUG_FLAGS experiments (ncu launch list)
mkdir -p gpurun_out
: > gpurun_out/ncu_flags.log
for f in "$@"; do
  MSCS_DEBUG_FLAGS=$f timeout -s KILL 200 ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg --clock-control none -k regex:k_sim_ -s 9 -c 3 --csv \
    python tools/stage_times.py cfg2 2>/dev/null | grep -E "k_sim" | awk -F'","' -v f=$f '{print f, $5, $(NF-2), $NF}' >> gpurun_out/ncu_flags.log
done
cat gpurun_out/ncu_flags.log
