#!/bin/bash
# One `ncu --set full` capture of the hot kernels of one cfg-2 step (7 launches), final code.
mkdir -p gpurun_out
timeout -s KILL 170 ncu --set full --clock-control none --import-source on -k regex:"k_sim_|k_gather|k_scatter|k_fy_select|k_label_hist" -s 42 -c 7 \
  -o gpurun_out/prof_final -f python bench.py --steps 2 --warmup 6 --no-cpu-baseline > gpurun_out/ncu_full_final.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full_final.log; ls -la gpurun_out/prof_final.ncu-rep
