#!/bin/bash
# Staged GPU test run (used through gpurun): safe kernels first, the tcgen05 kernels in their own
# processes under a hard timeout so a hang cannot take the box down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
echo "== stage 1: sampling / gather / SIMT" | tee gpurun_out/stage1.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 200 \
  -k "sampling or gather or simt" >> gpurun_out/stage1.log 2>&1
echo "stage1 exit $?" | tee -a gpurun_out/stage1.log
tail -5 gpurun_out/stage1.log
echo "== stage 2: tcgen05 kernels, tiny cases" | tee gpurun_out/stage2.log
timeout -s KILL 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 120 \
  -k "similarity_kernels and tc" >> gpurun_out/stage2.log 2>&1
echo "stage2 exit $?" | tee -a gpurun_out/stage2.log
tail -30 gpurun_out/stage2.log
echo "== stage 3: end to end" | tee gpurun_out/stage3.log
timeout -s KILL 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 120 \
  -k "end_to_end or properties" >> gpurun_out/stage3.log 2>&1
echo "stage3 exit $?" | tee -a gpurun_out/stage3.log
tail -30 gpurun_out/stage3.log
