#!/bin/bash
# the random-configuration GPU parity test without pytest start-up (a few seconds of box time)
cd "$(dirname "$0")/.." && mkdir -p gpurun_out
timeout ${2:-7} env MSCS_GPU_RANDOM=${1:-10} python -c "
import sys; sys.path[:0] = ['.', 'tests']
import test_zz_gpu_random as t
t.N_CASES = int('${1:-10}')
t.test_random_configs_vs_oracle()
print('RANDOM-GPU-OK')
" > gpurun_out/random_gpu.log 2>&1
echo "exit $?" >> gpurun_out/random_gpu.log
tail -5 gpurun_out/random_gpu.log
