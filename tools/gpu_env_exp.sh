#!/bin/bash
# A/B over an arbitrary environment variable: gpu_env_exp.sh VAR v1 v2 ...
mkdir -p gpurun_out
var=$1; shift
for v in "$@"; do
  echo -n "$var=$v "
  env $var=$v timeout -s KILL 120 python tools/stage_times.py cfg2 2>&1 | tail -1
done
