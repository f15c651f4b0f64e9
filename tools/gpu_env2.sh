#!/bin/bash
# A/B over full environment assignments: gpu_env2.sh "A=1 B=2" "A=0" ...
for e in "$@"; do
  echo -n "[$e] "
  env $e timeout -s KILL 120 python tools/stage_times.py cfg2 2>&1 | grep -E "^0 |wall" | tr '\n' ' '; echo
done
