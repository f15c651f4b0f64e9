#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
MSCS_FWD_TIMELINE=1 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/pooled_timeline.py > gpurun_out/pooled_timeline_${N}gpu.txt 2>&1
echo "exit $?"; grep "^rank" gpurun_out/pooled_timeline_${N}gpu.txt
