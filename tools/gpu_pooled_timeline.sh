#!/bin/bash
# per-rank forward timeline of the pooled cfg-5 step at N GPUs; with "prof" as the 2nd argument the profiling build
# (make prof) also prints the per-CTA spans of the sweeps.  usage (under gpurun --gpus N): bash tools/gpu_pooled_timeline.sh N [prof]
N=${1:-8}
mkdir -p gpurun_out
PKG=$(ls -d eccv2022*_b200)
if [ "$2" = prof ]; then export MSCS_LIB=$PWD/$PKG/libmscs_prof.so; fi
MSCS_FWD_TIMELINE=1 timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/pooled_timeline.py > gpurun_out/pooled_timeline_${N}gpu$2.txt 2>&1
echo "exit $?"; grep "rank [0-9]" gpurun_out/pooled_timeline_${N}gpu$2.txt | sort
