#!/bin/bash
# multi-GPU: pooled-mode parity over real NVLink ranks + the bench line with its pooled record.  usage: gpu_r2_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
CHECK_MODE=${2:-cfg5}
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/pooled_check.py $CHECK_MODE > gpurun_out/pooled_check_${N}gpu.txt 2>&1
echo "pooled_check x$N exit $?"; grep -v "Warn\|warn" gpurun_out/pooled_check_${N}gpu.txt | tail -30
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench x$N exit $?"; tail -c 2500 gpurun_out/bench_${N}gpu.json; tail -5 gpurun_out/bench_${N}gpu.err
