#!/bin/bash
# bench line with its pooled record at N GPUs (torchrun), then the per-rank forward timeline.
# usage (under gpurun --gpus N): bash tools/gpu_bench_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 40 --warmup 8 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "bench x$N exit $?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<EOF
import json
for l in open('gpurun_out/bench_${N}gpu.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print(json.dumps(d.get('pooled'),indent=1))
        print(d['ms_per_step'], d['value'], d['n_gpus'], d['roofline']['stage_ms'])
EOF
bash tools/gpu_pooled_timeline.sh $N
