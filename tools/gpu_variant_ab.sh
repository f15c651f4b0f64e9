#!/bin/bash
# A/B of library variants on one GPU: bash tools/gpu_variant_ab.sh name1 name2 ...  ("default" = libmscs.so,
# otherwise libmscs_<name>.so built by `make variant NAME=<name> DEFS=...`; "<name>@VAR=value" also sets an environment
# switch for that run).  Two interleaved rounds of bench.py.
mkdir -p gpurun_out
PKG=$(ls -d eccv2022*_b200)
for round in 1 2; do
  for v in "$@"; do
    lib=${v%%@*}; envset=""; if [ "$lib" != "$v" ]; then envset=${v#*@}; fi
    if [ "$lib" = default ]; then unset MSCS_LIB; else export MSCS_LIB=$PWD/$PKG/libmscs_$lib.so; fi
    env $envset python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-pooled > gpurun_out/bench_${v}_$round.json 2> gpurun_out/bench_${v}_$round.err || tail -5 gpurun_out/bench_${v}_$round.err
    python - <<EOF
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${v}_$round.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print("$v", $round, "ms/step %.4f" % d["ms_per_step"], {k: round(x, 4) for k, x in r["stage_ms"].items()}, "fwdfrac %.3f" % r["fwd"]["frac"], flush=True)
except Exception as e:
    print("$v", $round, "FAILED", e)
EOF
  done
done
unset MSCS_LIB
