#!/bin/bash
# A/B of programmatic dependent launch (MSCS_PDL): parity tests under both settings, bench under both (twice, interleaved).
mkdir -p gpurun_out
for p in 1 0; do
  MSCS_PDL=$p timeout -s KILL 300 python -m pytest tests -m gpu -x -q --timeout 200 > gpurun_out/pytest_pdl$p.log 2>&1
  echo "pytest MSCS_PDL=$p exit $?"; tail -2 gpurun_out/pytest_pdl$p.log
done
for rep in 1 2; do for p in 1 0; do
  MSCS_PDL=$p timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_pdl${p}_$rep.json 2> gpurun_out/bench_pdl${p}_$rep.err
  echo "bench MSCS_PDL=$p rep $rep exit $?"
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pdl${p}_$rep.json"))
print("  ms/step", round(d["ms_per_step"],4), {k: round(v,4) for k,v in d["roofline"]["stage_ms"].items()})
PY
done; done
MSCS_PDL=1 timeout -s KILL 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --layout nhwc > gpurun_out/bench_pdl1_nhwc.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/bench_pdl1_nhwc.json'));print('nhwc pdl1 ms/step', d['ms_per_step'], d['roofline']['stage_ms'])"
MSCS_PDL=1 timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
