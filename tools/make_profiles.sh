#!/bin/bash
# Turns the outputs of tools/gpu_round_end.sh (gpurun_out/) into the tracked evidence files of round R (default r02).
# usage: bash tools/make_profiles.sh [r02]
R=${1:-r02}
G=gpurun_out
set -e
python - <<EOF
import json
main = json.loads([l for l in open("$G/bench.json") if l.startswith("{")][-1])
json.dump(main, open("profiles/${R}_bench_cfg2.json", "w"), indent=1)
ref = json.loads([l for l in open("$G/bench_reference.json") if l.startswith("{")][-1])
json.dump(ref, open("profiles/${R}_bench_reference.json", "w"), indent=1)
print("main arm: %.4f ms/step, %.3e pairs/s, roofline frac %.3f; reference arm: %.3e pairs/s" %
      (main["ms_per_step"], main["value"], main["roofline"]["frac"], ref["value"]))
EOF
cp $G/launches.csv profiles/${R}_launches_cfg2.csv
{
  echo "# Round ${R#r} -- kernels of one cfg-2 step (ncu \`gpu__time_duration.sum\`, \`--clock-control none\`)"
  echo
  echo "Command: \`ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 40 --csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pooled\` (tools/gpu_round_end.sh; raw list: \`${R}_launches_cfg2.csv\`; table by tools/launch_table.py)."
  echo "Per-launch times are cold-cache and serialised (ncu also serialises the programmatic dependent launches and the side stream): compare shares, not absolutes.  \`k_mt_stream\` runs on a side stream in the real step (on the SM the persistent kernels leave free); the \`FillFunctor\` launches are autograd's \`ones_like\` of the loss and the zero fill of the gradient-row accumulators."
  echo
  python tools/launch_table.py $G/launches.csv
} > profiles/${R}_step_breakdown.md
{
  echo "# Round ${R#r} -- \`ncu --set full --clock-control none --import-source on\` of the hot kernels (cfg-2, one step)"
  echo
  echo "Command: tools/gpu_round_end.sh (\`-k regex:k_sim_|k_gather|k_dense|k_dx_rows|k_fy_select|k_label_hist -s 42 -c 8\` around \`python bench.py --steps 2 --warmup 6 --no-cpu-baseline --no-pooled\`); table by tools/ncu_summary.py from the raw page."
  echo "Numbers taken under the profiler are not bench values (the tensor kernels ran at 1.6 GHz here, 1.96 GHz in the bench loop)."
  echo
  python tools/ncu_summary.py $G/prof_full_raw.csv
} > profiles/${R}_ncu_kernels.md
{
  echo "# Round ${R#r} -- warp-stall shares of the tensor kernels (source page of the same capture)"
  echo
  python tools/stall_summary.py $G/prof_full_source.csv
} > profiles/${R}_stall_summary.md
{ tail -4 $G/pytest_gpu.log; grep -c "PASSED\|passed" $G/pytest_gpu.log >/dev/null; } > profiles/${R}_gpu_tests.txt
grep -i "random case\|worst\|rel err\|cos" $G/pytest_gpu.log | tail -100 >> profiles/${R}_gpu_tests.txt || true
python tools/sass_evidence.py > profiles/${R}_sass_evidence.md 2>/dev/null || true
ls -la profiles/${R}_*
