"""CPU: the JSON-line contract of bench.py.  The reference arm (`--impl reference`: the reference's own loss files on
the host cores -- oracle/_ref or /root/reference, else the torch port of the oracle) runs here for one bounded step; the
main arm needs a B200, so its contract is checked on the line committed under profiles/ (written by the same bench.py
on the GPU box)."""
import json
import os
import subprocess
import sys

import pytest

from helpers import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


@pytest.mark.timeout(300)
def test_reference_arm_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=280, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["unit"] == "anchor-pairs/s" and d["higher_is_better"] is True
    assert d["config"]["workload"].startswith("cfg2: HRNet-W48 Cityscapes")
    cb = d["cpu_baseline"]
    from oracle import ref_loader
    assert cb["kind"] == ("reference" if ref_loader.find_root() else "port")
    assert cb["cores"] == os.cpu_count() and cb["value"] == d["value"] > 0
    assert "images of the cfg2 inputs" in cb["sample"]
    assert d["steps"] == 1 and d["warmup"] == 0          # exactly what was asked for, on a sample sized to the budget
    assert set(d["config"]) == {"workload", "layout", "per_gpu_batch", "l2", "parallelism"}
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None
    # rank != 0 of a torchrun launch exits 0 without work or output
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=60, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_main_arm_line_has_the_contract_keys():
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_cfg2.json")))
    assert BASE_KEYS | {"roofline", "clocks", "cpu_baseline"} <= set(d)
    assert d["dtype"] == "bf16" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["n_gpus"] == 1
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    # achieved = algorithmic FLOPs per launch / measured launch time (4*C per anchor pair backward)
    pairs, C = d["detail"]["anchor_pairs_per_step"], 256       # (config is identical on both arms: run facts live in detail)
    assert rf["algorithmic_flops_per_launch"] == 4 * C * pairs
    assert abs(rf["achieved"] - 4 * C * pairs / (rf["launch_ms"] * 1e-3) / 1e12) < 1e-6 * rf["achieved"]
    assert abs(d["value"] - pairs / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 5e8 and e["d2h_bytes_per_step"] == 4 and 0 < e["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["gpu_launches"] == 15 * d["steps"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
