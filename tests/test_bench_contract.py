"""The JSON contract of bench.py: the lines recorded on the GPU boxes (profiles/r1_scaling.json) and
the reference arm run here on the CPU carry every key the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e"}


def _check_common(line):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["unit"] == "rays/s" and line["higher_is_better"] is True and line["scaling"] == "weak"
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "f32"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["value"] > 0 and line["ms_per_step"] > 0


def test_recorded_bench_lines_follow_the_contract():
    rec = json.load(open(os.path.join(ROOT, "profiles", "r1_scaling.json")))
    lines = {k: v for k, v in rec.items() if isinstance(v, dict)}
    assert {"n1", "n2_slab", "n4_slab", "n8_slab"} <= set(lines)
    for name, line in lines.items():
        _check_common(line)
        assert line["gpu_launches"] > 0
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(line["clocks"])
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = line["roofline"]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm"
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
        assert line["e2e"]["value"] <= line["value"] * 1.02          # copies + read-back are not free
        # value = global rays per step / device time
        assert abs(line["value"] - line["config"]["global_batch"] / (line["ms_per_step"] * 1e-3)) < 1e-3 * line["value"]
    n1 = lines["n1"]
    assert n1["n_gpus"] == 1 and n1["roofline"]["frac"] >= 0.6           # north-star target
    assert {"value", "unit", "cores", "kind", "sample"} <= set(n1["cpu_baseline"])
    assert n1["cpu_baseline"]["kind"] == "port"
    for n in (2, 4, 8):
        assert lines[f"n{n}_slab"]["n_gpus"] == n
        assert lines[f"n{n}_slab"]["config"]["global_batch"] == 5000 * n


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` on the smallest workload (CPU port of the whole step)."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--workload", "small", "--steps", "1", "--warmup", "1", "--cpu-rays", "256",
                          "--no-ref-pytorch"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(line)
    assert line["impl"] == "reference"
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "rays/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
