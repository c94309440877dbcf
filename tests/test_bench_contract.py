"""CPU: the bench line keeps the driver's contract -- checked on the last bench captured on a B200 (profiles/) and on a live
run of the CPU reference arm at a tiny size."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = {"metric": str, "value": float, "unit": str, "n_gpus": int, "steps": int, "warmup": int, "ms_per_step": float,
        "higher_is_better": bool, "scaling": str, "dtype": str, "data": str, "config": dict}


def _check_base(line):
    for k, t in BASE.items():
        assert k in line and isinstance(line[k], t), (k, line.get(k))
    assert "vs_baseline" in line and "workload" in line["config"]
    assert line["unit"] == "agent-timesteps/s" and line["higher_is_better"] is True and line["scaling"] == "weak"


def test_committed_b200_bench_lines_have_every_contract_key():
    ours = sorted(p for p in glob.glob(os.path.join(ROOT, "profiles", "r[0-9]*_bench.json")))
    assert ours, "no bench line captured on a B200 under profiles/"
    line = json.load(open(ours[-1]))
    _check_base(line)
    assert line["warmup"] >= 3 and line["gpu_launches"] > 0 and line["vs_baseline"] is None
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] != line["value"]
    roof = line["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in roof, k
    assert roof["bound"] in ("hbm", "tensor") and abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9
    cpu = line["cpu_baseline"]
    assert cpu["kind"] in ("port", "reference") and cpu["cores"] >= 1 and cpu["value"] > 0 and cpu["sample"]
    clocks = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(clocks)
    assert not set(clocks["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line_live():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--agents", "3", "--k", "3", "--num_gens", "2"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    _check_base(line)
    assert line["impl"] == "reference"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cpu = line["cpu_baseline"]
    assert cpu["value"] == line["value"] and cpu["cores"] >= 1 and cpu["sample"]
    staged = os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "mggan")) or os.path.isdir("/root/reference/mggan")
    # the unmodified reference trainer when a copy is present (staged by __graft_entry__.build()), else the oracle port
    assert cpu["kind"] == ("reference" if staged else "port")
    if staged:
        assert cpu["port"]["kind"] == "port" and cpu["port"]["value"] > 0
