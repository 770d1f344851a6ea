"""bench.py's JSON contract, checked on the CPU: the reference arm prints one valid line (bounded sample of the
reference's own C), and the committed GPU bench record carries every key the driver reads."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_lib as R  # noqa: E402

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config"}


@pytest.mark.skipif(not R.available(), reason="oracle/_ref/libns_ref.so not built")
def test_reference_arm_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
                        "--budget", "3", "--grid", "64"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["metric"] == "timesteps_per_sec" and d["unit"] == "steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["dtype"] == "f64" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
    # every timed sample is a full RK4Step: the line reports what was actually timed, nothing is extrapolated
    assert d["extrapolated"] is False and d["steps"] >= 1 and "full RK4Step at 64^3" in cb["sample"]
    assert d["steps_requested"] == 1


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.parametrize("name", ["r01_bench_final_1gpu.json", "r02_bench_1gpu.json"])
def test_committed_gpu_bench_record_has_the_contract_keys(name):
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip("%s not recorded yet" % name)
    d = json.loads([l for l in open(path) if l.startswith("{")][-1])
    assert BASE_KEYS <= set(d)
    assert d["n_gpus"] == 1 and d["config"]["N"] == 512 and d["dtype"] == "f64" and d["data"] == "synthetic"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and 0 < e["value"] < d["value"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert d["gpu_launches"] > 0 and d["steps"] >= 1 and d["warmup"] >= 3
    c = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c)
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"]))
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] > 0
    if name.startswith("r02"):
        assert d["parity"]["ok"] is True and d["parity"]["max_rel_err"] < 1e-12
        assert "save_interval_%d" % d["steps"] in e and "step_frac_204S" not in r
