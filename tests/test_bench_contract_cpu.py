"""bench.py's reference arm (the reference algorithm on the host cores) prints exactly ONE JSON line with the contract's keys;
the helpers that label workloads and summarise clock samples behave."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "images/sec at 640x640 bs32 WeDetect-Base" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == dict(value=d["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "configs[1]" in d["config"]["workload"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_workload_labels_and_clock_summary():
    sys.path.insert(0, ROOT)
    import bench
    import types
    a = types.SimpleNamespace(size="base", batch=32, res=640, classes=80)
    assert bench.workload_names(a) == ("images/sec at 640x640 bs32 WeDetect-Base", "WeDetect-Base bs32/GPU 640x640 K=80 (BASELINE configs[1])")
    a = types.SimpleNamespace(size="large", batch=16, res=800, classes=1203)
    assert "configs[2]" in bench.workload_names(a)[1]
    a = types.SimpleNamespace(size="base", batch=8, res=640, classes=80)
    assert "side measurement" in bench.workload_names(a)[1]
    c = bench.ClockSampler(0, None)
    c.rows = [["1700", "1965", "950.0", "Not Active", "Not Active", "Not Active", "Active"], ["1800", "1965", "960.5", "Not Active", "Not Active", "Not Active", "Active"],
              ["1750", "1965", "955.0", "Not Active", "Not Active", "Not Active", "Not Active"]]
    s = c.summary()
    assert s["sm_mhz"] == 1750.0 and s["sm_min_mhz"] == 1700.0 and s["sm_max_mhz"] == 1965.0 and s["reasons"] == ["sw_power_cap"] and s["samples"] == 3
    assert s["power_w_max"] == 960.5
