"""bench.py contract on a machine without a GPU: the reference arm (`--impl reference`) runs the reference's CPU path
on a small workload and prints exactly one JSON line with the keys the driver reads."""
import json
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1_sphere_box_128", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(ROOT))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"] == "c1_sphere_box_128"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_workload_table_covers_baseline_configs():
    sys.path.insert(0, str(ROOT))
    import bench

    names = set(bench.WORKLOADS)
    assert {"c1_sphere_box_128", "c2_sd_obj_512", "c3_many1024_1024", "c4_mandelbulb_2048", "c5_animated_1024"} <= names
    assert bench.DEFAULT_WORKLOAD == "c3_many1024_1024"
