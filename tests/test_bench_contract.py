"""bench.py's reference arm (the reference's algorithm on the host cores through the oracle port) must put exactly ONE JSON line on stdout
with the keys the driver reads; everything else (layer tables, library banners) goes to stderr.  Runs on the CPU in ~30 s."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1", "--num_envs", "512"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, (out.stdout[-500:], out.stderr[-500:])
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    if "unavailable" in d:
        return
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
              "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["unit"] == "env-steps/s" and d["value"] > 0 and d["vs_baseline"] is None and d["config"]["workload"].startswith("go2 rough-terrain")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the arm runs what it prints (1024 envs were labelled 4096, warm-up 1 labelled 5): the step / warm-up counts are the ones that ran,
    # and `config` is the very object this repo's arm prints for the same command line
    sys.path.insert(0, ROOT)
    import bench
    assert d["steps"] == 2 and d["warmup"] == 1
    assert d["config"] == bench.workload_config("go2", 512, 1)
    assert "512-env workload" in d["cpu_baseline"]["sample"] and abs(d["value"] - 2 * 512 * 24 / (d["ms_per_step"] * 2e-3)) < 1e-6 * d["value"]


def test_reference_arm_says_unavailable_for_tasks_the_port_does_not_restate():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--task", "go2_moe_cts"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1 and "unavailable" in json.loads(lines[0])
