"""bench.py contract (CPU part): the reference arm prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"] == "KP2DTiny-S frames/s @240x320" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["scaling"] == "weak"
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and "workload" in d["config"]
    # the config object is the native arm's, verbatim (the driver compares the two arms' configs)
    assert set(d["config"]) == {"workload", "batch_per_gpu", "global_batch", "parallelism", "conv_backend", "l2"}
    assert d["config"]["parallelism"] == "frame-dp1" and d["config"]["batch_per_gpu"] == 256
    cb = d["cpu_baseline"]
    staged = os.path.isdir(os.path.join(REPO, "oracle", "_ref", "src", "kp2dtiny"))
    assert cb["kind"] == ("reference" if staged else "port")  # the reference's own modules when oracle/_ref is staged
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps",
                        "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=REPO, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
