"""bench.py pieces that run without a GPU: the byte accounting and the CPU (`--impl reference`) arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_survey():
    sys.path.insert(0, ROOT)
    import bench

    kws = {k: v["kw"] for k, v in bench.WORKLOADS.items()}
    assert bench.algorithmic_bytes(kws["C2"]) == 1006
    assert bench.algorithmic_bytes(kws["C5b"]) == 3646
    assert bench.algorithmic_bytes(kws["C4"]) == 28430
    assert bench.algorithmic_bytes(kws["C5a"]) == 84878


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "20",
                          "--warmup", "3"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    d = json.loads(out[0])
    assert d["impl"] == "reference" and d["metric"] == "env-steps/sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "5", "--warmup", "3"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
