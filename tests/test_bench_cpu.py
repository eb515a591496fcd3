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
                          "--warmup", "3", "--py-seconds", "0"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
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


def test_both_arms_print_the_same_config():
    sys.path.insert(0, ROOT)
    import bench

    c = bench.headline_config("C2")
    assert c == {"workload": bench.WORKLOADS["C2"]["desc"], "envs_per_gpu": 4096}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": headline_config(args.workload)') == 2  # the b200 arm and the reference arm


def test_replica_plan_never_revisits_a_replica_inside_l2():
    sys.path.insert(0, ROOT)
    import bench

    for steps in (1, 3, 20, 30, 200, 2000):
        for per_step in (4096 * 1006, 65536 * 1006, 65536 * 3646, 262144 * 28430):
            R, L = bench.replica_plan(steps, per_step)
            assert L % steps == 0 and L % R == 0 and R * per_step >= 2.5 * bench.L2_BYTES or R == 1 and per_step >= 2.5 * bench.L2_BYTES


def test_traffic_is_read_from_profiles():
    sys.path.insert(0, ROOT)
    import bench

    t, src = bench.ncu_traffic_bytes("C5a")
    assert src.startswith("profiles/") and 10e9 < t < 12e9  # rgb: ~11.1 GB per launch, no wasted traffic


def test_python_reference_timing_runs():
    """Row d' of the coverage table: the unmodified reference, single env and per-core vector stand-in."""
    sys.path.insert(0, ROOT)
    from oracle.ref_shim import reference_available

    if not reference_available():
        import pytest

        pytest.skip("reference not installed (baseline/_ref) or mounted")
    from oracle.ref_python_bench import per_core_vector_rate, single_env_rate

    r1, n1 = single_env_rate(dict(reward_step=True, advanced_clears=True), budget_s=0.6)
    rv, nv = per_core_vector_rate(dict(reward_step=True, advanced_clears=True), workers=2, budget_s=0.5)
    assert r1 > 1000 and n1 > 0 and rv > 1000 and nv > 0


def test_profile_readers_find_the_round_profiles():
    """roofline.traffic and roofline.issue come from the committed ncu summaries, never from constants in bench.py."""
    import bench

    for name in ("C2", "C3", "C4", "C5a", "C5b", "C2_T32", "C3_T32"):
        traffic, src = bench.ncu_traffic_bytes(name)
        assert traffic and traffic > 0 and src.startswith("profiles/r"), name
        issue = bench.ncu_issue_metrics(name)
        assert issue and 0 < issue["issue_active_pct"] <= 100 and issue["warp_instructions_per_launch"] > 0, name
    assert bench.ncu_traffic_bytes("no_such_workload") == (None, None)
    assert bench.ncu_issue_metrics("no_such_workload") is None
