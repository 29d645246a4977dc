"""`bench.py --impl reference` (task section 4): the CPU implementation of the path timed on the host cores, one JSON line
with the native arm's metric / unit / config; under a multi-rank launch only rank 0 runs and prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(env_extra, *args):
    env = dict(os.environ, **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *args], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=600)


def test_reference_arm_prints_one_contract_line():
    r = run({}, "--steps", "2", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["value"] == d["e2e"]["value"] == d["cpu_baseline"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "1048576 envs" in d["cpu_baseline"]["sample"]
    assert "1,048,576 drones per GPU" in d["config"]["workload"] and d["gpu_launches"] == 0
    # same config object as the native arm prints (the driver compares them): no sample key, full batch
    import bench
    assert d["config"] == bench.workload_config(1)


def test_reference_arm_runs_on_rank_0_only():
    r = run({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, "--gpus", "2", "--steps", "2", "--warmup", "1")
    assert r.returncode == 0 and r.stdout.strip() == ""
