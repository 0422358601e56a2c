"""bench.py's JSON-line contract, on the arm that runs without a GPU (--impl reference)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-n", "512", "--cpu-n2", "2048", "--cpu-burn2", "8", "--cpu-steps2", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [l for l in r.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "particle_obs_updates_per_sec" and j["unit"] == "updates/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["cpu_baseline"]["kind"] in ("port", "reference") and j["cpu_baseline"]["cores"] >= 1
    assert j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["vs_baseline"] is None and j["dtype"] == "f64" and j["data"] == "synthetic"
    # the CPU arm says which population it really ran, and carries the second (larger) sample
    assert j["config"]["n_particles_run"] == 512 and j["cpu_baseline_out_of_cache"]["value"] > 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--cpu-n", "512"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_committed_ncu_traffic_is_wired():
    sys.path.insert(0, ROOT)
    import bench
    t, src = bench.ncu_traffic("k_instantiate")
    assert t and t > 1e9 and "ncu" in src
    assert bench.ncu_traffic("k_nonexistent")[0] is None
