"""bench.py's JSON contract on the CPU-runnable arm (--impl reference)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "0", "--ref-seconds", "1.0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "instances/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("batched L-BFGS instances/sec")
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and "instances" in cb["sample"]
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"]


def test_non_zero_rank_of_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=60, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_algorithmic_bytes_model():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    # 1024*(26K - 110) bytes for K >= 10 iterations at d = 128, m = 10, fp64
    for K in (10, 11, 640):
        assert bench.algorithmic_bytes(np.array([K])) == 1024 * (26 * K - 110)
    assert bench.algorithmic_bytes(np.array([3])) == 8 * 128 * (6 * 3 + 2 * (0 + 1 + 2))


def test_reference_arm_ignores_torchrun_omp_num_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arms set their thread count explicitly."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "0", "--ref-seconds", "0.5"],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == cores and d["cpu_port_same_cores"]["cores"] == cores
    assert d["config"]["host_threads"] == cores and d["config"]["omp_num_threads_env"] == "1"
