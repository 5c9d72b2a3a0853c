"""bench.py contract (CPU part): the reference arm runs without a GPU and prints ONE JSON line with the keys the driver reads;
the B200 arm refuses to run without a CUDA device instead of falling back."""
import json
import subprocess
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--model", "pythia-70m", "--steps", "1", "--warmup", "1",
                          "--cpu-sample-tokens", "32", "--config1-tokens", "8,16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-1500:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, res.stdout
    d = json.loads(lines[0])
    # BASELINE.json configs[0] (Pythia-160m fp32 on the host) at the reference-faithful OMP_NUM_THREADS=1 and at all cores
    c1 = d["cpu_baseline"]["config1"]
    assert "pythia-160m" in c1["workload"] and c1["omp_num_threads_1"]["cores"] == 1 and c1["omp_num_threads_1"]["tokens_per_s"] > 0
    assert c1["all_cores"]["cores"] >= 1 and c1["all_cores"]["tokens_per_s"] > 0
    assert d["impl"] == "reference" and d["metric"] == "train_tokens_per_s" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_uses_all_cores_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; rank 0 of the reference arm must still use every core it may run on."""
    import os

    env = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", OMP_NUM_THREADS="1")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--model", "pythia-70m", "--steps", "1",
                          "--warmup", "0", "--cpu-sample-tokens", "16", "--no-config1"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert res.returncode == 0, res.stderr[-1500:]
    d = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0)) and d["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_silently():
    import os

    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--model", "pythia-70m"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_b200_arm_needs_a_gpu():
    if torch.cuda.is_available():
        return
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--steps", "1", "--warmup", "0", "--model", "pythia-70m", "--no-cpu-baseline"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode != 0 and not any(ln.startswith("{") for ln in res.stdout.splitlines())
