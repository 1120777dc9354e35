"""CPU: the parts of bench.py's contract that do not need a GPU -- the reference arm (`--impl reference`) prints one JSON
line with the agreed keys, honours --steps/--warmup, and never touches CUDA; the GPU arm refuses to run without a
device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "4", "--warmup", "3",
                        "--cpu-lattice", "40"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "FSP Action() GB/s" and j["unit"] == "GB/s"
    assert j["steps"] == 4 and j["warmup"] == 3 and j["higher_is_better"] is True and j["dtype"] == "f64"
    assert j["value"] > 0 and j["cpu_baseline"]["value"] == j["value"] and j["cpu_baseline"]["kind"] == "port"
    assert j["cpu_baseline"]["cores"] >= 1 and "sample" in j["cpu_baseline"]
    assert j["e2e"] == {"value": j["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert j["gpu_launches"] == 0 and j["vs_baseline"] is None


def test_reference_arm_uses_all_cores_under_torchrun_env():
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every core it may run on
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--lattice", "40"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads(r.stdout.strip().splitlines()[-1])
    assert j["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert j["config"]["same_config"] is True and j["config"]["states"] == 40 ** 3 and j["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
