"""bench.py contract on the CPU side: the reference arm (`--impl reference`) runs without a GPU and prints one JSON line
with the keys the driver reads; the GPU arm refuses to run without CUDA (no CPU fallback)."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                          "--warmup", "1", "--cpu-steps", "200"], capture_output=True, text=True, timeout=600, cwd=REPO)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "env-steps/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["n_gpus"] == 1 and line["scaling"] == "weak" and line["data"] == "synthetic"
    assert line["config"]["workload"].startswith("configs[2]")
    cb = line["cpu_baseline"]
    live = os.path.isfile(os.path.join(REPO, "baseline", "_ref", "sustaindc_env.py"))     # the unmodified reference tree, when shipped
    assert cb["kind"] == ("reference" if live else "port") and cb["cores"] >= 1 and cb["value"] == line["value"]
    assert ("baseline/_ref" if live else "OracleEnv.step") in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, cwd=REPO, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_gpu_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--steps", "1"], capture_output=True, text=True, timeout=300,
                         cwd=REPO)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr
