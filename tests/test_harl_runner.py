"""BASELINE config 5 / SURVEY 8b: the UNMODIFIED reference runner (harl/runners/on_policy_ha_runner.py under baseline/_ref,
copied from the reference by oracle/setup_baseline_ref.py) drives CudaShareVecEnv through dc_rl_b200.harl_runner.install(),
which only rebinds make_train_env / make_eval_env / the logger registry entry.  On the CPU suite the env is bound to the
serial hostsim build (SDC_B200_LIB); under `-m gpu` to libsdc_b200.so."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HARL = os.path.join(REPO, "baseline", "_ref")
needs_harl = pytest.mark.skipif(not os.path.isdir(os.path.join(HARL, "harl")), reason="baseline/_ref (unmodified reference tree) not present")


def _run(extra, env, tmp_path):
    cmd = [sys.executable, "-W", "ignore", "-m", "dc_rl_b200.harl_runner", "--harl-root", HARL, "--out", str(tmp_path)] + extra
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO, env=env)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    return json.loads(out.stdout.strip().splitlines()[-1]), out.stdout


@needs_harl
@pytest.mark.parametrize("logger", ["vector", "reference"])
def test_unmodified_happo_runner_trains_on_the_batched_env(tmp_path, logger):
    """collect -> envs.step -> logger.per_step -> insert -> compute -> train of the reference runner, two iterations, with the
    vectorised logger and with the reference's own per-env logger loop reading InfoBatch rows."""
    import hostsim_build
    env = dict(os.environ, SDC_B200_LIB=hostsim_build.build())
    extra = ["--n-envs", "6", "--episode-length", "12", "--episodes", "2"] + (["--reference-logger"] if logger == "reference" else [])
    line, stdout = _run(extra, env, tmp_path)
    assert line["env_steps"] == 6 * 12 * 2 and line["env_steps_per_s_into_buffers"] > 0
    assert line["logger"] == ("dc_rl_b200.logger" if logger == "vector" else "harl.envs.sustaindc.sustaindc_logger")
    assert "updates 2/2 episodes" in stdout and "Avg Net Energy=" in stdout


@needs_harl
@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["runner", "device"])
def test_happo_rollout_on_the_cuda_library(tmp_path, mode):
    """The same on libsdc_b200.so, and the device-resident rollout (step_torch + one hand-over per episode) feeding the
    runner's unmodified compute() / train()."""
    from conftest import cuda_lib_or_skip
    cuda_lib_or_skip()
    extra = ["--n-envs", "512", "--episode-length", "16", "--episodes", "2"] + (["--device-rollout"] if mode == "device" else [])
    line, stdout = _run(extra, dict(os.environ), tmp_path)
    assert line["env_steps"] == 512 * 16 * 2 and line["env_steps_per_s_into_buffers"] > 0
    assert "updates 2/2 episodes" in stdout
