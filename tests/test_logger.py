"""The vectorised SustainDCLogger (dc_rl_b200/logger.py) against the scalars the REFERENCE logger wrote for the same inputs
(tests/golden/logger_golden.json, minted by oracle/make_golden_r2.py running harl/envs/sustaindc/sustaindc_logger.py and
harl/common/base_logger.py on seeded oracle roll-outs).  Three feeding paths: lists of dicts (the reference's own loop),
InfoBatch columns, and the device-side accumulators (`attach`); the last two replay the roll-outs on the engine -- the serial
hostsim build in the CPU suite, libsdc_b200.so under `-m gpu`."""
import json
import os
import sys

import numpy as np
import pytest

from conftest import lib_params, resolve_lib
from helpers import GOLDEN

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


class Writer:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, value, step):
        self.scalars[tag] = [float(value), int(step)]


class CriticBuffer:
    def get_mean_rewards(self):
        return 0.25


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(GOLDEN, "logger_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def rollouts(golden):
    import make_golden_r2
    rec = {}
    steps = make_golden_r2.logger_rollout_inputs(golden["spec"], rec)
    return steps, rec


def _logger(golden, tmp_path):
    from dc_rl_b200.logger import SustainDCLogger
    spec = golden["spec"]
    K = len(spec["envs"])
    algo_args = {"train": {"n_rollout_threads": K, "episode_length": spec["steps"], "num_env_steps": 10 * K * spec["steps"]},
                 "eval": {"n_eval_rollout_threads": K}}
    w = Writer()
    lg = SustainDCLogger({"env": "sustaindc", "algo": "happo", "exp_name": "golden"}, algo_args, {"location": "ny"}, 3, w, str(tmp_path))
    return lg, w, K


def _compare(got, want, rel, skip=()):
    assert set(got) == set(want), set(got) ^ set(want)
    for tag, (v, step) in want.items():
        assert got[tag][1] == step, tag
        if tag in skip:
            continue
        assert abs(got[tag][0] - v) <= rel * max(1.0, abs(v)), (tag, got[tag][0], v)


def test_dict_infos_reproduce_the_reference_logger(golden, rollouts, tmp_path):
    steps, _ = rollouts
    lg, w, K = _logger(golden, tmp_path)
    lg.init(10)
    lg.episode_init(1)
    for data in steps:
        lg.per_step(data)
    lg.episode_log([{"policy_loss": 0.5}, {"policy_loss": 0.25}, {"policy_loss": 0.125}], {"value_loss": 1.5}, None, CriticBuffer())
    _compare(w.scalars, golden["train"], 1e-12)
    w.scalars = {}
    lg.eval_init()
    for data in steps:
        lg.eval_per_step((None, None, data[2], data[3], data[4], None))
        for k in range(K):
            if data[3][k].all():
                lg.eval_thread_done(k)
    lg.eval_log(K)
    _compare(w.scalars, golden["eval"], 1e-12)
    assert abs(lg.avg_eval_episode_reward - golden["avg_eval_episode_reward"]) <= 1e-12
    lg.close()


@pytest.mark.parametrize("kind", lib_params())
@pytest.mark.parametrize("attached", [False, True])
def test_infobatch_and_device_accumulators_reproduce_the_reference_logger(golden, rollouts, tmp_path, kind, attached):
    """The same roll-outs replayed on the engine (per-env location / month, staged episodes): the logger fed with InfoBatch
    columns, or attached to the device accumulators, writes the reference logger's scalars (fp32 info columns: 1e-6; the
    attached HVAC mean / max / p90 come from the 4096-bin device histogram: one bin)."""
    from dc_rl_b200.vec_env import CudaShareVecEnv
    from replay import location_traces
    lib = resolve_lib(kind)
    steps, rec = rollouts
    spec = golden["spec"]
    K, t_ep = len(spec["envs"]), spec["days"] * 96
    locs = [e[0] for e in spec["envs"]]
    args = {"location": locs, "days_per_episode": spec["days"], "traces": {l: location_traces(l) for l in set(locs)},
            "nonoverlapping_shared_obs_space": True}
    v = CudaShareVecEnv(args, K, months=[e[1] for e in spec["envs"]], lib=lib)
    by_step = {}
    for ep in rec["episodes"]:
        by_step.setdefault(ep[0], []).append(ep)

    def stage(first_step):
        for _, k, day, hour, temp, wetb, tmin, tmax in by_step.get(first_step, []):
            v.engine.stage_episode([k], [day], [hour], temp[None], wetb[None], [tmin], [tmax])
    stage(0)
    v.reset()
    lg, w, _ = _logger(golden, tmp_path)
    if attached:
        lg.attach(v)
    lg.init(10)
    lg.episode_init(1)
    for s, data in enumerate(steps):
        stage(s + 1)                         # episodes that start after this step (auto-reset inside the step)
        obs, share, rew, dones, infos, avail = v.step(rec["actions"][s].reshape(K, 3, 1))
        assert np.array_equal(dones, data[3])
        assert np.max(np.abs(rew - data[2])) <= 1e-4
        lg.per_step((obs, share, rew, dones, infos, avail, None, None, None, None, None))
    lg.episode_log([{"policy_loss": 0.5}, {"policy_loss": 0.25}, {"policy_loss": 0.125}], {"value_loss": 1.5}, None, CriticBuffer())
    hist_tags = ("metrics/Average HVAC Power on use", "metrics/Max HVAC Power on use", "metrics/Percentile 90% HVAC Power on use")
    _compare(w.scalars, golden["train"], 2e-6 if not attached else 1e-5, skip=("train/average_step_rewards",) + (hist_tags if attached else ()))
    assert abs(w.scalars["train/average_step_rewards"][0] - golden["train"]["train/average_step_rewards"][0]) <= 1e-3
    if attached:
        width = v.engine.hvac_histogram()[1] / 4096
        for tag in hist_tags:
            assert abs(w.scalars[tag][0] - golden["train"][tag][0]) <= 1.01 * width, (tag, w.scalars[tag][0], golden["train"][tag][0])
    lg.close()
    v.close()
