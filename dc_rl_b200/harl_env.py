"""HARL-facing entry points: drop-in replacements for the names `harl.runners` import from
`harl.utils.envs_tools` (reference harl/utils/envs_tools.py:49-103) and for `HARLSustainDCEnv`
(reference harl/envs/sustaindc/harlsustaindc_env.py:10-210).

    from dc_rl_b200.harl_env import make_train_env, make_eval_env     # same signatures as the reference

`n_threads` becomes the number of envs batched on the GPU.  Per-env month and seed follow the reference's rules
(month: env_args['month'] if given, else rank % 12 for rank < 12, else rank % 3 + 5; seed: seed + rank*1000 for
training, seed*50000 + rank*10000 for evaluation).
"""
import numpy as np

from ._lib import N_AGENTS, OBS_DIM
from .vec_env import AGENTS, CudaShareVecEnv


def _device_from(env_args):
    return int(env_args.get("device", 0))


def make_train_env(env_name, seed, n_threads, env_args):
    if env_name != "sustaindc":
        print("Can not support the " + env_name + "environment.")
        raise NotImplementedError
    return CudaShareVecEnv(env_args, n_threads, seed=seed, device=_device_from(env_args), lib=env_args.get("_lib"))


def make_eval_env(env_name, seed, n_threads, env_args):
    if env_name != "sustaindc":
        print("Can not support the " + env_name + "environment.")
        raise NotImplementedError
    ids = np.arange(n_threads)
    seeds = (int(seed) * 50000 + ids * 10000).astype(np.uint64)
    return CudaShareVecEnv(env_args, n_threads, seeds=seeds, device=_device_from(env_args), lib=env_args.get("_lib"))


class HARLSustainDCEnv:
    """Single-env HARL adapter (lists per agent), for callers that build their own vec-env around it."""

    def __init__(self, env_args, device=0, lib=None):
        self.env_args = env_args
        self._vec = CudaShareVecEnv(env_args, 1, device=device, lib=lib)
        self.n_agents = N_AGENTS
        self.agents = list(AGENTS)
        self.observation_space = self._vec.observation_space
        self.share_observation_space = self._vec.share_observation_space
        self.action_space = self._vec.action_space
        self.discrete = True
        self._seed = 0
        self.cur_step = 0

    def seed(self, seed):
        self._seed = seed

    def reset(self):
        self._seed += 1
        self.cur_step = 0
        obs, s_obs, avail = self._vec.reset()
        return list(obs[0]), list(s_obs[0]), [list(a) for a in avail[0]]

    def step(self, actions):
        obs, s_obs, rew, dones, infos, avail = self._vec.step(np.asarray(actions).reshape(1, N_AGENTS))
        row = infos[0]
        o, s = obs[0], s_obs[0]
        if dones[0, 0]:                      # like the reference adapter: no auto-reset at this level
            o, s = row[0]["original_obs"], row[0]["original_state"]
        info = [row[a].to_dict() for a in range(N_AGENTS)]
        return list(o), list(s), [[float(rew[0, a, 0])] for a in range(N_AGENTS)], list(dones[0]), info, [list(a) for a in avail[0]]

    def get_avail_actions(self):
        return [[1] * 3 for _ in range(N_AGENTS)]

    def close(self):
        self._vec.close()
