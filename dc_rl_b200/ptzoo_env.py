"""`SustainDCPettingZooEnv`: the reference's PettingZoo ParallelEnv surface (reference
harl/envs/sustaindc/sustaindc_ptzoo.py:5-101) over the CUDA-backed `SustainDC`.  Same constructor argument, attributes
(`possible_agents`, `agents`, `observation_spaces`, `action_spaces`, `share_observation_space`, `metadata`) and
`reset` / `step` return values, duck-typed: pettingzoo itself is not imported (it is not needed to BE a ParallelEnv for
callers that only use this surface, and it is not installed in this image).  For throughput use `CudaShareVecEnv`.
"""
import numpy as np

from ._lib import N_AGENTS, OBS_DIM, SHARE_DIM
from .sustaindc_env import SustainDC
from .vec_env import Box


class SustainDCPettingZooEnv:
    metadata = {"render.modes": []}

    def __init__(self, env_config, device=0, lib=None):
        if not env_config.get("partial_obs", True):
            raise NotImplementedError("Fully observable states are no longer supported. Please set 'partial_obs' to True.")   # :19
        self.env = SustainDC(env_config, device=device, lib=lib)
        self.possible_agents = self.env.agents
        self.agents = self.env.agents
        self.observation_spaces = dict(zip(self.possible_agents, self.env.observation_space))
        self.action_spaces = dict(zip(self.possible_agents, self.env.action_space))
        if env_config.get("nonoverlapping_shared_obs_space", False):
            self.share_observation_space = {a: Box(-2.0, 2.0, (SHARE_DIM,)) for a in self.possible_agents}                # :30-31
        else:
            self.share_observation_space = {a: Box(0.0, 1.0, (OBS_DIM * N_AGENTS,)) for a in self.possible_agents}          # :33-43

    def observation_space(self, agent):
        return self.observation_spaces[agent]

    def action_space(self, agent):
        return self.action_spaces[agent]

    def reset(self, seed=None, options=None):
        """Returns the observation dict only, like the reference (:58-63)."""
        if seed is not None:
            np.random.seed(seed)
        return self.env.reset()

    def step(self, actions):
        obs, rewards, dones, truncateds, infos = self.env.step(actions)
        sel = lambda d: {a: d[a] for a in self.possible_agents}          # noqa: E731
        return sel(obs), sel(rewards), sel(dones), sel(truncateds), sel(infos)

    def render(self, mode="human"):
        raise NotImplementedError("rendering is not part of the batched simulation path")

    def close(self):
        self.env.close()
