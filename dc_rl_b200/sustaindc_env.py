"""`SustainDC`: the reference's single-env Gymnasium surface (reference sustaindc_env.py:91-238, 436-621) over a
1-env CUDA engine.  Same constructor config keys, same dict-keyed reset()/step() returns; the simulation itself runs in
libsdc_b200.so.  For throughput use `CudaShareVecEnv` (N envs per launch) -- this class exists so that single-env
callers (scripts, notebooks, PettingZoo-style wrappers) keep working unchanged.
"""
import numpy as np

from .vec_env import AGENTS, OBS_WIDTH, CudaShareVecEnv, Box, Discrete

DEFAULT_CONFIG = {          # reference sustaindc_env.py:38-80
    "agents": ["agent_ls", "agent_dc", "agent_bat"], "location": "ny", "cintensity_file": "NYIS_NG_&_avgCI.csv",
    "weather_file": "USA_NY_New.York-Kennedy.epw", "workload_file": "Alibaba_CPU_Data_Hourly_1.csv",
    "datacenter_capacity_mw": 1, "timezone_shift": 0, "days_per_episode": 7, "max_bat_cap_Mw": 2,
    "dc_config_file": "dc_config.json", "individual_reward_weight": 0.8, "flexible_load": 0.1,
    "ls_reward": "default_ls_reward", "dc_reward": "default_dc_reward", "bat_reward": "default_bat_reward",
    "evaluation": False, "actions_are_logits": False,
}
DO_NOTHING = {"agent_ls": 1, "agent_dc": 1, "agent_bat": 2}       # reference utils/base_agents.py:16,47,73


class SustainDC:
    def __init__(self, env_config, device=0, lib=None):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(env_config)
        if cfg.get("month") is None:
            raise TypeError("env_config['month'] is required (sustaindc_env.py:150 computes self.month + 1)")
        self.env_config = cfg
        self.agents = list(cfg["agents"])
        self._vec = CudaShareVecEnv(dict(cfg, nonoverlapping_shared_obs_space=True), 1, seed=cfg.get("seed", 0),
                                    months=[int(cfg["month"])], device=device, lib=lib)
        self.observation_space = [Box(-2.0, 2.0, (OBS_WIDTH[a],)) if a != "agent_dc" else Box(-5.0e9, 5.0e9, (OBS_WIDTH[a],))
                                  for a in self.agents]
        self.action_space = [Discrete(3) for _ in self.agents]
        self.infos = {}
        self._post_reset_obs = None

    def seed(self, seed=None):
        self.seed_value = seed

    def _obs_dict(self, rows):
        return {a: rows[i, :OBS_WIDTH[a]].copy() for i, a in enumerate(AGENTS) if a in self.agents}

    def reset(self, seed=None, options=None):
        """Returns the obs dict only, like the reference (sustaindc_env.py:531)."""
        if self._post_reset_obs is not None:          # the device already reset this env at the end of the last episode
            obs, self._post_reset_obs = self._post_reset_obs, None
            return obs
        obs, _, _ = self._vec.reset()
        return self._obs_dict(obs[0])

    def step(self, action_dict):
        a = np.array([[int(action_dict.get(ag, DO_NOTHING[ag])) for ag in AGENTS]], np.int32)
        obs, _, rew, dones, infos, _ = self._vec.step(a)
        done = bool(dones[0, 0])
        row = infos[0][0]
        info = row.to_dict()
        cur = info.pop("original_obs") if done else obs[0]
        info.pop("original_state", None); info.pop("original_avail_actions", None)
        if done:
            self._post_reset_obs = self._obs_dict(obs[0])
        obs_d = self._obs_dict(np.asarray(cur))
        rew_d = {ag: float(rew[0, i, 0]) for i, ag in enumerate(AGENTS) if ag in self.agents}
        term = {ag: False for ag in self.agents}; term["__all__"] = False
        trunc = {ag: done for ag in self.agents}; trunc["__all__"] = done      # episode end == truncation (:713-718)
        self.infos = {ag: dict(info) for ag in self.agents}
        self.infos["__common__"] = dict(info)
        return obs_d, rew_d, term, trunc, self.infos

    def close(self):
        self._vec.close()
