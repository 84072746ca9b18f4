"""TEST INFRASTRUCTURE: drives an Engine (CUDA library or the hostsim build) through a golden trajectory
in replay mode and compares every output with the recorded live-reference values."""
import numpy as np

from dc_rl_b200 import info_layout
from dc_rl_b200.dc_config import size_datacenter
from dc_rl_b200.engine import Engine
from dc_rl_b200.traces import LocationTraces
from helpers import GOLDEN, load_traj, reward_methods_of, traj_cfg

import functools
import os


@functools.lru_cache(maxsize=None)
def location_traces(loc, timezone_shift=0):
    return LocationTraces.from_npz(os.path.join(GOLDEN, "loc_%s.npz" % loc), loc, timezone_shift)


def scaled_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def make_engine(g, lib, n_envs=1, **kw):
    cfg = traj_cfg(g)
    dc_cfg = None
    if "dc_geometry" in cfg:                    # builder-authored geometry (dc_config.synthetic_dc_config)
        from dc_rl_b200.dc_config import synthetic_dc_config
        dc_cfg = synthetic_dc_config(*cfg["dc_geometry"])
    params, _ = size_datacenter(cfg["location"], dc_cfg)
    eng = Engine(n_envs, [location_traces(cfg["location"], cfg.get("timezone_shift", 0))], [params], months=cfg["month"],
                 days_per_episode=cfg["days_per_episode"], lib=lib, **kw)
    methods = reward_methods_of(cfg)
    if methods != ("default_ls_reward", "default_dc_reward", "default_bat_reward"):
        eng.set_reward_methods(*methods)
    return eng


def stage(eng, g, k, env_ids):
    n = len(env_ids)
    eng.stage_episode(env_ids, [int(g["reset_day"][k])] * n, [int(g["reset_hour"][k])] * n,
                      np.repeat(g["reset_temp"][k][None], n, 0), np.repeat(g["reset_wetb"][k][None], n, 0),
                      [float(g["reset_t_min30"][k])] * n, [float(g["reset_t_max30"][k])] * n)


def replay(name, lib, n_envs=1, max_steps=None, compact=False, **kw):
    """Returns dict of worst errors. All `n_envs` envs replay the same trajectory (they must agree)."""
    g = load_traj(name)
    eng = make_engine(g, lib, n_envs, **kw)
    ids = np.arange(n_envs, dtype=np.int32)
    t_ep = eng.ep_len
    stage(eng, g, 0, ids)
    obs, share = eng.reset_host()
    worst = dict(obs=0.0, rew=0.0, info=0.0, share=0.0, reset_obs=0.0, term_obs=0.0)

    def cmp_reset(k, obs):
        for a, key, w in ((0, "reset_obs_ls", 26), (1, "reset_obs_dc", 14), (2, "reset_obs_bat", 13)):
            worst["reset_obs"] = max(worst["reset_obs"], scaled_err(obs[:, a, :w], np.broadcast_to(g[key][k], (n_envs, w))))
            assert not obs[:, a, w:].any()
    cmp_reset(0, obs)
    k = 0
    in_ep = 0
    n_steps = int(g["n_steps"][0]) if max_steps is None else min(max_steps, int(g["n_steps"][0]))
    used = info_layout.INFO_K_USED
    for s in range(n_steps):
        if in_ep == t_ep - 1 and k + 1 < len(g["reset_day"]):
            stage(eng, g, k + 1, ids)
        act = np.broadcast_to(g["actions"][s].astype(np.int32), (n_envs, 3))
        obs, share, rew, done, info, term = eng.step_host(act)
        in_ep += 1
        assert bool(done[0]) == bool(g["trunc"][s]) and (done == done[0]).all()
        worst["rew"] = max(worst["rew"], scaled_err(rew, np.broadcast_to(g["rewards"][s], (n_envs, 3))))
        if not compact:
            cur = term if done[0] else obs
            for a, key, w in ((0, "obs_ls", 26), (1, "obs_dc", 14), (2, "obs_bat", 13)):
                e = scaled_err(cur[:, a, :w], np.broadcast_to(g[key][s], (n_envs, w)))
                worst["term_obs" if done[0] else "obs"] = max(worst["term_obs" if done[0] else "obs"], e)
            worst["info"] = max(worst["info"], scaled_err(info[:used].T, np.broadcast_to(g["info"][s], (n_envs, used))))
        else:
            e = scaled_err(info[info_layout.COL["bat_total_energy_with_battery_KWh"]], g["energy"][s])
            worst["info"] = max(worst["info"], e)
        if done[0]:
            k += 1
            in_ep = 0
            cmp_reset(k, obs)
        else:
            ref_share = np.concatenate([obs[:, 0, :], obs[:, 1, 11:12], obs[:, 1, 13:14], obs[:, 2, 25:26]], axis=1)
            worst["share"] = max(worst["share"], scaled_err(share, ref_share))
    worst["err_flags"] = int(np.bitwise_or.reduce(eng.read_state("err")))
    worst["engine"] = eng
    return worst
