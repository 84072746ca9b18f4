"""Pins oracle/sdc_oracle.py against the golden vectors minted from the LIVE reference
(oracle/make_golden.py). Bar: observations bit-identical after the fp32 cast; rewards and info to 1e-12."""
import random

import numpy as np
import pytest

import sdc_oracle
from dc_rl_b200.info_layout import INFO_COLUMNS, info_dict_to_row
from helpers import kat, load_traj, oracle_traces, rel_err, reward_methods_of, traj_cfg

AG = ("agent_ls", "agent_dc", "agent_bat")
FULL = ["ny_m0_s0", "ny_m3_s1", "az_m6_s2", "wa_m9_s3", "ny_m6_dc25x200",
        "ny_m6_tz5",                                    # timezone_shift = 5
        "ny_m2_altA", "az_m8_altB", "wa_m4_altC"]       # alternate reward methods (utils/reward_creator.py:133-318)


def _make(g):
    cfg = traj_cfg(g)
    dc_cfg = None
    if "dc_geometry" in cfg:                    # builder-authored geometry (dc_config.synthetic_dc_config)
        from dc_rl_b200.dc_config import synthetic_dc_config
        nested = synthetic_dc_config(*cfg["dc_geometry"])
        dc_cfg = {k: v for section in nested.values() for k, v in section.items()}     # the oracle takes the flat form
    return sdc_oracle.OracleEnv(oracle_traces(cfg["location"], cfg.get("timezone_shift", 0)), cfg["location"], cfg["month"],
                                cfg["days_per_episode"], dc_cfg=dc_cfg, reward_methods=reward_methods_of(cfg))


def _check_reset(g, k, obs):
    assert np.array_equal(obs["agent_ls"], g["reset_obs_ls"][k])
    assert np.array_equal(obs["agent_dc"], g["reset_obs_dc"][k])
    assert np.array_equal(obs["agent_bat"], g["reset_obs_bat"][k])


@pytest.mark.parametrize("name", FULL)
def test_seeded_trajectory_matches_live_reference(name):
    """Seeded mode: the oracle consumes `random` / `np.random` exactly like the reference."""
    g = load_traj(name)
    env = _make(g)
    seed = int(g["seed"][0])
    random.seed(seed); np.random.seed(seed)
    obs = env.reset()
    k = 0
    _check_reset(g, k, obs)
    assert (env.day, env.hour) == (g["reset_day"][k], g["reset_hour"][k])
    worst_r = worst_i = 0.0
    for s in range(int(g["n_steps"][0])):
        a = [int(np.random.randint(3)) for _ in range(3)]
        assert a == list(g["actions"][s])
        obs, rew, term, info = env.step(*a)
        assert np.array_equal(obs["agent_ls"], g["obs_ls"][s]), s
        assert np.array_equal(obs["agent_dc"], g["obs_dc"][s]), s
        assert np.array_equal(obs["agent_bat"], g["obs_bat"][s]), s
        worst_r = max(worst_r, rel_err(rew, g["rewards"][s]))
        worst_i = max(worst_i, rel_err(info_dict_to_row(info), g["info"][s]))
        assert term == bool(g["trunc"][s])
        if term:
            obs = env.reset()
            k += 1
            _check_reset(g, k, obs)
    assert worst_r <= 1e-12 and worst_i <= 1e-12, (worst_r, worst_i)


def test_replay_mode_matches_seeded():
    """Replay mode (injected start + weather windows) reproduces the same trajectory without RNG."""
    g = load_traj("wa_m9_s3")
    env = _make(g)
    k = 0
    def inject(k):
        env.inject_episode(int(g["reset_day"][k]), int(g["reset_hour"][k]), g["reset_temp"][k], g["reset_wetb"][k],
                           float(g["reset_t_min30"][k]), float(g["reset_t_max30"][k]))
    inject(0)
    obs = env.reset()
    _check_reset(g, 0, obs)
    for s in range(int(g["n_steps"][0])):
        obs, rew, term, info = env.step(*[int(x) for x in g["actions"][s]])
        assert np.array_equal(obs["agent_ls"], g["obs_ls"][s]), s
        assert rel_err(rew, g["rewards"][s]) <= 1e-12
        if term:
            k += 1
            inject(k)
            _check_reset(g, k, env.reset())


def test_long_run_reward_history_saturates():
    """11 000 steps: the 10 000-sample reward window fills and rolls (utils/reward_creator.py:5)."""
    g = load_traj("ny_m6_long")
    env = _make(g)
    seed = int(g["seed"][0])
    random.seed(seed); np.random.seed(seed)
    env.reset()
    worst = 0.0
    for s in range(int(g["n_steps"][0])):
        a = [int(np.random.randint(3)) for _ in range(3)]
        _, rew, term, info = env.step(*a)
        worst = max(worst, rel_err(rew, g["rewards"][s]), rel_err(info["bat_total_energy_with_battery_KWh"], g["energy"][s]))
        if term:
            env.reset()
    assert len(env.history) == 10000
    assert worst <= 1e-12, worst


def test_sizing_known_answers():
    k = kat()["sizing"]
    for loc in ("NY", "AZ", "WA"):
        dc, c = sdc_oracle.size_datacenter(loc)
        assert rel_err(dc.ctafr, k[loc]["ctafr"]) <= 1e-14
        assert rel_err(dc.ct_fan_ref_p, k[loc]["ct_fan_ref_p"]) <= 1e-14
        assert rel_err(c["power_lb_kw"], k[loc]["power_lb_kw"]) <= 1e-14
        assert rel_err(c["power_ub_kw"], k[loc]["power_ub_kw"]) <= 1e-14
        assert rel_err(c["bat_capacity"], k[loc]["max_battery_energy_mwh"]) <= 1e-14
    assert sdc_oracle.MONTH_INIT_DAY == kat()["init_day"]


def test_dc_model_known_answers():
    dc, _ = sdc_oracle.size_datacenter("NY")
    for row in kat()["dc_model"]:
        cpu, fan, out = dc.it_model(row["load"], row["sp"])
        assert rel_err(cpu, row["rack_cpu"]) <= 1e-14 and rel_err(fan, row["rack_fan"]) <= 1e-14
        assert rel_err(out, row["rack_out"]) <= 1e-14
        t_ret = dc.return_temp(out)
        ct, q, comp, cw, ctp = dc.hvac(row["sp"], t_ret, row["amb"], sum(cpu) + sum(fan))
        assert rel_err([ct, q, comp, cw, ctp], [row["ct"], row["crac_load"], row["comp"], row["cw_pump"], row["ct_pump"]]) <= 1e-14
        assert rel_err(dc.water_usage(t_ret, row["sp"], row["twb"]), row["water"]) <= 1e-14


def test_chiller_known_answers():
    for row in kat()["chiller"]:
        assert rel_err(sdc_oracle.chiller_power(row["cap"], row["load"], row["amb"]), row["power"]) <= 1e-14


def test_normalize_energy_known_answers():
    from collections import deque
    k = kat()["normalize_energy"]
    h = deque(maxlen=10000)
    for v, z in zip(k["values"], k["z"]):
        h.append(v)
        assert abs(sdc_oracle.normalize_energy(h, v) - z) <= 1e-13
