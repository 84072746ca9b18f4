"""TEST INFRASTRUCTURE ONLY -- mints the golden vectors under tests/golden/ by running the LIVE
reference (read-only /root/reference) in this container through oracle/live_ref.py.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz and kat.json

The reference ships no tests, goldens or KATs of its own, so these files are the
pin for oracle/sdc_oracle.py, and through it for the CUDA path.  Contents:
  loc_<loc>.npz    hourly input columns as parsed by the reference managers (workload cpu_load,
                   avg_CI, EPW dry bulb / RH / pressure; utils/managers.py:168-174,345-351,521-528)
  traj_<name>.npz  step-by-step trajectories of sustaindc_env.SustainDC (reset/step) under fixed seeds
  kat.json         known-answer values of individual reference functions (IT/HVAC model, chiller,
                   battery, reward normaliser, sizing)
"""
import json
import os
import random
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)
sys.path.insert(0, _HERE)
sys.path.insert(0, _REPO)

import live_ref  # noqa: E402
from dc_rl_b200.info_layout import INFO_COLUMNS, info_dict_to_row  # noqa: E402

GOLDEN = os.path.join(_REPO, "tests", "golden")
AGENTS = ("agent_ls", "agent_dc", "agent_bat")


def dump_locations(locations=("ny", "az", "wa")):
    live_ref.import_reference()
    import pandas as pd
    from utils.utils_cf import obtain_paths
    root = live_ref.REFERENCE_ROOT
    cpu = pd.read_csv(root + "/data/Workload/Alibaba_CPU_Data_Hourly_1.csv")["cpu_load"].values[:8760].astype(float)
    for loc in locations:
        ci_loc, wea_file = obtain_paths(loc)
        ci = pd.read_csv(root + f"/data/CarbonIntensity/{ci_loc}_NG_&_avgCI.csv")["avg_CI"].values[:8760].astype(float)
        wea = pd.read_csv(root + f"/data/Weather/{wea_file}", skiprows=8, header=None).values
        np.savez_compressed(
            os.path.join(GOLDEN, f"loc_{loc}.npz"),
            cpu_load=cpu, avg_ci=ci,
            dry_bulb=wea[:, 6].astype(float), rel_hum=wea[:, 8].astype(float), pressure=wea[:, 9].astype(float),
            source=np.array([f"Alibaba_CPU_Data_Hourly_1.csv|{ci_loc}_NG_&_avgCI.csv|{wea_file}"]))
        print("wrote loc", loc)


def _reset_record(env, t_ep):
    wm, cm = env.weather_m, env.ci_m
    t0 = int(wm.time_step)
    return dict(day=int(env.t_m.day), hour=int(env.t_m.hour), t0=t0,
                temp=np.array(wm.temperature_data[t0:t0 + t_ep + 18], dtype=np.float64),
                wetb=np.array(wm.wet_bulb_data[t0:t0 + t_ep + 18], dtype=np.float64),
                t_min30=float(np.min(wm.temperature_data[t0:t0 + 2880])),
                t_max30=float(np.max(wm.temperature_data[t0:t0 + 2880])),
                ci_min30=float(np.min(cm.carbon_smooth[t0:t0 + 2880])),
                ci_max30=float(np.max(cm.carbon_smooth[t0:t0 + 2880])))


def record_trajectory(name, cfg, seed, n_steps, compact=False, dc_geometry=None):
    """dc_geometry = (rows, racks_per_row, cpus_per_rack): run the reference on the builder-authored geometry of
    dc_rl_b200.dc_config.synthetic_dc_config (BASELINE config 3: 25 racks x 40 CPUs) instead of utils/dc_config.json.
    The reference reader joins its utils/ directory with `dc_config_file`, so an absolute path selects our JSON."""
    live_cfg = dict(cfg)
    if dc_geometry is not None:
        import tempfile
        from dc_rl_b200.dc_config import synthetic_dc_config
        tmp = tempfile.NamedTemporaryFile("w", suffix=".json", delete=False)
        # the reference reader wants every key of its own file: start from it and overlay the synthetic geometry
        with open(os.path.join(live_ref.REFERENCE_ROOT, "utils", "dc_config.json")) as f:
            full = json.load(f)
        for section, values in synthetic_dc_config(*dc_geometry).items():
            full.setdefault(section, {}).update(values)
        json.dump(full, tmp)
        tmp.close()
        live_cfg["dc_config_file"] = tmp.name
        cfg = dict(cfg, dc_geometry=list(dc_geometry))
    env = live_ref.fresh_env(live_cfg)
    t_ep = cfg["days_per_episode"] * 96
    random.seed(seed)
    np.random.seed(seed)
    obs = env.reset()
    resets = [_reset_record(env, t_ep)]
    reset_obs = [[np.asarray(obs[a], np.float32) for a in AGENTS]]
    reset_at = [0]
    actions = np.zeros((n_steps, 3), np.int8)
    rewards = np.zeros((n_steps, 3), np.float64)
    trunc = np.zeros(n_steps, np.bool_)
    energy = np.zeros(n_steps, np.float64)
    obs_ls = np.zeros((n_steps, 26), np.float32)
    obs_dc = np.zeros((n_steps, 14), np.float32)
    obs_bat = np.zeros((n_steps, 13), np.float32)
    info = np.zeros((n_steps, len(INFO_COLUMNS)), np.float64)
    for s in range(n_steps):
        a = [int(np.random.randint(3)) for _ in range(3)]
        actions[s] = a
        o, r, term, tr, inf = env.step(dict(zip(AGENTS, a)))
        assert not any(term.values())
        rewards[s] = [r[k] for k in AGENTS]
        trunc[s] = tr["__all__"]
        energy[s] = inf["agent_ls"]["bat_total_energy_with_battery_KWh"]
        if not compact:
            obs_ls[s], obs_dc[s], obs_bat[s] = (o[k] for k in AGENTS)
            info[s] = info_dict_to_row(inf["agent_ls"])
        if tr["__all__"]:
            obs = env.reset()
            resets.append(_reset_record(env, t_ep))
            reset_obs.append([np.asarray(obs[k], np.float32) for k in AGENTS])
            reset_at.append(s + 1)
    dc = env.dc_env
    cfg_o = dc.DC_Config
    out = dict(
        cfg_json=np.array([json.dumps(cfg)]), seed=np.array([seed]), n_steps=np.array([n_steps]),
        actions=actions, rewards=rewards, trunc=trunc, energy=energy,
        reset_at=np.array(reset_at), reset_day=np.array([r["day"] for r in resets]),
        reset_hour=np.array([r["hour"] for r in resets]), reset_t0=np.array([r["t0"] for r in resets]),
        reset_temp=np.stack([r["temp"] for r in resets]), reset_wetb=np.stack([r["wetb"] for r in resets]),
        reset_t_min30=np.array([r["t_min30"] for r in resets]), reset_t_max30=np.array([r["t_max30"] for r in resets]),
        reset_ci_min30=np.array([r["ci_min30"] for r in resets]), reset_ci_max30=np.array([r["ci_max30"] for r in resets]),
        reset_obs_ls=np.stack([o[0] for o in reset_obs]), reset_obs_dc=np.stack([o[1] for o in reset_obs]),
        reset_obs_bat=np.stack([o[2] for o in reset_obs]),
        rack_full=np.array([float(r.full_load_pwr[0]) for r in dc.dc.racks_list]),
        rack_idle=np.array([float(r.idle_pwr[0]) for r in dc.dc.racks_list]),
        rack_ncpu=np.array([int(r.num_CPUs) for r in dc.dc.racks_list]),
        ctafr=np.array([cfg_o.CT_REFRENCE_AIR_FLOW_RATE]), ct_fan_ref_p=np.array([cfg_o.CT_FAN_REF_P]),
        power_lb_kw=np.array([dc.power_lb_kW]), power_ub_kw=np.array([dc.power_ub_kW]),
        bat_capacity=np.array([env.bat_env.battery.capacity]),
        info_columns=np.array(INFO_COLUMNS))
    if not compact:
        out.update(obs_ls=obs_ls, obs_dc=obs_dc, obs_bat=obs_bat, info=info)
    np.savez_compressed(os.path.join(GOLDEN, f"traj_{name}.npz"), **out)
    print("wrote traj", name, "steps", n_steps, "resets", len(resets))


def known_answers():
    live_ref.import_reference()
    import envs.datacenter as DC
    from envs.bat_env_fwd_view import BatteryEnvFwd
    from utils import reward_creator
    from utils.dc_config_reader import DC_Config
    from utils.make_envs_pyenv import make_dc_pyeplus_env
    from utils.utils_cf import get_init_day
    kat = {}
    # -- sizing per location (utils/make_envs_pyenv.py:149-218)
    sizing = {}
    for loc in ("NY", "AZ", "WA"):
        dc_env, _ = make_dc_pyeplus_env(month=1, location=loc, dc_config_file="dc_config.json", use_ls_cpu_load=True,
                                        add_cpu_usage=False)
        c = dc_env.DC_Config
        sizing[loc] = dict(ctafr=c.CT_REFRENCE_AIR_FLOW_RATE, ct_fan_ref_p=c.CT_FAN_REF_P,
                           power_lb_kw=dc_env.power_lb_kW, power_ub_kw=dc_env.power_ub_kW,
                           max_battery_energy_mwh=dc_env.ranges["max_battery_energy_Mwh"],
                           zone_air=list(dc_env.ranges["Zone Air Temperature(West Zone)"]))
    kat["sizing"] = sizing
    kat["init_day"] = [get_init_day(m) for m in range(12)]
    # -- IT + HVAC model (envs/datacenter.py:250-317,432-474,325-353) at NY sizing
    dc_env, _ = make_dc_pyeplus_env(month=1, location="NY", dc_config_file="dc_config.json", use_ls_cpu_load=True,
                                    add_cpu_usage=False)
    cfgo = dc_env.DC_Config
    rows = []
    rng = np.random.RandomState(7)
    cases = [(18.0, 50.0, 20.0, 15.0), (21.6, 100.0, 35.0, 25.0), (15.0, 0.0, 2.0, 1.0)]
    cases += [(float(rng.uniform(15, 21.6)), float(rng.uniform(0, 100)), float(rng.uniform(-5, 45)), float(rng.uniform(0, 30)))
              for _ in range(61)]
    for sp, load, amb, twb in cases:
        cpu, fan, out = dc_env.dc.compute_datacenter_IT_load_outlet_temp([load] * cfgo.NUM_RACKS, sp)
        ret = DC.calculate_avg_CRAC_return_temp(cfgo.RACK_RETURN_APPROACH_TEMP_LIST, out)
        p_it = sum(cpu) + sum(fan)
        _, ct, crac_load, comp, cw, ctp = DC.calculate_HVAC_power(sp, ret, amb, p_it, cfgo)
        dc_env.dc.hot_water_temp, dc_env.dc.cold_water_temp, dc_env.dc.wet_bulb_temp = ret, sp, twb
        water = float(dc_env.dc.calculate_cooling_tower_water_usage())
        rows.append(dict(sp=sp, load=load, amb=amb, twb=twb, p_it=float(p_it), t_out_mean=float(np.mean(out)),
                         t_ret=float(ret), ct=float(ct), crac_load=float(crac_load), comp=float(comp),
                         water=water, cw_pump=float(cw), ct_pump=float(ctp),
                         rack_cpu=[float(x) for x in cpu], rack_fan=[float(x) for x in fan],
                         rack_out=[float(x) for x in out]))
    kat["dc_model"] = rows
    # -- chiller alone (envs/datacenter.py:356-429), including the min-PLR cycling branch
    ch = []
    for cap, load, amb in [(2.3e6, 1.0e6, 20.0), (2.3e6, 5.0e4, 10.0), (2.3e6, 3.0e6, 40.0), (2.3e6, 0.0, 25.0),
                           (1.0e6, 2.0e5, -5.0), (2307120.481120018, 933307.154154, 20.0)]:
        ch.append(dict(cap=cap, load=load, amb=amb, power=float(DC.calculate_chiller_power(cap, load, amb))))
    kat["chiller"] = ch
    # -- battery (envs/bat_env_fwd_view.py:84-126,194-284; envs/battery_model.py:94-139)
    bat = BatteryEnvFwd({"n_fwd_steps": 8, "max_bat_cap": 4.8313716318897475, "charging_rate": 0.5,
                         "max_dc_pw_MW": 4.83, "dcload_max": 1200.0, "dcload_min": 147.0})
    bat.reset()
    rng = np.random.RandomState(11)
    seq = []
    for i in range(200):
        a = int(rng.randint(3)) if i >= 40 else 0          # charge for a while first
        dcl = float(rng.uniform(0.6, 2.4))
        ci = float(rng.uniform(150, 380))
        bat.set_dcload(dcl)
        bat.update_ci(ci, 0.0)
        _, _, _, _, inf = bat.step(a)
        seq.append(dict(a=a, dcl=dcl, ci=ci, soc=float(inf["bat_SOC"]), co2=float(inf["bat_CO2_footprint"]),
                        e=float(inf["bat_total_energy_with_battery_KWh"]), load=float(bat.battery.current_load)))
    kat["battery"] = seq
    # -- reward normaliser (utils/reward_creator.py:16-45)
    reward_creator.energy_history.clear()
    rng = np.random.RandomState(5)
    vals = (330 + 40 * rng.randn(600)).tolist()
    zs = []
    for v in vals:
        reward_creator.update_energy_history(v)
        zs.append(float(reward_creator.normalize_energy(v)))
    kat["normalize_energy"] = dict(values=vals, z=zs)
    reward_creator.energy_history.clear()
    with open(os.path.join(GOLDEN, "kat.json"), "w") as f:
        json.dump(kat, f)
    print("wrote kat.json")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if "--geometry-only" in sys.argv:       # adds the non-default geometry trajectory without rewriting the other fixtures
        record_trajectory("ny_m6_dc25x200", {"location": "ny", "month": 6, "days_per_episode": 2}, seed=5, n_steps=450, dc_geometry=(5, 5, 200))
        return
    dump_locations()
    known_answers()
    base = {"location": "ny", "month": 0, "days_per_episode": 7}
    record_trajectory("ny_m0_s0", dict(base), seed=0, n_steps=1344 + 10)
    record_trajectory("ny_m3_s1", dict(base, month=3, days_per_episode=3), seed=1, n_steps=700)
    record_trajectory("az_m6_s2", dict(base, location="az", month=6, days_per_episode=3), seed=2, n_steps=700)
    record_trajectory("wa_m9_s3", dict(base, location="wa", month=9, days_per_episode=2), seed=3, n_steps=500)
    # long run: reward history saturates at 10 000 samples (utils/reward_creator.py:5)
    record_trajectory("ny_m6_long", dict(base, month=6), seed=4, n_steps=11000, compact=True)
    # non-default geometry: 25 racks (5 x 5) of 200 CPUs, builder-authored JSON.  (BASELINE config 3 names 25 racks x 40
    # CPUs; the reference itself refuses that geometry: racks that small trip its outlet-temperature guard,
    # envs/datacenter.py:295-300.)
    record_trajectory("ny_m6_dc25x200", dict(base, month=6, days_per_episode=2), seed=5, n_steps=450, dc_geometry=(5, 5, 200))


if __name__ == "__main__":
    main()
