"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU (numpy, fp64) restatement of the reference algorithm for one
SustainDC env-step / reset, written from SURVEY.md Appendix A and checked line by line against the
reference sources cited below.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this file; the product path never does.

PARITY PIN: tests/test_oracle_golden.py checks this restatement against tests/golden/traj_*.npz and
kat.json, which oracle/make_golden.py minted by running the live reference in the build container
(observations bit-identical after the fp32 cast, rewards / info to <= 1e-12 relative).  The one unpinned
input is the wet-bulb trace (psychrolib is absent everywhere; see dc_rl_b200/psychro.py).

Reference map (all paths relative to the reference root):
  traces / managers   utils/managers.py:66-88,116-147,183-185,220-244,247-314,401-483,581-666
  load shifting       envs/carbon_ls.py:63-73,76-169,172-324
  data centre         envs/dc_gym.py:142-237 ; envs/datacenter.py:31-49,65-74,157-181,250-317,325-353,
                      356-429,432-474,476-529,531-541
  sizing              utils/make_envs_pyenv.py:124-218 ; utils/dc_config_reader.py:39-145 ; utils/utils_cf.py:56-77
  battery             envs/bat_env_fwd_view.py:84-126,194-284 ; envs/battery_model.py:70-139
  reward              utils/reward_creator.py:5-130
  env                 sustaindc_env.py:266-433 (obs), 436-531 (reset), 533-737 (step)
  HARL view           harl/envs/sustaindc/harlsustaindc_env.py:42-131 ; harl/envs/env_wrappers.py:168-219
"""
import math
import random as _pyrandom
from collections import deque

import numpy as np

# ----------------------------------------------------------------------------------------------
# Default data-centre description (values of the reference's utils/dc_config.json)
# ----------------------------------------------------------------------------------------------
DEFAULT_DC_CONFIG = {
    "NUM_ROWS": 4, "NUM_RACKS_PER_ROW": 5, "CPUS_PER_RACK": 200,
    "RACK_SUPPLY_APPROACH_TEMP_LIST": [5.3] * 5 + [5.0] * 10 + [5.3] * 5,
    "RACK_RETURN_APPROACH_TEMP_LIST": [-3.7] * 5 + [-2.5] * 10 + [-3.7] * 5,
    "C_AIR": 1006, "RHO_AIR": 1.225, "CRAC_SUPPLY_AIR_FLOW_RATE_pu": 0.00005663,
    "CRAC_REFRENCE_AIR_FLOW_RATE_pu": 0.00009438, "CRAC_FAN_REF_P": 150,
    "CW_PRESSURE_DROP": 300000, "CW_WATER_FLOW_RATE": 0.0011, "CW_PUMP_EFFICIENCY": 0.87,
    "CT_PRESSURE_DROP": 300000, "CT_WATER_FLOW_RATE": 0.0011, "CT_PUMP_EFFICIENCY": 0.87,
    "CPU_POWER_RATIO_LB": [0.01, 1.00], "CPU_POWER_RATIO_UB": [0.03, 1.02],
    "IT_FAN_AIRFLOW_RATIO_LB": [0.01, 0.225], "IT_FAN_AIRFLOW_RATIO_UB": [0.225, 1.0],
    "IT_FAN_FULL_LOAD_V": 0.051, "ITFAN_REF_V_RATIO": 1.0, "ITFAN_REF_P": 10.0, "INLET_TEMP_RANGE": [16, 28],
    "DEFAULT_SERVER_POWER_CHARACTERISTICS": [[170, 20], [120, 10]] + [[130, 10]] * 11 + [[170, 10]]
    + [[130, 10]] * 2 + [[110, 10]] + [[170, 10]] * 3,
}
MONTH_INIT_DAY = [0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334]   # utils/utils_cf.py:56-77 on a 2022 index
MAX_AMBIENT = {"ny": 30.0, "az": 50.0, "wa": 20.0}                        # utils/make_envs_pyenv.py:149-157
SP_MIN, SP_MAX, SP_INIT = 15.0, 21.6, 18                                   # utils/make_envs_pyenv.py:124-126
QUEUE_MAX = 1000                                                           # sustaindc_env.py:148-149
HIST_MAX = 10000                                                           # utils/reward_creator.py:5
STEPS_PER_DAY = 96
YEAR_STEPS = 35040


# ----------------------------------------------------------------------------------------------
# Data-centre IT + HVAC model
# ----------------------------------------------------------------------------------------------
class DCModel:
    """Rack tables + curve parameters (envs/datacenter.py:31-49,65-135) for one dc_config."""

    def __init__(self, cfg=None, datacenter_capacity_mw=1):
        c = dict(DEFAULT_DC_CONFIG if cfg is None else cfg)
        self.cfg = c
        self.n_racks = c["NUM_ROWS"] * c["NUM_RACKS_PER_ROW"]
        max_w = int(datacenter_capacity_mw * 1e6 / self.n_racks)          # dc_config_reader.py:52-53
        lo, hi = c["INLET_TEMP_RANGE"]
        self.m_cpu = (c["CPU_POWER_RATIO_UB"][0] - c["CPU_POWER_RATIO_LB"][0]) / (hi - lo)      # datacenter.py:36
        self.c_cpu = c["CPU_POWER_RATIO_UB"][0] - self.m_cpu * hi                                 # :37
        self.shift_cpu = c["CPU_POWER_RATIO_LB"][1] - c["CPU_POWER_RATIO_LB"][0]                  # :39
        self.m_fan = (c["IT_FAN_AIRFLOW_RATIO_UB"][0] - c["IT_FAN_AIRFLOW_RATIO_LB"][0]) / (hi - lo)   # :46
        self.c_fan = c["IT_FAN_AIRFLOW_RATIO_UB"][0] - self.m_fan * hi                            # :47
        self.shift_fan = c["IT_FAN_AIRFLOW_RATIO_LB"][1] - c["IT_FAN_AIRFLOW_RATIO_LB"][0]        # :49
        self.full, self.idle, self.ncpu = [], [], []
        for full, idle in c["DEFAULT_SERVER_POWER_CHARACTERISTICS"][:self.n_racks]:
            # datacenter.py:65-74 -- CPUs are appended while the running sum of full-load power stays
            # below the rack cap; the CPU that reaches the cap is popped.
            n, load = 0, 0
            for _ in range(int(c["CPUS_PER_RACK"])):
                n += 1
                load += full
                if load >= max_w:
                    n -= 1
                    break
            self.full.append(float(full)); self.idle.append(float(idle)); self.ncpu.append(n)
        self.supply = [max(3.8, min(s, 5.3)) for s in c["RACK_SUPPLY_APPROACH_TEMP_LIST"]]        # :209-215
        self.ret = list(c["RACK_RETURN_APPROACH_TEMP_LIST"])
        self.ctafr = None          # CT_REFRENCE_AIR_FLOW_RATE after sizing
        self.ct_fan_ref_p = None   # CT_FAN_REF_P after sizing

    def it_model(self, load_pct, setpoint):
        """Per-rack CPU W, fan W, outlet degC (datacenter.py:157-181,250-317). Uses numpy arrays of C_r
        identical CPUs and np.sum, exactly like the reference's vectorised rack."""
        c = self.cfg
        cpu_w, fan_w, out_t = [], [], []
        for r in range(self.n_racks):
            n = self.ncpu[r]
            t_in = self.supply[r] + setpoint
            m_cpu = np.full(n, self.m_cpu); c_cpu = np.full(n, self.c_cpu)
            base = (m_cpu + 0.05) * t_in + c_cpu
            ratio = base + np.full(n, self.shift_cpu) * (load_pct / 100)
            p_cpu = np.maximum(np.full(n, self.idle[r]), np.full(n, self.full[r]) * ratio)
            v_base = np.full(n, self.m_fan) * 10 * t_in + np.full(n, self.c_fan) * 5
            v = v_base + np.full(n, self.shift_fan) * (load_pct / 20)
            p_fan = np.full(n, c["ITFAN_REF_P"]) * (v / np.full(n, c["ITFAN_REF_V_RATIO"]))
            v_fan = np.full(n, c["IT_FAN_FULL_LOAD_V"]) * v
            pc, pf = np.sum(p_cpu), np.sum(p_fan)
            power_term = (pc + pf) ** 1.096
            airflow_term = c["C_AIR"] * c["RHO_AIR"] * np.sum(v_fan) ** 0.824 * 0.526
            t_out = t_in + 1.918 * power_term / airflow_term + (-14.01)
            cpu_w.append(pc); fan_w.append(pf); out_t.append(t_out)
        return cpu_w, fan_w, out_t

    def return_temp(self, out_t):
        return sum([i + j for i, j in zip(self.ret, out_t)]) / len(self.ret)     # datacenter.py:540-541

    def hvac(self, setpoint, t_ret, ambient, p_it):
        """(CT fan W, CRAC load W, compressor W, CW pump W, CT pump W) -- datacenter.py:432-474."""
        c = self.cfg
        m_sys = c["RHO_AIR"] * c["CRAC_SUPPLY_AIR_FLOW_RATE_pu"] * p_it
        q = m_sys * c["C_AIR"] * max(0.0, t_ret - setpoint)
        comp = chiller_power(self.ct_fan_ref_p, q, ambient)
        cw = (c["CW_PRESSURE_DROP"] * c["CW_WATER_FLOW_RATE"]) / c["CW_PUMP_EFFICIENCY"]
        ctp = (c["CT_PRESSURE_DROP"] * c["CT_WATER_FLOW_RATE"]) / c["CT_PUMP_EFFICIENCY"]
        if ambient < 5:
            return 0.0, q, comp, cw, ctp
        delta = max(50 - (ambient - setpoint), 1)
        v_air = q / (c["C_AIR"] * delta) / c["RHO_AIR"]
        ct = self.ct_fan_ref_p * (min(v_air / self.ctafr, 1)) ** 3
        return ct, q, comp, cw, ctp

    @staticmethod
    def water_usage(t_ret, setpoint, wet_bulb):
        """L per 15 min (datacenter.py:325-353)."""
        w = 0.044 * wet_bulb + (0.3528 * (t_ret - setpoint) + 0.101)
        w = np.clip(w, 0, None)
        w += w * 0.01
        return np.round((w * 1000) / 4, 4)


def chiller_power(max_cooling_cap, load, ambient):
    """EnergyPlus-style electric chiller (datacenter.py:356-429)."""
    cap_c = [0.94483600, -0.05700880, 0.00185486]
    pow_c = [2.333, -1.975, 0.6121]
    flf = [0.03303, 0.6852, 0.2818]
    min_plr, max_plr = 0.05, 1.0
    d_t = (ambient - 35.0) / 2.778 - (6.67 - 35.0)
    rat = cap_c[0] + cap_c[1] * d_t + cap_c[2] * d_t ** 2
    avail = max_cooling_cap * rat if rat != 0 else 0
    fpr = pow_c[0] + pow_c[1] * rat + pow_c[2] * rat ** 2
    plr = max(min_plr, min(load / avail, max_plr)) if avail > 0 else 0
    ffl = flf[0] + flf[1] * plr + flf[2] * plr ** 2
    if avail > 0:
        opl = load / avail if load / avail < min_plr else plr
    else:
        opl = 0.0
    frac = min(1.0, opl / min_plr) if opl < min_plr else 1.0
    power = ffl * fpr * avail / 3.0 * frac
    return power if opl > 0 else 0


def size_datacenter(location, cfg=None, datacenter_capacity_mw=1):
    """Init-time sizing (utils/make_envs_pyenv.py:149-218; datacenter.py:476-529). Returns a sized DCModel
    plus the derived constants the env needs."""
    dc = DCModel(cfg, datacenter_capacity_mw)
    loc = location.lower()
    max_amb = 30.0 if "ny" in loc else 50.0 if "az" in loc else 20.0 if "wa" in loc else 50.0
    # chiller_sizing(dc_config, 15.0, 21.6, max_amb)
    cpu, fan, out = dc.it_model(100.0, SP_MAX)
    t_ret = dc.return_temp(out)
    p_it = sum(cpu) + sum(fan)
    m_sys = dc.cfg["RHO_AIR"] * dc.cfg["CRAC_SUPPLY_AIR_FLOW_RATE_pu"] * p_it
    crac_load = m_sys * dc.cfg["C_AIR"] * max(0.0, t_ret - SP_MIN)
    delta = max(50 - (max_amb - SP_MIN), 1)
    dc.ctafr = crac_load / (dc.cfg["C_AIR"] * delta) / dc.cfg["RHO_AIR"]
    dc.ct_fan_ref_p = crac_load
    # 8 x 11 sweep (make_envs_pyenv.py:168-178)
    ite, amb = [], []
    for sp in range(15, 23):
        for load in range(0, 110, 10):
            cpu, fan, out = dc.it_model(load, sp)
            ite.append(sum(cpu) + sum(fan))
            amb.append(sum(out) / len(out))
    chiller_max = chiller_power(dc.ct_fan_ref_p, max(ite), max_amb)
    max_dc_power_w = 1.1 * max(ite) + 1.1 * dc.ct_fan_ref_p + 1.1 * chiller_max
    hvac_rng = [0.0, 1.1 * dc.ct_fan_ref_p + 1.1 * chiller_max]
    it_rng = [0.9 * min(ite), 1.1 * max(ite)]
    consts = dict(
        power_lb_kw=(it_rng[0] + hvac_rng[0]) / 1e3, power_ub_kw=(it_rng[1] + hvac_rng[1]) / 1e3,   # dc_gym.py:86-87
        bat_capacity=(max_dc_power_w / 4) * (4 * 1) / 1e6,                                           # :190-197
        zone_air=[0.9 * min(amb), 1.1 * max(amb)])
    return dc, consts


# ----------------------------------------------------------------------------------------------
# Exogenous traces (managers)
# ----------------------------------------------------------------------------------------------
def interp15(hourly):
    """Hourly -> 15 min (managers.py:183-185): endpoint-inclusive linspace, clamped at the end."""
    hourly = np.asarray(hourly, dtype=float)
    x = range(0, len(hourly))
    xn = np.linspace(0, len(hourly), len(hourly) * 4)
    return np.interp(xn, x, hourly)


class Traces:
    """Per-location traces at 15-min resolution built from the hourly input columns."""

    def __init__(self, cpu_load, avg_ci, dry_bulb, wet_bulb_hourly, timezone_shift=0):
        roll = lambda x: np.roll(x, -1 * timezone_shift * 4)              # noqa: E731  managers.py:188,377,557-558
        cpu = roll(interp15(cpu_load[:8760]))
        # workload: percentile rescale + 16-tap smoothing at every reset, deterministic (managers.py:220-244,268-271)
        p5, p95 = np.percentile(cpu, 5), np.percentile(cpu, 95)
        scaled = np.clip(0.2 + ((cpu - p5) * (0.8 - 0.2) / (p95 - p5)), 0, 1)
        self.workload = np.convolve(scaled, np.ones(16) / 16, mode="same")
        ci = np.asarray(avg_ci[:8760], dtype=float)
        if np.isnan(ci).any():
            ci = np.nan_to_num(ci, nan=np.nanmean(ci))                     # managers.py:359-361
        self.ci = np.clip(roll(interp15(ci)), 0, None)                      # managers.py:417
        self.temp_base = roll(interp15(dry_bulb))                           # managers.py:550
        self.wetb_base = roll(interp15(wet_bulb_hourly))                    # managers.py:547

    @staticmethod
    def from_golden(npz, wet_bulb_fn, timezone_shift=0):
        wb = [wet_bulb_fn(t, rh / 100, p) for t, rh, p in zip(npz["dry_bulb"], npz["rel_hum"], npz["pressure"])]
        return Traces(npz["cpu_load"], npz["avg_ci"], npz["dry_bulb"], wb, timezone_shift)


def weather_reset(traces, t0, np_rng=np.random):
    """Weather_Manager.reset (managers.py:594-613): coherent noise, day roll, clip, 30-day min/max.
    Consumes np_rng exactly like the reference: normal(35040) then randint(0, 14)."""
    n = len(traces.temp_base)
    steps = np_rng.normal(loc=0, scale=1, size=n)                            # managers.py:45
    walk = np.cumsum(0.02 * steps)
    noise = 0 + (walk / np.std(walk)) * 0.75                                 # managers.py:46-48
    temp = traces.temp_base + noise
    wetb = traces.wetb_base + noise
    roll = np_rng.randint(0, 14)
    temp = np.roll(temp, roll * 96)
    wetb = np.roll(wetb, roll * 96)
    temp = np.clip(temp, 0, 45)
    wetb = np.clip(wetb, 0, 45)
    t_max, t_min = np.max(temp[t0:2880 + t0]), np.min(temp[t0:2880 + t0])
    return temp, wetb, t_min, t_max


def hour_features(hour):
    """sc_obs (managers.py:66-88), first two entries only: (cos, sin) of the rounded day fraction."""
    ang = round(hour / 24, 3) * (np.pi * 2)
    return np.cos(ang) * 0.5 + 0.5, np.sin(ang) * 0.5 + 0.5


# ----------------------------------------------------------------------------------------------
# Observation features (sustaindc_env.py:266-433)
# ----------------------------------------------------------------------------------------------
def _trend_features(values, current):
    mean = np.mean(values)
    std = np.std(values)
    grad = np.gradient(np.hstack((current, values)))
    peaks = np.where((grad[:-1] > 0) & (grad[1:] <= 0))[0]
    valleys = np.where((grad[:-1] < 0) & (grad[1:] >= 0))[0]
    t_peak = peaks[0] if len(peaks) > 0 else len(values)
    t_valley = valleys[0] if len(valleys) > 0 else len(values)
    return np.array([mean, std, (current - mean) / (std + 1e-8), t_peak / len(values), t_valley / len(values)])


def ci_features(cur, fut, past):
    sm_f = np.convolve(np.hstack((cur, fut[:16])), np.ones(4), "valid") / 4
    sm_p = np.convolve(np.hstack((past, cur)), np.ones(4), "valid") / 4
    f_slope = np.polyfit(range(len(sm_f)), sm_f, 1)[0]
    p_slope = np.polyfit(range(len(sm_p)), sm_p, 1)[0]
    return np.hstack([f_slope, p_slope, _trend_features(fut, cur)])


def temp_features(cur, nxt_n):
    slope = np.polyfit(range(len(nxt_n) + 1), np.hstack([cur, nxt_n]), 1)[0]
    return np.hstack([slope, _trend_features(nxt_n, cur)])


# ----------------------------------------------------------------------------------------------
# The env
# ----------------------------------------------------------------------------------------------
class OracleEnv:
    """One SustainDC env. `history` is private to the instance (the reference keeps one per process)."""

    def __init__(self, traces, location="ny", month=0, days_per_episode=7, dc_cfg=None,
                 py_rng=_pyrandom, np_rng=np.random, reward_methods=None):
        self.tr = traces
        self.dc, self.consts = size_datacenter(location, dc_cfg)
        self.cap = self.consts["bat_capacity"]
        self.t_ep = days_per_episode * STEPS_PER_DAY
        init_day = MONTH_INIT_DAY[month]
        self.day_range = (max(0, init_day - 7), min(364, init_day + 7))       # sustaindc_env.py:197-198
        self.py_rng, self.np_rng = py_rng, np_rng
        self.history = deque(maxlen=HIST_MAX)
        self.setpoint = SP_INIT                                              # dc_gym.py:77, survives reset()
        self.injected = None
        # (ls, dc, bat) reward method names, sustaindc_env.py:137-144 / utils/reward_creator.py:322-334
        self.reward_methods = tuple(reward_methods or ("default_ls_reward", "default_dc_reward", "default_bat_reward"))

    # -- reset ---------------------------------------------------------------------------------
    def inject_episode(self, day, hour, temp_window, wetb_window, t_min30, t_max30):
        """Replay mode: the next reset() uses these instead of drawing from the RNGs."""
        self.injected = (day, hour, np.asarray(temp_window, float), np.asarray(wetb_window, float), t_min30, t_max30)

    def reset(self):
        if self.injected is not None:
            day, hour, tw, ww, t_min, t_max = self.injected
            self.injected = None
            self.t = day * 96 + hour * 4
            self.temp = np.full(YEAR_STEPS + 32, np.nan); self.wetb = np.full(YEAR_STEPS + 32, np.nan)
            self.temp[self.t:self.t + len(tw)] = tw
            self.wetb[self.t:self.t + len(ww)] = ww
        else:
            day = self.py_rng.randint(max(0, self.day_range[0]), min(364, self.day_range[1]))   # :454
            hour = self.py_rng.randint(0, 23)                                                    # :455
            self.t = day * 96 + hour * 4
            self.np_rng.random()                               # Workload_Manager.reset draws one unused value (managers.py:260)
            self.temp, self.wetb, t_min, t_max = weather_reset(self.tr, self.t, self.np_rng)
        self.day, self.hour = day, hour
        self.t_min, self.t_max = t_min, t_max
        self.norm_temp = (self.temp - t_min) / (t_max - t_min)                                  # managers.py:608
        ci = self.tr.ci
        c_max, c_min = np.max(ci[self.t:2880 + self.t]), np.min(ci[self.t:2880 + self.t])
        self.norm_ci = (ci - c_min) / (c_max - c_min)                                            # managers.py:435-437
        self.ci_min, self.ci_max = c_min, c_max
        self.step_in_ep = 0
        self.queue = deque(maxlen=QUEUE_MAX)                                                     # carbon_ls.py:85
        self.run, self.last_delta, self.scale = 0, None, 1                                       # dc_gym.py:114-116
        self.bat_load = 0                                                                        # battery_model.py:90-91
        self.ls = dict(norm_q=0, oldest=0.0, avg=0.0, hist=np.zeros(5), overdue=0)               # carbon_ls.py:154-167
        return self._observe(soc=self.bat_load / self.cap)

    # -- observation ---------------------------------------------------------------------------
    def _observe(self, soc):
        t = self.t
        cos_h, sin_h = hour_features(self.hour)
        cur = self.norm_ci[t]
        fut = self.norm_ci[t + 1:t + 9]
        past = self.norm_ci[t - 16:t]
        f_ci = ci_features(cur, fut, past)
        w, w_next = self.tr.workload[t], self.tr.workload[t + 1]
        nt, nt_next = self.norm_temp[t], self.norm_temp[t + 1]
        f_t = temp_features(nt, self.norm_temp[t + 1:t + 17])
        ls = np.float32(np.hstack((cos_h, sin_h, cur, f_ci, self.ls["oldest"], self.ls["avg"], self.ls["norm_q"],
                                   w, nt, f_t, self.ls["hist"])))
        dc = np.float32(np.hstack((cos_h, sin_h, cur, f_ci, w, w_next, nt, nt_next)))
        bat = np.float32(np.hstack((cos_h, sin_h, cur, f_ci, w, nt, soc)))
        return {"agent_ls": ls, "agent_dc": dc, "agent_bat": bat}

    # -- one step ------------------------------------------------------------------------------
    def step(self, a_ls, a_dc, a_bat):
        t = self.t
        info = {}
        # ---- load shifting (carbon_ls.py:172-324) at date (day, hour) of index t
        w = self.tr.workload[t]
        if w < 0 or w > 1:
            raise ValueError("The workload should be between 0 and 1")
        ns = int(math.ceil(w * 0.8 * 100))
        sh = int(math.floor(w * 0.2 * 100))
        day, hour = self.day, self.hour
        age = lambda task: (day - task[0]) * 24 + (hour - task[1])          # noqa: E731
        overdue = [task for task in self.queue if age(task) > 24]
        over = len(overdue)
        cap = 90 - (ns + sh)
        otp = 0
        if cap > 0 and over > 0:
            otp = min(over, cap)
            for task in overdue[:otp]:
                self.queue.remove(task)
        cap = 90 - (ns + sh + otp)
        dropped = proc = 0
        if a_ls == 0:
            add = min(sh, QUEUE_MAX - len(self.queue))
            dropped = sh - add
            self.queue.extend([(day, hour)] * add)
            util = (otp + (sh - add)) / 100
        elif a_ls == 2:
            if cap >= 1:
                proc = min(sh, cap, len(self.queue))
                if proc == len(self.queue):
                    self.queue.clear()
                else:
                    for _ in range(proc):
                        self.queue.popleft()
                util = (sh + proc + otp) / 100
            else:
                util = (sh + otp) / 100
        else:
            util = (sh + otp) / 100
        util += ns / 100
        qlen = len(self.queue)
        if qlen > 0:
            ages = [age(task) for task in self.queue]
            oldest, avg = max(ages), sum(ages) / len(ages)
        else:
            ages, oldest, avg = [], 0.0, 0.0
        hist, _ = np.histogram(ages, bins=[0, 6, 12, 18, 24, np.inf])
        hist = hist / max(qlen, 1)
        hist[-1] = 1 if hist[-1] > 0 else 0
        self.ls = dict(norm_q=qlen / QUEUE_MAX, oldest=oldest / 24, avg=avg / 24, hist=hist, overdue=over)
        info.update(ls_original_workload=w, ls_shifted_workload=util, ls_action=a_ls, ls_norm_load_left=0,
                    ls_unasigned_day_load_left=0, ls_penalty_flag=0, ls_queue_max_len=QUEUE_MAX,
                    ls_tasks_in_queue=qlen, ls_norm_tasks_in_queue=qlen / QUEUE_MAX, ls_tasks_dropped=dropped,
                    ls_current_hour=hour, ls_tasks_processed=proc, ls_enforced=0, ls_oldest_task_age=oldest / 24,
                    ls_average_task_age=avg / 24, ls_overdue_penalty=over, ls_computed_tasks=int(util * 100),
                    ls_task_age_histogram=hist)
        # ---- data centre (dc_gym.py:142-237)
        assert 0.0 <= util <= 1.0, "CPU load out of bounds"
        delta = {0: -1, 1: 0, 2: 1}[a_dc]
        if delta == self.last_delta and a_dc != 0:
            self.run += 1
        else:
            self.run = 1
            self.scale = 1
        if self.run > 3:
            self.scale += 1
        self.setpoint += delta * self.scale
        self.setpoint = max(min(self.setpoint, SP_MAX), SP_MIN)
        self.last_delta = delta
        sp = self.setpoint
        ambient, wet_bulb = self.temp[t], self.wetb[t]
        cpu_w, fan_w, out_t = self.dc.it_model(util * 100, sp)
        for o, s in zip(out_t, self.dc.supply):
            if o - (s + sp) < 2:
                raise Exception("Sorry, no numbers below 2")               # datacenter.py:295-300
        t_ret = self.dc.return_temp(out_t)
        p_it = sum(cpu_w) + sum(fan_w)
        ct, q, comp, cw, ctp = self.dc.hvac(sp, t_ret, ambient, p_it)
        water = DCModel.water_usage(t_ret, sp, wet_bulb)
        total_kw = (p_it + ct + comp) / 1e3
        info.update(dc_ITE_total_power_kW=p_it / 1e3, dc_CT_total_power_kW=ct / 1e3,
                    dc_Compressor_total_power_kW=comp / 1e3, dc_HVAC_total_power_kW=(ct + comp) / 1e3,
                    dc_total_power_kW=total_kw, dc_crac_setpoint_delta=delta, dc_crac_setpoint=sp,
                    dc_cpu_workload_fraction=util, dc_int_temperature=np.mean(out_t),
                    dc_exterior_ambient_temp=ambient, dc_power_lb_kW=self.consts["power_lb_kw"],
                    dc_power_ub_kW=self.consts["power_ub_kw"], dc_CW_pump_power_kW=cw, dc_CT_pump_power_kW=ctp,
                    dc_water_usage=water)
        # ---- battery (bat_env_fwd_view.py:84-126,194-284; battery_model.py:94-139)
        dcl = total_kw / 1e3
        ci_now = self.tr.ci[t]
        cap_b, b = self.cap, self.bat_load
        soc = (b - 0) / (cap_b - 0)
        sig = lambda x: 1 / (1 + np.exp(-x))                                # noqa: E731
        if a_bat == 0:
            t_u = np.round(0.5 * (1 - sig(10 * (soc - 0.5))), 4) * 15 / 60
            max_c = min((cap_b / 1) * 0.1, (1 * cap_b - b) / ((1 * t_u) - (-0.04)))
            chg = min(max_c, cap_b) * 1 * t_u
            b = np.round(b + chg, 8)
            energy = dcl * 1e3 * 0.25 + chg * 1e3
            co2 = energy * ci_now
        elif a_bat == 1:
            t_u = max(0.5, 4 * sig(10 * (soc - 0.25))) * 15 / 60
            max_d = min((cap_b / 1) * 1, (b - 0 * cap_b) / (0.01 + (1 * t_u)), dcl / 4)
            b = np.round(b - (min(max_d, cap_b) * 1 * t_u), 8)
            dis = max_d * t_u if max_d < cap_b else cap_b * t_u
            assert dcl * 1e3 * 0.25 >= dis * 1e3
            energy = dcl * 1e3 * 0.25 - dis * 1e3
            co2 = max(energy, 0) * ci_now
        else:
            energy = dcl * 1e3 * 0.25
            co2 = energy * ci_now
        self.bat_load = b
        info.update(bat_action=a_bat, bat_SOC=b / cap_b, bat_CO2_footprint=co2, bat_avg_CI=ci_now,
                    bat_total_energy_without_battery_KWh=dcl * 1e3 * 0.25, bat_total_energy_with_battery_KWh=energy,
                    bat_max_bat_cap=cap_b, bat_a_t={0: "charge", 1: "discharge", 2: "idle"}[a_bat],
                    bat_dcload_min=self.consts["power_lb_kw"] / 4, bat_dcload_max=self.consts["power_ub_kw"] / 4)
        # ---- managers advance (managers.py:127-147,285-302,452-474,633-654)
        self.t = t + 1
        self.step_in_ep += 1
        self.hour += 0.25
        if self.hour >= 24:
            self.hour = 0
            self.day += 1
        terminal = self.step_in_ep >= self.t_ep
        obs = self._observe(soc=b / cap_b)
        # ---- rewards (reward_creator.py:16-130; ls first: it alone appends to the history)
        norm_ci_next = self.norm_ci[self.t + 1]
        info.update(outside_temp=self.temp[self.t], day=self.day, hour=self.hour, norm_CI=norm_ci_next,
                    forecast_CI=self.norm_ci[self.t + 1:self.t + 9], isterminal=terminal)
        # sustaindc_env.py:721-737: the three methods are called in the order ls, dc, bat on the same params
        rewards = tuple(self._reward(name, info) for name in self.reward_methods)
        return obs, rewards, terminal, info

    def _reward(self, name, p):
        """utils/reward_creator.py:48-318 (the methods usable as shipped)."""
        if name in ("default_ls_reward", "default_dc_reward", "default_bat_reward"):
            energy = p["bat_total_energy_with_battery_KWh"]
            if name == "default_ls_reward":
                self.history.append(energy)                                 # :62-63 -- the only place the window grows
            foot = -1.0 * (p["norm_CI"] * normalize_energy(self.history, energy) / 0.50)
            if name != "default_ls_reward":
                return foot
            return np.clip(foot + (-0.3 * np.sqrt(p["ls_overdue_penalty"]) + 0.3) + (-0.1 * p["ls_oldest_task_age"]), -10, 10)
        if name == "custom_agent_reward":
            return 0.0                                                      # :133-146
        if name == "tou_reward":                                            # :154-202 (KeyError off the full hour)
            tou = {h: v for hs, v in (((0, 1, 2, 3, 4, 5, 22, 23), 0.25), ((6, 7, 8, 9, 10), 0.41), ((11, 12, 13, 14, 15), 0.30),
                                      ((16, 17, 18, 19, 20, 21), 0.27)) for h in hs}
            return -1.0 * p["bat_total_energy_with_battery_KWh"] * tou[p["hour"]]
        if name == "energy_efficiency_reward":
            return p["dc_ITE_total_power_kW"] / p["dc_total_power_kW"]      # :227-243
        if name == "energy_PUE_reward":                                     # :246-268
            pue = p["dc_total_power_kW"] / p["dc_ITE_total_power_kW"] if p["dc_ITE_total_power_kW"] != 0 else float("inf")
            return -abs(pue - 1)
        if name == "water_usage_efficiency_reward":
            return -0.01 * p["dc_water_usage"]                              # :297-318
        raise AssertionError("%s needs keys the env never provides (reward_creator.py:217,283)" % name)


def normalize_energy(history, value):
    """utils/reward_creator.py:16-45."""
    if len(history) < 2:
        return 0.0
    h = np.array(history)
    q1 = np.percentile(h, 25)
    q3 = np.percentile(h, 75)
    iqr = q3 - q1
    c = np.clip(h, q1 - 1.5 * iqr, q3 + 1.5 * iqr)
    mean, std = np.mean(c), np.std(c)
    return (value - mean) / (std if std > 0 else 1)


# ----------------------------------------------------------------------------------------------
# HARL view (harl/envs/sustaindc/harlsustaindc_env.py:42-131, supersuit zero padding)
# ----------------------------------------------------------------------------------------------
def harl_view(obs):
    """obs dict -> (obs[3,26] zero padded, share_obs[3,29]) for nonoverlapping_shared_obs_space=True."""
    rows = np.zeros((3, 26), np.float32)
    for i, k in enumerate(("agent_ls", "agent_dc", "agent_bat")):
        rows[i, :len(obs[k])] = obs[k]
    share = np.array(list(rows[0]) + [rows[1][11], rows[1][13]] + [rows[2][-1]], dtype=np.float32)
    return rows, np.stack([share] * 3)
