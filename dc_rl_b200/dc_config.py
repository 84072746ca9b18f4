"""Data-centre description -> kernel constants (host side, runs once per (dc_config, location) pair).

Mirrors what the reference does at construction time:
  * ``DC_Config`` reading ``utils/dc_config.json``            (reference utils/dc_config_reader.py:39-145)
  * rack population under the per-rack power cap               (reference envs/datacenter.py:65-74)
  * CPU / IT-fan curve coefficients                            (reference envs/datacenter.py:31-49)
  * chiller + cooling-tower sizing and the 8x11 IT sweep       (reference utils/make_envs_pyenv.py:149-218,
                                                                 envs/datacenter.py:476-529)
The result is the plain-C ``sdc_dc_params`` struct of include/sdc_b200.h: racks with identical
parameters are merged into classes with a multiplicity (racks are internally homogeneous, so a class
costs two ``pow`` per step instead of two per rack).

Rack order: the reference shuffles racks through ``concurrent.futures.as_completed``
(utils/dc_config_reader.py:100-105); this loader always uses the JSON order.  Unlike the shipped reader
it also accepts ``dc_config_dc{1,2,3}.json`` (missing CHILLER_COP_BASE, list lengths != NUM_RACKS:
lists are cycled to NUM_RACKS entries).
"""
import json
import math

import numpy as np

from ._lib import DcParams, MAX_RACK_CLASSES

SP_MIN, SP_MAX, SP_INIT = 15.0, 21.6, 18.0      # utils/make_envs_pyenv.py:124-126
MAX_AMBIENT = {"ny": 30.0, "az": 50.0, "wa": 20.0}   # utils/make_envs_pyenv.py:149-157 (else 50)

DEFAULT_DC_CONFIG = {
    "data_center_configuration": {
        "NUM_ROWS": 4, "NUM_RACKS_PER_ROW": 5, "CPUS_PER_RACK": 200,
        "RACK_SUPPLY_APPROACH_TEMP_LIST": [5.3] * 5 + [5.0] * 10 + [5.3] * 5,
        "RACK_RETURN_APPROACH_TEMP_LIST": [-3.7] * 5 + [-2.5] * 10 + [-3.7] * 5,
    },
    "hvac_configuration": {
        "C_AIR": 1006, "RHO_AIR": 1.225, "CRAC_SUPPLY_AIR_FLOW_RATE_pu": 0.00005663,
        "CRAC_REFRENCE_AIR_FLOW_RATE_pu": 0.00009438, "CRAC_FAN_REF_P": 150,
        "CHILLER_COP_BASE": 5.0, "CHILLER_COP_K": 0.1, "CHILLER_COP_T_NOMINAL": 25.0,
        "CT_FAN_REF_P": 1000, "CT_REFRENCE_AIR_FLOW_RATE": 2.8315,
        "CW_PRESSURE_DROP": 300000, "CW_WATER_FLOW_RATE": 0.0011, "CW_PUMP_EFFICIENCY": 0.87,
        "CT_PRESSURE_DROP": 300000, "CT_WATER_FLOW_RATE": 0.0011, "CT_PUMP_EFFICIENCY": 0.87,
    },
    "server_characteristics": {
        "CPU_POWER_RATIO_LB": [0.01, 1.00], "CPU_POWER_RATIO_UB": [0.03, 1.02],
        "IT_FAN_AIRFLOW_RATIO_LB": [0.01, 0.225], "IT_FAN_AIRFLOW_RATIO_UB": [0.225, 1.0],
        "IT_FAN_FULL_LOAD_V": 0.051, "ITFAN_REF_V_RATIO": 1.0, "ITFAN_REF_P": 10.0, "INLET_TEMP_RANGE": [16, 28],
        "DEFAULT_SERVER_POWER_CHARACTERISTICS": (
            [[170, 20], [120, 10]] + [[130, 10]] * 11 + [[170, 10]] + [[130, 10]] * 2 + [[110, 10]] + [[170, 10]] * 3),
    },
}


def load_dc_config(path_or_dict=None):
    """Flattens a dc_config JSON (sectioned like the reference file, or already flat) into one dict."""
    if path_or_dict is None:
        raw = DEFAULT_DC_CONFIG
    elif isinstance(path_or_dict, dict):
        raw = path_or_dict
    else:
        with open(path_or_dict) as f:
            raw = json.load(f)
    flat = {}
    for key, val in raw.items():
        if isinstance(val, dict):
            flat.update(val)
        else:
            flat[key] = val
    for section in DEFAULT_DC_CONFIG.values():          # tolerate files that lack newer keys
        for key, val in section.items():
            flat.setdefault(key, val)
    return flat


def _cycled(values, n):
    values = list(values)
    return [values[i % len(values)] for i in range(n)]


def tolerant_flat_config(path_or_dict=None):
    """The flat dc_config this loader actually simulates: missing keys defaulted, per-rack lists cycled to
    NUM_ROWS * NUM_RACKS_PER_ROW entries.  The shipped reader rejects utils/dc_config_dc{1,2,3}.json (missing
    CHILLER_COP_BASE, list lengths != NUM_RACKS); parity tests feed the oracle this same dict."""
    c = dict(load_dc_config(path_or_dict))
    n = int(c["NUM_ROWS"]) * int(c["NUM_RACKS_PER_ROW"])
    for key in ("RACK_SUPPLY_APPROACH_TEMP_LIST", "RACK_RETURN_APPROACH_TEMP_LIST", "DEFAULT_SERVER_POWER_CHARACTERISTICS"):
        c[key] = _cycled(c[key], n)
    return c


class RackModel:
    """Per-rack closed form of the reference's per-CPU vectorised model (all CPUs of a rack are identical)."""

    def __init__(self, cfg, datacenter_capacity_mw=1.0):
        c = self.cfg = load_dc_config(cfg)
        self.n_racks = int(c["NUM_ROWS"]) * int(c["NUM_RACKS_PER_ROW"])
        cap_w = int(datacenter_capacity_mw * 1e6 / self.n_racks)                 # dc_config_reader.py:52-53
        t_lo, t_hi = c["INLET_TEMP_RANGE"]
        lb, ub = c["CPU_POWER_RATIO_LB"], c["CPU_POWER_RATIO_UB"]
        self.m_cpu = (ub[0] - lb[0]) / (t_hi - t_lo)
        self.c_cpu = ub[0] - self.m_cpu * t_hi
        self.shift_cpu = lb[1] - lb[0]
        flb, fub = c["IT_FAN_AIRFLOW_RATIO_LB"], c["IT_FAN_AIRFLOW_RATIO_UB"]
        self.m_fan = (fub[0] - flb[0]) / (t_hi - t_lo)
        self.c_fan = fub[0] - self.m_fan * t_hi
        self.shift_fan = flb[1] - flb[0]
        servers = np.asarray(_cycled(c["DEFAULT_SERVER_POWER_CHARACTERISTICS"], self.n_racks), dtype=np.float64)
        self.full, self.idle = servers[:, 0].copy(), servers[:, 1].copy()
        # CPUs are added while the running full-load sum stays below the cap; the one reaching it is dropped
        fits = np.ceil(cap_w / self.full) - 1
        self.ncpu = np.minimum(float(int(c["CPUS_PER_RACK"])), fits)
        self.supply = np.clip(np.asarray(_cycled(c["RACK_SUPPLY_APPROACH_TEMP_LIST"], self.n_racks), np.float64), 3.8, 5.3)
        self.ret = np.asarray(_cycled(c["RACK_RETURN_APPROACH_TEMP_LIST"], self.n_racks), np.float64)

    def it_power_and_outlet(self, load_pct, setpoint):
        """(cpu W, fan W, outlet degC) arrays over racks -- envs/datacenter.py:157-181,250-317."""
        c = self.cfg
        t_in = self.supply + setpoint
        ratio = ((self.m_cpu + 0.05) * t_in + self.c_cpu) + self.shift_cpu * (load_pct / 100)
        cpu_w = np.maximum(self.idle, self.full * ratio) * self.ncpu
        v = (self.m_fan * 10 * t_in + self.c_fan * 5) + self.shift_fan * (load_pct / 20)
        fan_w = (c["ITFAN_REF_P"] * (v / c["ITFAN_REF_V_RATIO"])) * self.ncpu
        flow = (c["IT_FAN_FULL_LOAD_V"] * v) * self.ncpu
        outlet = t_in + 1.918 * (cpu_w + fan_w) ** 1.096 / (c["C_AIR"] * c["RHO_AIR"] * flow ** 0.824 * 0.526) + (-14.01)
        return cpu_w, fan_w, outlet

    def crac_load(self, setpoint, return_temp, it_power):
        c = self.cfg
        return c["RHO_AIR"] * c["CRAC_SUPPLY_AIR_FLOW_RATE_pu"] * it_power * c["C_AIR"] * max(0.0, return_temp - setpoint)


def chiller_power(max_cooling_cap, load, ambient):
    """EnergyPlus-style electric chiller, reference envs/datacenter.py:356-429 (host copy used for sizing only)."""
    d_t = (ambient - 35.0) / 2.778 - (6.67 - 35.0)
    cap_ratio = 0.94483600 - 0.05700880 * d_t + 0.00185486 * d_t * d_t
    avail = max_cooling_cap * cap_ratio if cap_ratio != 0 else 0.0
    full_pow_ratio = 2.333 - 1.975 * cap_ratio + 0.6121 * cap_ratio * cap_ratio
    plr = min(max(load / avail, 0.05), 1.0) if avail > 0 else 0.0
    frac_full = 0.03303 + 0.6852 * plr + 0.2818 * plr * plr
    oper = (load / avail if load / avail < 0.05 else plr) if avail > 0 else 0.0
    cycling = min(1.0, oper / 0.05) if oper < 0.05 else 1.0
    return frac_full * full_pow_ratio * avail / 3.0 * cycling if oper > 0 else 0.0


def size_datacenter(location, cfg=None, datacenter_capacity_mw=1.0):
    """Returns (DcParams ctypes struct, dict of derived constants) for one (dc_config, location)."""
    rm = RackModel(cfg, datacenter_capacity_mw)
    c = rm.cfg
    loc = location.lower()
    max_amb = next((v for k, v in MAX_AMBIENT.items() if k in loc), 50.0)
    # chiller_sizing: 100 % load at the highest set-point, return air against the lowest set-point
    cpu_w, fan_w, outlet = rm.it_power_and_outlet(100.0, SP_MAX)
    load = rm.crac_load(SP_MIN, float(np.sum(rm.ret + outlet) / rm.n_racks), float(np.sum(cpu_w) + np.sum(fan_w)))
    ctafr = load / (c["C_AIR"] * max(50 - (max_amb - SP_MIN), 1)) / c["RHO_AIR"]
    ct_fan_ref_p = load
    ite, zone = [], []
    for sp in range(15, 23):
        for pct in range(0, 110, 10):
            cpu_w, fan_w, outlet = rm.it_power_and_outlet(float(pct), float(sp))
            ite.append(float(np.sum(cpu_w) + np.sum(fan_w)))
            zone.append(float(np.sum(outlet) / rm.n_racks))
    chiller_max = chiller_power(ct_fan_ref_p, max(ite), max_amb)
    max_dc_power_w = 1.1 * max(ite) + 1.1 * ct_fan_ref_p + 1.1 * chiller_max
    hvac_hi = 1.1 * ct_fan_ref_p + 1.1 * chiller_max
    derived = dict(ctafr=ctafr, ct_fan_ref_p=ct_fan_ref_p, power_lb_kw=(0.9 * min(ite) + 0.0) / 1e3,
                   power_ub_kw=(1.1 * max(ite) + hvac_hi) / 1e3, bat_capacity_mwh=(max_dc_power_w / 4) * (4 * 1) / 1e6,
                   zone_air=[0.9 * min(zone), 1.1 * max(zone)], max_ambient=max_amb, n_racks=rm.n_racks)
    # merge identical racks into classes (first-appearance order)
    classes = {}
    for r in range(rm.n_racks):
        key = (rm.full[r], rm.idle[r], rm.ncpu[r], rm.supply[r])
        classes[key] = classes.get(key, 0) + 1
    if len(classes) > MAX_RACK_CLASSES:
        raise ValueError("more than %d distinct rack classes" % MAX_RACK_CLASSES)
    p = DcParams()
    p.n_classes, p.n_racks = len(classes), rm.n_racks
    for i, ((full, idle, ncpu, supply), mult) in enumerate(classes.items()):
        p.cls_full[i], p.cls_idle[i], p.cls_ncpu[i], p.cls_supply[i], p.cls_mult[i] = full, idle, ncpu, supply, mult
    p.ret_mean = float(np.sum(rm.ret) / rm.n_racks)
    p.m_cpu, p.c_cpu, p.shift_cpu = rm.m_cpu, rm.c_cpu, rm.shift_cpu
    p.m_fan, p.c_fan, p.shift_fan = rm.m_fan, rm.c_fan, rm.shift_fan
    p.itfan_ref_p, p.itfan_ref_v_ratio, p.itfan_full_load_v = c["ITFAN_REF_P"], c["ITFAN_REF_V_RATIO"], c["IT_FAN_FULL_LOAD_V"]
    p.c_air, p.rho_air, p.crac_supply_flow_pu = c["C_AIR"], c["RHO_AIR"], c["CRAC_SUPPLY_AIR_FLOW_RATE_pu"]
    p.cw_pump_w = (c["CW_PRESSURE_DROP"] * c["CW_WATER_FLOW_RATE"]) / c["CW_PUMP_EFFICIENCY"]
    p.ct_pump_w = (c["CT_PRESSURE_DROP"] * c["CT_WATER_FLOW_RATE"]) / c["CT_PUMP_EFFICIENCY"]
    p.ctafr, p.ct_fan_ref_p = ctafr, ct_fan_ref_p
    p.power_lb_kw, p.power_ub_kw, p.bat_capacity_mwh = derived["power_lb_kw"], derived["power_ub_kw"], derived["bat_capacity_mwh"]
    return p, derived


def synthetic_dc_config(num_rows=5, racks_per_row=5, cpus_per_rack=40):
    """Builder-authored geometry for BASELINE config 3 (25 racks x 40 CPUs): the default server mix and
    approach temperatures cycled to the requested rack count."""
    cfg = json.loads(json.dumps(DEFAULT_DC_CONFIG))
    n = num_rows * racks_per_row
    d = cfg["data_center_configuration"]
    d["NUM_ROWS"], d["NUM_RACKS_PER_ROW"], d["CPUS_PER_RACK"] = num_rows, racks_per_row, cpus_per_rack
    d["RACK_SUPPLY_APPROACH_TEMP_LIST"] = _cycled([5.3, 5.0], n)
    d["RACK_RETURN_APPROACH_TEMP_LIST"] = _cycled([-3.7, -2.5], n)
    cfg["server_characteristics"]["DEFAULT_SERVER_POWER_CHARACTERISTICS"] = _cycled([[170, 20], [120, 10], [130, 10], [110, 10]], n)
    return cfg


MONTH_INIT_DAY = [0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334]     # utils/utils_cf.py:56-77


def start_day_range(month):
    """Admissible episode start days for a month: init_day +- 7 clipped to [0, 364] (sustaindc_env.py:197-198,454)."""
    init = MONTH_INIT_DAY[int(month)]
    return max(0, init - 7), min(364, init + 7)


def _selfcheck():
    assert math.isclose(size_datacenter("ny")[1]["ctafr"], 53.489453509150, rel_tol=1e-9)


if __name__ == "__main__":
    _selfcheck()
    print(size_datacenter("ny")[1])
