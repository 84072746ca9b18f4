"""Exogenous traces: hourly inputs -> the 15-min device tables of one location (host side, init time).

What the reference managers do when they are constructed / reset (reference utils/managers.py):
  * hourly -> 15 min by ``np.interp(np.linspace(0, n, 4n), range(n), x)``           :183-185 (same for CI, weather)
  * workload: 5/95-percentile rescale to [0.2, 0.8], clip [0,1], 16-tap moving average   :220-244, 208, 268-271
  * carbon intensity: NaN -> mean, clip >= 0; normalised per episode by the min/max of the next 30 days :359-361,417,435-437
  * weather: dry bulb + wet bulb (psychrolib from RH and pressure), per-episode noise/roll  :521-550, 594-613
All of these are deterministic per file except the weather noise, so they are built once here.  The
integer task counts of the load-shifting env (``ceil(w*0.8*100)``, ``floor(w*0.2*100)``, reference
envs/carbon_ls.py:194-195) are evaluated here in fp64 so that the device never re-derives an integer
from a rounded float.
"""
import os

import numpy as np

from . import psychro
from ._lib import TRACE_PAD, YEAR_STEPS

HOURS = 8760
NORM_WINDOW = 30 * 96          # managers.py:435, 606


def interp15(hourly):
    hourly = np.asarray(hourly, dtype=np.float64)
    n = len(hourly)
    return np.interp(np.linspace(0, n, n * 4), np.arange(n), hourly)


def _forward_extreme(x, window, fn):
    """fn over x[t : t+window] for every t (slice truncated at the array end, like the reference's)."""
    n = len(x)
    fill = np.inf if fn is np.min else -np.inf
    padded = np.concatenate([x, np.full(window - 1, fill)])
    view = np.lib.stride_tricks.sliding_window_view(padded, window)
    out = np.empty(n)
    step = 2048
    for s in range(0, n, step):
        out[s:s + step] = fn(view[s:s + step], axis=1)
    return out


def _pad(x, dtype):
    x = np.asarray(x, dtype=dtype)
    return np.ascontiguousarray(np.concatenate([x, np.full(TRACE_PAD, x[-1], dtype=dtype)]))


def hour_table():
    """(cos, sin) of sc_obs for the 96 quarter-hours, with Python's round() as in managers.py:66-88."""
    cos, sin = np.empty(96), np.empty(96)
    for k in range(96):
        ang = round((k * 0.25) / 24, 3) * (np.pi * 2)
        cos[k], sin[k] = np.cos(ang) * 0.5 + 0.5, np.sin(ang) * 0.5 + 0.5
    return cos, sin


class LocationTraces:
    """Device-ready tables of one location (all length YEAR_STEPS + TRACE_PAD)."""

    def __init__(self, cpu_load, avg_ci, dry_bulb, wet_bulb, name="custom", timezone_shift=0):
        """timezone_shift (hours): every 15-min trace is rolled by -4*shift samples right after the interpolation, as the
        reference managers do (managers.py:188,377,557-558)."""
        self.name = name
        self.timezone_shift = int(timezone_shift)
        sh = -4 * self.timezone_shift
        cpu = np.roll(interp15(np.asarray(cpu_load, np.float64)[:HOURS]), sh)
        p5, p95 = np.percentile(cpu, 5), np.percentile(cpu, 95)
        scaled = np.clip(0.2 + ((cpu - p5) * (0.8 - 0.2) / (p95 - p5)), 0, 1)
        workload = np.convolve(scaled, np.ones(16) / 16, mode="same")
        ci_h = np.asarray(avg_ci, np.float64)[:HOURS]
        if np.isnan(ci_h).any():
            ci_h = np.nan_to_num(ci_h, nan=np.nanmean(ci_h))
        ci = np.clip(np.roll(interp15(ci_h), sh), 0, None)
        temp = np.roll(interp15(np.asarray(dry_bulb, np.float64)[:HOURS]), sh)
        wetb = np.roll(interp15(np.asarray(wet_bulb, np.float64)[:HOURS]), sh)
        for arr in (workload, ci, temp, wetb):
            if len(arr) != YEAR_STEPS:
                raise ValueError("traces must cover 8760 hours")
        if workload.min() < 0 or workload.max() > 1:
            raise ValueError("The workload should be between 0 and 1")       # envs/carbon_ls.py:333-336
        self.workload = _pad(workload, np.float64)
        self.ns_tasks = _pad(np.ceil(workload * 0.8 * 100), np.uint8)
        self.sh_tasks = _pad(np.floor(workload * 0.2 * 100), np.uint8)
        self.ci = _pad(ci, np.float64)
        self.ci_min30 = _pad(_forward_extreme(ci, NORM_WINDOW, np.min), np.float64)
        self.ci_max30 = _pad(_forward_extreme(ci, NORM_WINDOW, np.max), np.float64)
        self.temp_base = _pad(temp, np.float64)
        self.wetb_base = _pad(wetb, np.float64)

    @classmethod
    def from_hourly_columns(cls, cpu_load, avg_ci, dry_bulb, rel_hum_pct, pressure_pa, name="custom", timezone_shift=0):
        """Wet bulb from (dry bulb, RH %, station pressure) like managers.py:521-530."""
        wet = [psychro.wet_bulb_from_rel_hum(t, rh / 100, p) for t, rh, p in zip(dry_bulb, rel_hum_pct, pressure_pa)]
        return cls(cpu_load, avg_ci, dry_bulb, wet, name, timezone_shift)

    @classmethod
    def from_npz(cls, path, name=None, timezone_shift=0):
        """Hourly columns saved as npz (keys cpu_load, avg_ci, dry_bulb, rel_hum, pressure)."""
        z = np.load(path, allow_pickle=False)
        return cls.from_hourly_columns(z["cpu_load"], z["avg_ci"], z["dry_bulb"], z["rel_hum"], z["pressure"],
                                       name or os.path.basename(path), timezone_shift)

    @classmethod
    def from_reference_data(cls, data_root, location, workload_file="Alibaba_CPU_Data_Hourly_1.csv", timezone_shift=0):
        """Reads the reference's data/ tree: Workload/*.csv (cpu_load), CarbonIntensity/<loc>_NG_&_avgCI.csv
        (avg_CI), Weather/*.epw (columns 6, 8, 9 after 8 header rows) -- managers.py:168-174,345-351,521-528;
        file choice per location as utils/utils_cf.py:11-40."""
        ci_loc, epw = LOCATION_FILES[location_key(location)]
        wl_path = os.path.join(data_root, "Workload", workload_file)
        ci_path = os.path.join(data_root, "CarbonIntensity", "%s_NG_&_avgCI.csv" % ci_loc)
        epw_path = os.path.join(data_root, "Weather", epw)
        try:
            # The reference parses with pandas.read_csv, whose default float parser is NOT correctly rounded (it can differ
            # from float() in the last bit): use the same parser when pandas is there, so the tables are bit-identical.
            import pandas as pd
            cpu = pd.read_csv(wl_path)["cpu_load"].values[:HOURS].astype(np.float64)
            ci = pd.read_csv(ci_path)["avg_CI"].values[:HOURS].astype(np.float64)
            wea = pd.read_csv(epw_path, skiprows=8, header=None).values[:, [6, 8, 9]].astype(np.float64)
        except ImportError:
            def column(path, name):
                with open(path) as f:
                    header = f.readline().strip().split(",")
                    idx = header.index(name)
                    return np.array([float(line.split(",")[idx] or "nan") for line in f if line.strip()])

            cpu, ci = column(wl_path, "cpu_load"), column(ci_path, "avg_CI")
            rows = []
            with open(epw_path) as f:
                for i, line in enumerate(f):
                    if i >= 8 and line.strip():
                        parts = line.split(",")
                        rows.append((float(parts[6]), float(parts[8]), float(parts[9])))
            wea = np.array(rows)
        return cls.from_hourly_columns(cpu, ci, wea[:, 0], wea[:, 1], wea[:, 2], location, timezone_shift)

    @classmethod
    def synthetic(cls, location="ny", seed=1234, timezone_shift=0):
        """Seeded synthetic 1-year traces with the moments of the shipped files (SURVEY.md section 8d,
        config 3): used by bench.py, where the reference's data files are not available."""
        rng = np.random.default_rng(seed + {"ny": 0, "az": 1, "wa": 2}.get(location.lower(), 3))
        h = np.arange(HOURS)
        day, year = 2 * np.pi * (h % 24) / 24, 2 * np.pi * h / HOURS
        mean_t = {"ny": 13.3, "az": 23.9, "wa": 11.0}.get(location.lower(), 15.0)
        cpu = 0.45 + 0.18 * np.sin(day - 2.0) + 0.08 * np.sin(2 * day + 0.5) + 0.05 * rng.standard_normal(HOURS)
        cpu += 0.06 * np.sin(2 * np.pi * h / (24 * 7))
        cpu = np.clip(cpu, 0.02, 0.98)
        slow = np.cumsum(rng.standard_normal(HOURS)) * 1.5
        slow -= np.linspace(slow[0], slow[-1], HOURS)
        ci = 272.8 + 22 * np.sin(day - 1.0) + 12 * np.sin(year * 2) + np.clip(slow, -60, 60) * 0.4 + 6 * rng.standard_normal(HOURS)
        ci = np.clip(ci, 120, 420)
        # synoptic fronts: AR(1) with a ~3-day correlation time, sigma ~3.5 C (never pins a month below freezing)
        shocks = rng.standard_normal(HOURS) * 3.5 * np.sqrt(1 - 0.986 ** 2)
        front = np.empty(HOURS)
        acc = 0.0
        for i in range(HOURS):
            acc = 0.986 * acc + shocks[i]
            front[i] = acc
        dry = mean_t - 11.5 * np.cos(year) + 4.0 * np.sin(day - 2.4) + np.clip(front, -9, 9)
        wet = dry - rng.uniform(1.0, 5.0, HOURS)
        return cls(cpu, ci, dry, wet, "synthetic-" + location, timezone_shift)


# location -> (carbon-intensity region, EPW file), reference utils/utils_cf.py:11-40
LOCATION_FILES = {
    "az": ("AZ", "USA_AZ_Phoenix-Sky.Harbor.epw"), "ca": ("CA", "USA_CA_San.Jose-Mineta.epw"),
    "ga": ("GA", "USA_GA_Atlanta-Hartsfield-Jackson.epw"), "il": ("IL", "USA_IL_Chicago.OHare.epw"),
    "ny": ("NY", "USA_NY_New.York-LaGuardia.epw"), "tx": ("TX", "USA_TX_Dallas-Fort.Worth.epw"),
    "va": ("VA", "USA_VA_Leesburg.Exec.epw"), "wa": ("WA", "USA_WA_Seattle-Tacoma.epw"),
}


def location_key(location):
    """The reference matches by substring in this order (utils/utils_cf.py:22-40)."""
    loc = location.lower()
    for key in ("az", "ca", "ga", "il", "ny", "tx", "va", "wa"):
        if key in loc:
            return key
    raise ValueError("Location not found, please define the location %s" % location)


def generate_weather_window(traces, t0, win_len, rng):
    """One episode's realised weather exactly as Weather_Manager.reset builds it (managers.py:35-48,594-613),
    drawing from a legacy ``np.random``-style generator: normal(35040) then randint(0, 14).  Host helper for
    replay mode (sdc_stage_episode); the on-device generator (k_reset) uses Philox instead.
    Returns (temp_window, wetb_window, t_min30, t_max30)."""
    n = YEAR_STEPS
    steps = rng.normal(loc=0, scale=1, size=n)
    walk = np.cumsum(0.02 * steps)
    noise = 0 + (walk / np.std(walk)) * 0.75
    roll = rng.randint(0, 14)
    temp = np.clip(np.roll(traces.temp_base[:n] + noise, roll * 96), 0, 45)
    wetb = np.clip(np.roll(traces.wetb_base[:n] + noise, roll * 96), 0, 45)
    seg = temp[t0:t0 + NORM_WINDOW]
    out_t, out_w = np.zeros(win_len), np.zeros(win_len)
    k = min(win_len, n - t0)
    out_t[:k], out_w[:k] = temp[t0:t0 + k], wetb[t0:t0 + k]
    return out_t, out_w, float(np.min(seg)), float(np.max(seg))
