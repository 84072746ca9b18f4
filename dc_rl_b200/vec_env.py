"""`CudaShareVecEnv`: N SustainDC envs on one GPU behind harl's ``ShareVecEnv`` surface.

Replaces ``ShareSubprocVecEnv([HARLSustainDCEnv] * N)`` (reference harl/envs/env_wrappers.py:222-297 over
harl/envs/sustaindc/harlsustaindc_env.py:10-210): same attributes, same ``reset()`` / ``step(actions)`` return
shapes and dtypes, same auto-reset contract (when an env finishes, the returned obs / share_obs are the post-reset
ones and ``infos[i][0]`` carries ``original_obs`` / ``original_state`` / ``original_avail_actions``,
env_wrappers.py:173-192), so ``harl.runners`` drive it unchanged.  One process, one CUDA library handle, no pipes.
"""
import json
import os

import numpy as np

from . import info_layout
from ._lib import INFO_STRIDE, N_AGENTS, OBS_DIM, SHARE_DIM
from .dc_config import load_dc_config, size_datacenter
from .engine import Engine
from .traces import LocationTraces, location_key

AGENTS = ("agent_ls", "agent_dc", "agent_bat")
OBS_WIDTH = {"agent_ls": 26, "agent_dc": 14, "agent_bat": 13}


class Box:
    """Minimal stand-in for gymnasium.spaces.Box (the runners only read the class name, shape, low, high)."""

    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        shape = tuple(shape) if shape is not None else np.shape(low)
        self.shape = shape
        self.low = np.broadcast_to(np.asarray(low, self.dtype), shape).copy()
        self.high = np.broadcast_to(np.asarray(high, self.dtype), shape).copy()

    def sample(self):
        return np.random.uniform(self.low, self.high).astype(self.dtype)

    def __repr__(self):
        return "Box(%s, %s, %s, %s)" % (self.low.min(), self.high.max(), self.shape, self.dtype)


class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self.shape = ()
        self.dtype = np.dtype(np.int64)

    def sample(self):
        return int(np.random.randint(self.n))

    def __repr__(self):
        return "Discrete(%d)" % self.n


def make_spaces(nonoverlapping_shared_obs_space=True):
    """Spaces as the HARL adapter exposes them after supersuit padding (harlsustaindc_env.py:25-34,
    sustaindc_ptzoo.py:24-44; sub-env boxes: carbon_ls.py:40-45, make_envs_pyenv.py:114-132, bat_env_fwd_view.py:23-27)."""
    obs = [Box(-2.0, 2.0, (OBS_DIM,)), Box(-5.0e9, 5.0e9, (OBS_DIM,)), Box(-2.0, 2.0, (OBS_DIM,))]
    if nonoverlapping_shared_obs_space:
        share = [Box(-2.0, 2.0, (SHARE_DIM,)) for _ in range(N_AGENTS)]
    else:
        share = [Box(0.0, 1.0, (OBS_DIM * N_AGENTS,)) for _ in range(N_AGENTS)]
    act = [Discrete(3) for _ in range(N_AGENTS)]
    return obs, share, act


def resolve_traces(location, workload_file="Alibaba_CPU_Data_Hourly_1.csv", data_root=None, traces=None, timezone_shift=0):
    """Traces for a location: an explicit LocationTraces (or a {location key: LocationTraces} dict), the reference's
    data/ tree (``data_root`` or $SDC_DATA_ROOT), or -- only when asked for with traces='synthetic' -- the seeded
    synthetic year.  timezone_shift rolls the traces like the reference managers (managers.py:188,377,557-558)."""
    if isinstance(traces, dict):
        traces = traces[location_key(location)]
    if isinstance(traces, LocationTraces):
        if int(timezone_shift) != traces.timezone_shift:
            raise ValueError("explicit LocationTraces were built with timezone_shift=%d, env_args asks for %d" % (
                traces.timezone_shift, int(timezone_shift)))
        return traces
    if traces == "synthetic":
        return LocationTraces.synthetic(location_key(location), timezone_shift=timezone_shift)
    root = data_root or os.environ.get("SDC_DATA_ROOT")
    if root and os.path.isdir(root):
        return LocationTraces.from_reference_data(root, location, workload_file, timezone_shift)
    raise FileNotFoundError(
        "no trace data for location %r: pass env_args['data_root'] (the reference's data/ directory), set SDC_DATA_ROOT, "
        "or request env_args['traces']='synthetic'" % location)


class InfoRow:
    """Dict-like view of one env's info (the reference merges all sub-env infos into every agent's dict,
    sustaindc_env.py:676-710).  Supports [], get, in, keys, items, assignment of extra keys."""
    __slots__ = ("_b", "_i", "_extra")

    def __init__(self, batch, i, extra):
        self._b, self._i, self._extra = batch, i, extra

    def _lookup(self, key):
        if key in self._extra:
            return self._extra[key]
        b, i = self._b, self._i
        col = info_layout.COL.get(key)
        if col is not None:
            v = float(b.table[col, i])
            return bool(v) if key == info_layout.INFO_TERMINAL_KEY else v
        if key == info_layout.INFO_HIST_KEY:
            return b.table[info_layout.HIST0:info_layout.HIST0 + 5, i].astype(np.float64)
        if key == info_layout.INFO_FORECAST_KEY:
            return b.table[info_layout.FORECAST0:info_layout.FORECAST0 + 8, i].astype(np.float64)
        if key == "bat_a_t":
            return info_layout.BAT_ACTION_NAMES[int(b.table[info_layout.COL["bat_action"], i])]
        raise KeyError(key)

    def __getitem__(self, key):
        return self._lookup(key)

    def __setitem__(self, key, value):
        self._extra[key] = value

    def get(self, key, default=None):
        try:
            return self._lookup(key)
        except KeyError:
            return default

    def __contains__(self, key):
        return key in self._extra or key in info_layout.COL or key in _VECTOR_KEYS

    def keys(self):
        return list(info_layout.INFO_SCALAR_KEYS) + [info_layout.INFO_HIST_KEY] + list(info_layout.INFO_DC_KEYS) + list(
            info_layout.INFO_COMMON_KEYS) + [info_layout.INFO_FORECAST_KEY, info_layout.INFO_TERMINAL_KEY, "bat_a_t"] + list(self._extra)

    def items(self):
        return [(k, self._lookup(k)) for k in self.keys()]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def to_dict(self):
        return dict(self.items())


_VECTOR_KEYS = (info_layout.INFO_HIST_KEY, info_layout.INFO_FORECAST_KEY, "bat_a_t")


class _FinishedExtras:
    """The auto-reset extras of a step (`original_obs` / `original_state` / `original_avail_actions` of the envs that finished,
    env_wrappers.py:173-192), built per env on first access: a 65 536-env batch finishes ~100 envs per step, and eagerly
    building their dicts was a third of `CudaShareVecEnv.step`."""

    def __init__(self, finished, rows, term_share, avail, nonoverlapping):
        self._pos = {int(i): j for j, i in enumerate(finished)}
        self._rows, self._term_share, self._avail, self._nonoverlapping = rows, term_share, avail, nonoverlapping
        self._built = {}

    def _build(self, i):
        j = self._pos[i]
        o = self._rows[j]
        s = (np.repeat(self._term_share[j][None], N_AGENTS, 0) if self._nonoverlapping else np.repeat(o.reshape(1, -1), N_AGENTS, 0))
        return {"original_obs": o, "original_state": s, "original_avail_actions": self._avail[i].copy()}

    def setdefault(self, i, default):
        d = self._built.get(i)
        if d is None:
            d = self._build(i) if i in self._pos else default
            self._built[i] = d
        return d

    def __contains__(self, i):
        return i in self._pos or i in self._built

    def __getitem__(self, i):
        if i not in self:
            raise KeyError(i)
        return self.setdefault(i, {})

    def keys(self):
        return sorted(set(self._pos) | set(self._built))

    def __len__(self):
        return len(self.keys())


class InfoBatch:
    """``infos`` of one vec-env step: ``len(infos) == N``, ``infos[i][agent]`` is an InfoRow.  Backed by the step's
    [64, N] float table, which stays ON THE DEVICE until something is read: `column(key)` copies one column (what a logger
    needs: a dozen columns, not 16 MB per step), indexing a row copies the table once.  Valid until the env's next step."""

    def __init__(self, table, extras, n_envs=None, fetch=None):
        self._table = table                     # [INFO_STRIDE, N] float32 (owned), or None while it is still on the device
        self._fetch = fetch                     # fetch(first_col, n_cols) -> [n_cols, N]
        self._n = int(n_envs if n_envs is not None else table.shape[1])
        self._extras = extras                   # {env index: dict of extra keys for agent 0}
        self._rows = {}
        self._cols = {}

    @property
    def table(self):
        if self._table is None:
            self._table = self._fetch(0, info_layout.INFO_K_USED)
        return self._table

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[j] for j in range(*i.indices(len(self)))]
        i = int(i)
        if i < 0:
            i += len(self)
        row = self._rows.get(i)
        if row is None:
            extra0 = self._extras.setdefault(i, {})
            row = [InfoRow(self, i, extra0)] + [InfoRow(self, i, {}) for _ in range(N_AGENTS - 1)]
            self._rows[i] = row
        return row

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def column(self, key):
        """Vectorised access: one info key for all envs (fast path for loggers)."""
        col = info_layout.COL[key]
        if self._table is not None:
            return self._table[col]
        if col not in self._cols:
            self._cols[col] = self._fetch(col, 1)[0]
        return self._cols[col]


def _cyclic(value, index):
    """Per-env value of an env_args entry: scalars apply to every env, lists / tuples are cycled by `index`."""
    if isinstance(value, (list, tuple)):
        return [value[int(i) % len(value)] for i in index]
    return [value] * len(index)


class CudaShareVecEnv:
    """harl ``ShareVecEnv`` surface over one `Engine` (see module docstring).

    env_args are the reference's (sustaindc_env.py:38-80; harl/configs/envs_cfgs/sustaindc.yaml) plus additive keys:
    ``traces`` / ``data_root`` (where the year traces come from), ``device``.  ``location`` and ``dc_config_file`` may be
    LISTS: env with global id i then gets ``location[i % len(location)]`` and ``dc_config_file[(i // len(location)) %
    len(dc_config_file)]`` (BASELINE config 4: {az, ny, wa} x {dc1, dc2, dc3}); every distinct (dc_config, location) pair is
    sized once (utils/make_envs_pyenv.py:149-218 uses the location's design ambient) and selected per env on the device.
    Dead keys of the reference (weather_file, cintensity_file, flexible_load, individual_reward_weight, max_bat_cap_Mw,
    evaluation) are accepted and ignored, as there."""

    def __init__(self, env_args, n_envs, seed=0, months=None, seeds=None, device=0, lib=None, first_env_id=0, engine=None):
        """engine: wrap an existing `Engine` (its envs, traces and sizing) instead of building one from env_args."""
        args = dict(env_args)
        self.env_args = args
        self.num_envs = int(n_envs)
        self.n_agents = N_AGENTS
        self.nonoverlapping = bool(args.get("nonoverlapping_shared_obs_space", False))
        self.observation_space, self.share_observation_space, self.action_space = make_spaces(self.nonoverlapping)
        self.closed = False
        self._avail = np.ones((self.num_envs, N_AGENTS, 3), np.float32)
        self.lazy_info = args.get("info", "lazy") == "lazy"
        self.views = bool(args.get("output_views", False))
        if engine is not None:
            if engine.n_envs != self.num_envs:
                raise ValueError("engine holds %d envs, not %d" % (engine.n_envs, self.num_envs))
            self.engine = engine
            engine.set_tuning(lazy_info=int(self.lazy_info))
            return
        ids = np.arange(first_env_id, first_env_id + self.num_envs)
        location = args.get("location", "ny")
        n_loc_cycle = len(location) if isinstance(location, (list, tuple)) else 1
        env_loc = [location_key(x) for x in _cyclic(location, ids)]
        env_cfg = _cyclic(args.get("dc_config_file", "dc_config.json"), ids // n_loc_cycle)
        tz = int(args.get("timezone_shift", 0))
        loc_keys, traces = [], []
        cfg_keys, params, self.derived_all = [], [], []
        loc_id, cfg_id = np.zeros(self.num_envs, np.uint8), np.zeros(self.num_envs, np.uint8)
        flat_cache = {}
        for i, (lk, cf) in enumerate(zip(env_loc, env_cfg)):
            if lk not in loc_keys:
                loc_keys.append(lk)
                traces.append(resolve_traces(lk, args.get("workload_file", "Alibaba_CPU_Data_Hourly_1.csv"), args.get("data_root"),
                                             args.get("traces"), tz))
            ck = (json.dumps(cf, sort_keys=True) if isinstance(cf, dict) else str(cf), lk)
            if ck not in cfg_keys:
                if ck[0] not in flat_cache:
                    flat_cache[ck[0]] = cf if isinstance(cf, dict) else _find_dc_config(cf, args.get("data_root"))
                p, d = size_datacenter(lk, flat_cache[ck[0]], args.get("datacenter_capacity_mw", 1))
                cfg_keys.append(ck); params.append(p); self.derived_all.append(d)
            loc_id[i], cfg_id[i] = loc_keys.index(lk), cfg_keys.index(ck)
        self.derived = self.derived_all[0]
        self.env_location, self.env_dc_config = env_loc, [k[0] for k in (cfg_keys[c] for c in cfg_id)]
        if months is None:
            if "month" in args:                                   # harl/utils/envs_tools.py:56-62
                months = np.full(self.num_envs, int(args["month"]))
            else:
                months = np.where(ids < 12, ids % 12, ids % 3 + 5)
        if seeds is None:
            seeds = (int(seed) + ids * 1000).astype(np.uint64)   # envs_tools.py:67
        self.engine = Engine(self.num_envs, traces, params, loc_id=loc_id, cfg_id=cfg_id, months=months, seeds=seeds,
                             days_per_episode=int(args.get("days_per_episode", 7)), device=device, lib=lib)
        # additive keys: "info": "lazy" (default; the info table stays on the device until read) | "eager";
        # "output_views": True returns views of the pinned I/O buffers, valid until the next step, instead of fresh copies
        self.engine.set_tuning(lazy_info=int(self.lazy_info))
        rewards = [args.get(key, "default_%s" % key) for key in ("ls_reward", "dc_reward", "bat_reward")]
        if rewards != ["default_ls_reward", "default_dc_reward", "default_bat_reward"]:
            self.engine.set_reward_methods(*rewards)              # sustaindc_env.py:137-144

    # ---- ShareVecEnv API -----------------------------------------------------------------------
    def _share(self, obs, share):
        if self.nonoverlapping:
            return np.broadcast_to(share[:, None, :], (self.num_envs, N_AGENTS, SHARE_DIM)) if self.views else np.repeat(
                share[:, None, :], N_AGENTS, axis=1)
        flat = obs.reshape(self.num_envs, 1, N_AGENTS * OBS_DIM)
        return np.broadcast_to(flat, (self.num_envs, N_AGENTS, N_AGENTS * OBS_DIM)) if self.views else np.repeat(flat, N_AGENTS, axis=1)

    def reset(self):
        obs, share = self.engine.reset_host()
        if not self.views:
            obs = obs.copy()
        return obs, self._share(obs, share), (self._avail if self.views else self._avail.copy())

    def step_async(self, actions):
        a = np.asarray(actions).reshape(self.num_envs, N_AGENTS)
        self.engine.step_host_begin(a, want_info=not self.lazy_info, want_term=True)

    def step_wait(self):
        eng = self.engine
        obs, share, rew, done, info, term = eng.step_host_end()
        if not self.views:
            obs = obs.copy()
        share_obs = self._share(obs, share)
        extras = {}
        finished = np.nonzero(done)[0]
        if len(finished):
            rows = term[finished]                                     # copies only the finished envs' terminal rows
            term_share = np.concatenate([rows[:, 0, :], rows[:, 1, 11:12], rows[:, 1, 13:14], rows[:, 2, 25:26]], axis=1)
            extras = _FinishedExtras(finished, rows, term_share, self._avail, self.nonoverlapping)
        if self.lazy_info:
            step_id = eng.host_step_id

            def fetch(first, count):
                if eng.host_step_id != step_id:
                    raise RuntimeError("this InfoBatch belongs to an earlier step: its info table has been overwritten on the device")
                return eng.fetch_info(first, count)
            infos = InfoBatch(None, extras, self.num_envs, fetch)
        else:
            infos = InfoBatch(info.copy(), extras)
        if self.views:
            dones = np.broadcast_to(done.view(np.bool_)[:, None], (self.num_envs, N_AGENTS))
            return obs, share_obs, rew.reshape(self.num_envs, N_AGENTS, 1), dones, infos, self._avail
        dones = np.repeat(done.astype(bool)[:, None], N_AGENTS, axis=1)
        return obs, share_obs, rew.reshape(self.num_envs, N_AGENTS, 1).copy(), dones, infos, self._avail.copy()

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if not self.closed:
            self.engine.close()
            self.closed = True

    # ---- device-resident surface: torch CUDA tensors in, torch CUDA tensors out, nothing crosses PCIe -----
    def _torch_buffers(self):
        if not hasattr(self, "_tt"):
            import torch
            dev = torch.device("cuda", self.engine.device)
            n = self.num_envs
            z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)      # noqa: E731
            self._tt = dict(dev=dev, obs=z(n, N_AGENTS, OBS_DIM), share=z(n, SHARE_DIM), rew=z(n, N_AGENTS), done=z(n, dt=torch.uint8),
                            info=z(INFO_STRIDE, n), term=z(n, N_AGENTS, OBS_DIM))
        return self._tt

    def reset_torch(self):
        """reset() with device-resident results: (obs[N,3,26], share_obs[N,29]) torch CUDA tensors owned by the env."""
        import torch
        t = self._torch_buffers()
        self.engine.reset_device(t["obs"], t["share"], stream=torch.cuda.current_stream(t["dev"]).cuda_stream)
        return t["obs"], t["share"]

    def step_torch(self, actions, want_info=False):
        """step() for a device-resident rollout: `actions` is a CUDA tensor [N,3] (any integer / float dtype); returns
        (obs[N,3,26], share_obs[N,29], rewards[N,3], dones[N] uint8) CUDA tensors owned by the env and overwritten by the
        next call, enqueued on torch's current stream -- no host round trip, no synchronisation.  Auto-reset as in
        step(); the pre-reset observations of finished envs are in `terminal_obs_torch`; with want_info the [64,N] info
        table is in `info_torch` (row order: info_layout.INFO_COLUMNS)."""
        import torch
        t = self._torch_buffers()
        a = actions.reshape(self.num_envs, N_AGENTS).to(device=t["dev"], dtype=torch.int32).contiguous()
        self.engine.step_device(a, t["obs"], t["share"], t["rew"], t["done"], t["info"] if want_info else None, t["term"],
                                torch.cuda.current_stream(t["dev"]).cuda_stream)
        return t["obs"], t["share"], t["rew"], t["done"]

    @property
    def terminal_obs_torch(self):
        return self._torch_buffers()["term"]

    @property
    def info_torch(self):
        return self._torch_buffers()["info"]

    # ---- extras --------------------------------------------------------------------------------
    def metrics(self, clear=False):
        """Device-side running sums of the logger's per-step keys (sustaindc_logger.py:86-101) as a dict."""
        m = self.engine.metrics(clear)
        return dict(zip(METRIC_NAMES, m.tolist()))


    def hvac_power_stats(self, clear=False):
        """mean / max / 90th percentile of the positive dc_HVAC_total_power_kW samples since the last clear, from the
        device histogram (what SustainDCLogger.episode_log derives from its list of samples, sustaindc_logger.py:152-155).
        max is the upper edge of the highest occupied bin."""
        from .distributed import histogram_percentile
        counts, rng = self.engine.hvac_histogram(clear)
        total = float(counts.sum())
        if total == 0:
            return {"mean": 0.0, "max": 0.0, "p90": 0.0, "samples": 0}
        width = rng / len(counts)
        centres = (np.arange(len(counts)) + 0.5) * width
        top = int(np.nonzero(counts)[0][-1])
        return {"mean": float((centres * counts).sum() / total), "max": (top + 1) * width,
                "p90": histogram_percentile(counts, rng, 90.0), "samples": int(total)}


METRIC_NAMES = ("bat_total_energy_with_battery_KWh", "bat_CO2_footprint", "dc_water_usage", "ls_tasks_in_queue", "ls_tasks_dropped",
                "dc_ITE_total_power_kW", "dc_CT_total_power_kW", "dc_Compressor_total_power_kW", "dc_HVAC_total_power_kW",
                "env_steps", "reward_sum", "episodes", "reward_ls_sum", "reward_dc_sum", "ls_overdue_penalty", "dc_total_power_kW")


def _find_dc_config(name, data_root=None):
    """'dc_config.json' -> built-in default; other names are looked up next to the data root's utils/."""
    if name in (None, "dc_config.json"):
        return None
    if os.path.isfile(name):
        return load_dc_config(name)
    for base in filter(None, (data_root, os.environ.get("SDC_DATA_ROOT"))):
        for cand in (os.path.join(base, name), os.path.join(os.path.dirname(base.rstrip("/")), "utils", name)):
            if os.path.isfile(cand):
                return load_dc_config(cand)
    raise FileNotFoundError("dc_config_file %r not found" % name)
