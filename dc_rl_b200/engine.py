"""Thin object wrapper over the C ABI handle (include/sdc_b200.h): one `Engine` == N batched SustainDC envs
on one GPU.  Host code stays Python; all per-step arithmetic happens in the CUDA library.

Two call styles, same kernels:
  * device buffers (torch CUDA tensors, anything with ``data_ptr()``): ``reset_device`` / ``step_device``
  * host buffers (numpy): ``reset_host`` / ``step_host`` -- H2D, step, D2H through the library's pinned staging,
    i.e. what a numpy-based runner such as harl's sees.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import INFO_STRIDE, N_AGENTS, N_METRICS, OBS_COMPACT, OBS_DIM, SHARE_DIM
from .dc_config import start_day_range
from .traces import hour_table

_STATE_DTYPES = {
    "episode": np.uint32, "t": np.int32, "t0": np.int32, "step_in_ep": np.int32, "ci_min": np.float64, "ci_max": np.float64,
    "t_min": np.float64, "t_max": np.float64, "weather": np.float64, "ls_head": np.int32, "ls_len": np.int32,
    "ls_sum": np.int32, "ls_bins": np.uint16, "ls_ring": np.uint8, "setpoint": np.float64, "dc_run": np.int32,
    "dc_scale": np.int32, "dc_last": np.int8, "bat_load": np.float64, "hist": np.float32, "hist_ref": np.float64, "hist_len": np.int32,
    "hist_head": np.int32, "phase_clocks": np.uint64, "unit_log": np.uint32, "qlist": np.float32, "q_a": np.int32, "q_m": np.int32, "err": np.int32,
    "mom_s1": np.float64, "mom_s2": np.float64, "mom_c0": np.float64, "tails": np.float32, "tail_n": np.int32, "tail_nb": np.int32, "tail_bs": np.float64,
    "tail_thr": np.float32, "agg_n": np.int32, "agg_s": np.float64, "fast_cfg": np.uint32, "counters": np.int32, "pass_stats": np.int32, "pass_total": np.uint64,
    "pend_valid": np.uint8, "pend_weather": np.float64, "cur_buf": np.uint8, "pend_obs": np.float32, "metrics": np.float64, "hvac_hist": np.uint64,
}

_default_lib = None


def default_lib():
    global _default_lib
    if _default_lib is None:
        _default_lib = _lib.load()
    return _default_lib


def _ptr(x):
    if x is None:
        return None
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if isinstance(x, np.ndarray):
        if not x.flags["C_CONTIGUOUS"]:
            raise ValueError("buffer must be C-contiguous")
        return C.c_void_p(x.ctypes.data)
    if isinstance(x, int):
        return C.c_void_p(x)
    raise TypeError("expected a tensor, ndarray or raw address")


class SdcError(RuntimeError):
    pass


class Engine:
    def __init__(self, n_envs, locations, dc_params, loc_id=None, cfg_id=None, months=None, seeds=None,
                 days_per_episode=7, device=0, hist_cap=_lib.HIST_CAP, unit_envs=0, lib=None):
        """locations: list of traces.LocationTraces; dc_params: list of _lib.DcParams (one per (dc_config,
        location) pair); loc_id / cfg_id / months / seeds: per-env arrays (scalars broadcast)."""
        self.lib = lib or default_lib()
        self.device = int(device)
        self.n_envs = int(n_envs)
        self.ep_len = int(days_per_episode) * 96
        self.locations, self.dc_params = list(locations), list(dc_params)
        cfg = _lib.Config(self.n_envs, int(device), self.ep_len, len(self.locations), len(self.dc_params), int(hist_cap),
                          int(unit_envs), 0)
        handle = C.c_void_p()
        rc = self.lib.sdc_create(C.byref(cfg), C.byref(handle))
        if rc:
            raise SdcError("sdc_create: %s" % self.lib.sdc_last_error(None).decode())
        self._h = handle
        self.hist_cap = int(hist_cap)
        for i, loc in enumerate(self.locations):
            s = _lib.Location(*[loc_arr.ctypes.data for loc_arr in (
                loc.workload, loc.ns_tasks, loc.sh_tasks, loc.ci, loc.ci_min30, loc.ci_max30, loc.temp_base, loc.wetb_base)])
            self._check(self.lib.sdc_set_location(self._h, i, C.byref(s)))
        for i, p in enumerate(self.dc_params):
            self._check(self.lib.sdc_set_dc_params(self._h, i, C.byref(p)))
        cos, sin = hour_table()
        self._check(self.lib.sdc_set_hour_table(self._h, _ptr(cos), _ptr(sin)))
        n = self.n_envs
        self.loc_id = np.ascontiguousarray(np.broadcast_to(np.asarray(0 if loc_id is None else loc_id, np.uint8), (n,)))
        self.cfg_id = np.ascontiguousarray(np.broadcast_to(np.asarray(0 if cfg_id is None else cfg_id, np.uint8), (n,)))
        months = np.broadcast_to(np.asarray(0 if months is None else months, np.int64), (n,))
        ranges = np.array([start_day_range(m) for m in range(12)], np.int16)
        self.day_lo = np.ascontiguousarray(ranges[months, 0])
        self.day_hi = np.ascontiguousarray(ranges[months, 1])
        seeds = np.arange(n, dtype=np.uint64) if seeds is None else np.broadcast_to(np.asarray(seeds, np.uint64), (n,))
        self.seeds = np.ascontiguousarray(seeds)
        self._check(self.lib.sdc_assign(self._h, _ptr(self.loc_id), _ptr(self.cfg_id), _ptr(self.day_lo), _ptr(self.day_hi),
                                        _ptr(self.seeds)))
        self.win_len = self.lib.sdc_window_len(self._h)

    # ---- plumbing ------------------------------------------------------------------------------
    def _check(self, rc):
        if rc < 0:
            raise SdcError(self.lib.sdc_last_error(self._h).decode())
        return rc

    def close(self):
        if getattr(self, "_h", None):
            self.__dict__.pop("_hb", None)        # views into the handle's pinned buffers die with it
            self.lib.sdc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_tuning(self, **kw):
        for k, v in kw.items():
            self._check(self.lib.sdc_set_tuning(self._h, k.encode(), int(v)))

    def set_reward_methods(self, ls="default_ls_reward", dc="default_dc_reward", bat="default_bat_reward"):
        """Reward method per agent by the reference's names (utils/reward_creator.py:322-334)."""
        ids = []
        for name in (ls, dc, bat):
            if name in ("renewable_energy_reward", "temperature_efficiency_reward"):
                raise AssertionError("%s needs keys the reference env never provides (utils/reward_creator.py:217,283)" % name)
            if name not in _lib.REWARD_METHOD_IDS:
                raise AssertionError("Specified Reward Method %s not in REWARD_METHOD_MAP" % name)       # reward_creator.py:346
            ids.append(_lib.REWARD_METHOD_IDS[name])
        self._check(self.lib.sdc_set_reward_methods(self._h, *ids))
        self.reward_methods = (ls, dc, bat)

    # ---- episodes ------------------------------------------------------------------------------
    def stage_episode(self, env_ids, day, hour, temp_win, wetb_win, t_min30, t_max30):
        """Replay mode: the next reset of `env_ids` uses these starts / realised weather windows
        (each window ep_len + 18 long) instead of the device RNG."""
        env_ids = np.ascontiguousarray(env_ids, np.int32)
        k = len(env_ids)
        w = self.ep_len + 18
        temp = np.ascontiguousarray(np.asarray(temp_win, np.float64).reshape(k, -1)[:, :w])
        wetb = np.ascontiguousarray(np.asarray(wetb_win, np.float64).reshape(k, -1)[:, :w])
        if temp.shape[1] != w or wetb.shape[1] != w:
            raise ValueError("weather windows must hold at least ep_len + 18 = %d samples" % w)
        day = np.ascontiguousarray(day, np.int32)
        hour = np.ascontiguousarray(hour, np.int32)
        t_min30 = np.ascontiguousarray(t_min30, np.float64)
        t_max30 = np.ascontiguousarray(t_max30, np.float64)
        if not (len(day) == len(hour) == len(t_min30) == len(t_max30) == k):
            raise ValueError("stage_episode: per-env arrays must have len(env_ids) entries")
        self._check(self.lib.sdc_stage_episode(self._h, k, _ptr(env_ids), _ptr(day), _ptr(hour), _ptr(temp), _ptr(wetb),
                                               _ptr(t_min30), _ptr(t_max30)))

    def reset_device(self, obs, share, mask=None, stream=None):
        self._check(self.lib.sdc_reset(self._h, _ptr(mask), _ptr(obs), _ptr(share), _ptr(stream)))

    def step_device(self, actions, obs, share, rew, done, info=None, term_obs=None, stream=None):
        self._check(self.lib.sdc_step(self._h, _ptr(actions), _ptr(obs), _ptr(share), _ptr(rew), _ptr(done), _ptr(info),
                                      _ptr(term_obs), _ptr(stream)))

    def _host_buffers(self):
        """numpy views of the handle's page-locked I/O buffers (sdc_host_buffers): the host-side calls then move
        data straight between these and the device, with no staging copy."""
        if not hasattr(self, "_hb"):
            n = self.n_envs
            p = [C.c_void_p() for _ in range(7)]
            self._check(self.lib.sdc_host_buffers(self._h, *[C.byref(x) for x in p]))
            q = [C.c_void_p() for _ in range(2)]
            self._check(self.lib.sdc_host_buffers_compact(self._h, *[C.byref(x) for x in q]))

            def view(ptr, shape, ctype, dtype):
                count = int(np.prod(shape))
                return np.frombuffer((ctype * count).from_address(ptr.value), dtype=dtype).reshape(shape)
            self._hb = dict(actions=view(p[0], (n, N_AGENTS), C.c_int32, np.int32),
                            obs=view(p[1], (n, N_AGENTS, OBS_DIM), C.c_float, np.float32),
                            share=view(p[2], (n, SHARE_DIM), C.c_float, np.float32),
                            rew=view(p[3], (n, N_AGENTS), C.c_float, np.float32), done=view(p[4], (n,), C.c_uint8, np.uint8),
                            info=view(p[5], (INFO_STRIDE, n), C.c_float, np.float32),
                            term=view(p[6], (n, N_AGENTS, OBS_DIM), C.c_float, np.float32),
                            obs_c=view(q[0], (n, OBS_COMPACT), C.c_float, np.float32),
                            term_c=view(q[1], (n, OBS_COMPACT), C.c_float, np.float32))
        return self._hb

    def reset_host(self, mask=None):
        b = self._host_buffers()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        self._check(self.lib.sdc_reset_host(self._h, _ptr(m), _ptr(b["obs"]), _ptr(b["share"])))
        return b["obs"], b["share"]

    def step_host_begin(self, actions, want_info=True, want_term=True):
        """First half of step_host (ShareVecEnv.step_async): copies and the step are enqueued, the call returns."""
        b = self._host_buffers()
        a = b["actions"]
        a[...] = np.asarray(actions).reshape(self.n_envs, N_AGENTS)       # the one host copy: caller's actions -> pinned
        self._pending = (want_info, want_term)
        self._check(self.lib.sdc_step_host_begin(self._h, _ptr(a), _ptr(b["obs"]), _ptr(b["share"]), _ptr(b["rew"]), _ptr(b["done"]),
                                                 _ptr(b["info"]) if want_info else None, _ptr(b["term"]) if want_term else None))

    def step_host_end(self):
        b = self._host_buffers()
        want_info, want_term = self._pending
        self._check(self.lib.sdc_step_host_end(self._h))
        self.host_step_id = getattr(self, "host_step_id", 0) + 1
        return b["obs"], b["share"], b["rew"], b["done"], (b["info"] if want_info else None), (b["term"] if want_term else None)

    def step_host(self, actions, want_info=True, want_term=True):
        """actions int32 [N,3] -> (obs[N,3,26], share[N,29], rew[N,3], done[N], info[64,N] | None, term_obs | None).
        The returned arrays are views of the handle's pinned buffers, reused by the next call.  term_obs rows are valid
        for the envs that finished in this step."""
        self.step_host_begin(actions, want_info, want_term)
        return self.step_host_end()

    def step_compact_host(self, actions, want_info=False, want_term=True):
        """Compact host call: (obsc[N,29], rew[N,3], done[N], info | None, termc[N,29] | None) -- the 29 distinct observation
        values of every env (agent_ls[26] | workload(t+1) | norm T(t+1) | SoC; include/sdc_b200.h); `expand_obs` rebuilds the
        padded rows and the shared observation bit for bit."""
        b = self._host_buffers()
        a = b["actions"]
        a[...] = np.asarray(actions).reshape(self.n_envs, N_AGENTS)
        self._check(self.lib.sdc_step_compact_host(self._h, _ptr(a), _ptr(b["obs_c"]), _ptr(b["rew"]), _ptr(b["done"]),
                                                   _ptr(b["info"]) if want_info else None, _ptr(b["term_c"]) if want_term else None))
        self.host_step_id = getattr(self, "host_step_id", 0) + 1
        return b["obs_c"], b["rew"], b["done"], (b["info"] if want_info else None), (b["term_c"] if want_term else None)

    def step_compact_device(self, actions, obsc, rew, done, info=None, termc=None, stream=None):
        self._check(self.lib.sdc_step_compact(self._h, _ptr(actions), _ptr(obsc), _ptr(rew), _ptr(done), _ptr(info), _ptr(termc),
                                              _ptr(stream)))

    def expand_obs(self, obsc, want_share=True):
        """Host utility: compact rows -> (obs[n,3,26] zero padded, share[n,29] | None)."""
        c = np.ascontiguousarray(obsc, np.float32).reshape(-1, OBS_COMPACT)
        obs = np.empty((len(c), N_AGENTS, OBS_DIM), np.float32)
        share = np.empty((len(c), SHARE_DIM), np.float32) if want_share else None
        self.lib.sdc_expand_obs(_ptr(c), len(c), _ptr(obs), _ptr(share))
        return obs, share

    def fetch_info(self, first_col=0, n_cols=INFO_STRIDE):
        """Columns of the LAST host step's info table, [n_cols, N] (needs set_tuning(lazy_info=1) or want_info)."""
        out = np.empty((n_cols, self.n_envs), np.float32)
        self._check(self.lib.sdc_fetch_info(self._h, int(first_col), int(n_cols), _ptr(out)))
        return out

    # ---- metrics / state -----------------------------------------------------------------------
    def metrics(self, clear=False):
        out = np.zeros(N_METRICS, np.float64)
        self._check(self.lib.sdc_metrics(self._h, _ptr(out), int(bool(clear))))
        return out

    def hvac_histogram(self, clear=False):
        """(counts[HVAC_BINS] uint64, range_kw): histogram of the positive dc_HVAC_total_power_kW samples since the last
        clear, equal bins over [0, range_kw) -- what the logger's mean / max / p90 are computed from."""
        counts = np.zeros(_lib.HVAC_BINS, np.uint64)
        rng = C.c_double(0.0)
        self._check(self.lib.sdc_hvac_histogram(self._h, _ptr(counts), C.byref(rng), int(bool(clear))))
        return counts, float(rng.value)

    def prefill_history(self, values):
        """values: fp32 [count] (shared by all envs) or [N, count]."""
        if getattr(self, "reward_methods", ("default_ls_reward",))[0] != "default_ls_reward":
            raise NotImplementedError("a pre-filled reward window with a non-default ls_reward (the window never changes) is not supported")
        v = np.ascontiguousarray(values, np.float32)
        per_env = int(v.ndim == 2)
        count = v.shape[-1]
        self._check(self.lib.sdc_prefill_history(self._h, _ptr(v), count, per_env))

    def rebuild_brackets(self, stream=None):
        self._check(self.lib.sdc_rebuild_brackets(self._h, _ptr(stream)))

    def read_state(self, name):
        dt = np.dtype(_STATE_DTYPES[name])
        n = self.n_envs
        per_env = {"weather": 2 * self.win_len, "pend_weather": 2 * self.win_len, "pend_obs": N_AGENTS * OBS_DIM, "ls_bins": 4, "hist": self.hist_cap, "qlist": 2 * _lib.LIST_CAP, "q_a": 2, "q_m": 2,
                   "tail_n": 2, "tail_nb": 2, "tail_bs": 4, "tail_thr": 4, "agg_n": 2, "agg_s": 4, "tails": 2 * _lib.TAIL_CAP}.get(name, 1)
        if name == "phase_clocks":
            out = np.zeros(16, dt)
        elif name == "unit_log":
            out = np.zeros(((n + 7) // 8) * 8, dt)
        elif name == "counters":
            out = np.zeros(64, dt)
        elif name in ("pass_stats", "pass_total"):
            out = np.zeros(4, dt)
        elif name in ("metrics", "hvac_hist"):
            out = np.zeros(_lib.HVAC_BINS if name == "hvac_hist" else N_METRICS, dt)
        elif name == "ls_ring":
            out = np.zeros(n * 65536, dt)       # upper bound; trimmed below
        else:
            out = np.zeros(n * per_env, dt)
        got = self._check(self.lib.sdc_read_state(self._h, name.encode(), _ptr(out), out.nbytes))
        out = out[:got // dt.itemsize]
        if name in ("phase_clocks", "counters", "pass_stats", "pass_total", "metrics", "hvac_hist"):
            return out
        if name == "unit_log":
            return out.reshape(-1, 8)
        if name == "tails":
            return out.reshape(n, 2, _lib.TAIL_CAP)              # [env][side][slot], each band sorted ascending
        return out.reshape(n, -1) if out.size != n else out

    def write_state(self, name, values):
        v = np.ascontiguousarray(values, _STATE_DTYPES[name])
        self._check(self.lib.sdc_write_state(self._h, name.encode(), _ptr(v), v.nbytes))

    def get_state(self):
        nbytes = self.lib.sdc_state_bytes(self._h)
        blob = np.zeros(nbytes, np.uint8)
        self._check(self.lib.sdc_get_state(self._h, _ptr(blob), nbytes))
        return blob

    def set_state(self, blob):
        blob = np.ascontiguousarray(blob, np.uint8)
        self._check(self.lib.sdc_set_state(self._h, _ptr(blob), blob.nbytes))

    def kernel_times(self):
        """(timed steps, sum k_step ms, sum k_reset ms, max k_step ms) since the last call; needs set_tuning(timing=1)."""
        out = np.zeros(4, np.float64)
        self._check(self.lib.sdc_kernel_times(self._h, _ptr(out)))
        return out

    @property
    def error_flags_ptr(self):
        return self.lib.sdc_error_flags(self._h)

    @property
    def launch_count(self):
        return int(self.lib.sdc_launch_count(self._h))
