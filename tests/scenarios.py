"""TEST INFRASTRUCTURE: scenario bodies shared by the GPU parity tests (CUDA library, -m gpu) and by the CPU
tests of the same device logic through the serial hostsim build."""
import random

import numpy as np

from helpers import check_incremental_state


class Buffers:
    """I/O buffers of the device-pointer API: torch CUDA tensors on the GPU, numpy arrays for the hostsim."""

    def __init__(self, n, cuda):
        self.cuda = cuda
        if cuda:
            import torch
            self.torch = torch
            dev = torch.device("cuda:0")
            z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt, device=dev)      # noqa: E731
            self.obs, self.share, self.rew = z(n, 3, 26), z(n, 29), z(n, 3)
            self.done, self.info, self.term = z(n, dt=torch.uint8), z(64, n), z(n, 3, 26)
        else:
            z = lambda *s, dt=np.float32: np.zeros(s, dt)                                # noqa: E731
            self.obs, self.share, self.rew = z(n, 3, 26), z(n, 29), z(n, 3)
            self.done, self.info, self.term = z(n, dt=np.uint8), z(64, n), z(n, 3, 26)

    def actions(self, a):
        a = np.ascontiguousarray(a, np.int32)
        self._a = self.torch.tensor(a, device="cuda:0") if self.cuda else a
        return self._a

    def sync(self):
        if self.cuda:
            self.torch.cuda.synchronize()

    def np(self, x):
        return x.cpu().numpy() if self.cuda else x


def _flat_dc_cfg(geometry):
    """Flat (oracle) form of dc_config.synthetic_dc_config(*geometry); None = the default file."""
    if geometry is None:
        return None
    from dc_rl_b200.dc_config import synthetic_dc_config
    return {k: v for section in synthetic_dc_config(*geometry).values() for k, v in section.items()}


def _oracle_rollout(loc, month, days, seed, n_steps, geometry=None):
    """One oracle env under seeded RNGs; returns what is needed to replay it on the device."""
    import sdc_oracle
    from helpers import oracle_traces
    env = sdc_oracle.OracleEnv(oracle_traces(loc), loc, month, days, dc_cfg=_flat_dc_cfg(geometry))
    random.seed(seed); np.random.seed(seed)
    rng = np.random.RandomState(seed + 77)
    t_ep = days * 96
    rec = dict(resets=[], obs=[], rew=[], energy=[], actions=rng.randint(0, 3, size=(n_steps, 3)).astype(np.int32), done=[])

    def note_reset(o):
        t0 = env.t
        rec["resets"].append(dict(day=env.day, hour=env.hour, temp=env.temp[t0:t0 + t_ep + 18].copy(),
                                  wetb=env.wetb[t0:t0 + t_ep + 18].copy(), tmin=env.t_min, tmax=env.t_max,
                                  obs=np.concatenate([np.pad(o[k], (0, 26 - len(o[k]))) for k in ("agent_ls", "agent_dc", "agent_bat")])))
    note_reset(env.reset())
    for s in range(n_steps):
        o, r, term, info = env.step(*[int(x) for x in rec["actions"][s]])
        rec["obs"].append(np.concatenate([np.pad(o[k], (0, 26 - len(o[k]))) for k in ("agent_ls", "agent_dc", "agent_bat")]))
        rec["rew"].append(r); rec["energy"].append(info["bat_total_energy_with_battery_KWh"]); rec["done"].append(term)
        if term:
            note_reset(env.reset())
    return rec


def batched_mixed_locations_vs_oracle(lib, cuda, N=4096):
    """N = 4096 envs (BASELINE config 2 scale; the mix of BASELINE config 4) over {ny, az, wa} x months x three data-centre
    geometries (default 20 x 200, 25 x 200, 12 x 160 -- per-env cfg_id), device tensors through sdc_step: env i replays
    oracle rollout i mod K (K seeded oracle envs stepped on the CPU), with auto-resets."""
    from dc_rl_b200 import info_layout
    from dc_rl_b200.dc_config import size_datacenter, synthetic_dc_config
    from dc_rl_b200.engine import Engine
    from replay import location_traces, scaled_err
    locs = ["ny", "az", "wa"]
    geoms = [None, (5, 5, 200), (3, 4, 160)]
    days, n_steps, K = 1, 230, 9
    months = [0, 6, 9, 3, 7, 11, 1, 5, 8]
    combo = [(k % 3, (k // 3 + k) % 3) for k in range(K)]         # (location, geometry) of rollout k: all nine pairs
    rollouts = [_oracle_rollout(locs[combo[k][0]], months[k], days, 100 + k, n_steps, geoms[combo[k][1]]) for k in range(K)]
    params = [size_datacenter(locs[l], None if geoms[g] is None else synthetic_dc_config(*geoms[g]))[0] for l, g in combo]
    which = np.arange(N) % K
    eng = Engine(N, [location_traces(l) for l in locs], params, loc_id=np.array([c[0] for c in combo])[which], cfg_id=which,
                 months=0, days_per_episode=days, lib=lib)
    B = Buffers(N, cuda)
    obs, share, rew, done, info, term = B.obs, B.share, B.rew, B.done, B.info, B.term

    def stage(ep):
        for k in range(K):
            if ep < len(rollouts[k]["resets"]):
                r = rollouts[k]["resets"][ep]
                ids = np.nonzero(which == k)[0].astype(np.int32)
                n = len(ids)
                eng.stage_episode(ids, [r["day"]] * n, [r["hour"]] * n, np.repeat(r["temp"][None], n, 0),
                                  np.repeat(r["wetb"][None], n, 0), [r["tmin"]] * n, [r["tmax"]] * n)
    assert len(set(combo)) == 9
    stage(0)
    eng.reset_device(obs, share)
    B.sync()
    ref0 = np.stack([rollouts[k]["resets"][0]["obs"] for k in range(K)])[which].reshape(N, 3, 26)
    assert scaled_err(B.np(obs), ref0) <= 1e-6
    ep, worst_o, worst_r, worst_e = 0, 0.0, 0.0, 0.0
    acts = np.stack([r["actions"] for r in rollouts])       # [K, steps, 3]
    for s in range(n_steps):
        if (s + 1) % (days * 96) == 0:
            stage(ep + 1)
        a = B.actions(acts[which, s])
        eng.step_device(a, obs, share, rew, done, info, term)
        B.sync()
        d = bool(rollouts[0]["done"][s])
        assert (B.np(done) == int(d)).all()
        cur = B.np(term if d else obs)
        ref_o = np.stack([r["obs"][s] for r in rollouts])[which].reshape(N, 3, 26)
        worst_o = max(worst_o, scaled_err(cur, ref_o))
        ref_r = np.array([r["rew"][s] for r in rollouts], np.float64)[which]
        worst_r = max(worst_r, scaled_err(B.np(rew), ref_r))
        ref_e = np.array([r["energy"][s] for r in rollouts])[which]
        worst_e = max(worst_e, scaled_err(B.np(info)[info_layout.COL["bat_total_energy_with_battery_KWh"]], ref_e))
        if d:
            ep += 1
            ref_n = np.stack([r["resets"][ep]["obs"] for r in rollouts])[which].reshape(N, 3, 26)
            assert scaled_err(B.np(obs), ref_n) <= 1e-6
    assert worst_o <= 1e-6 and worst_e <= 1e-6 and worst_r <= 1e-4, (worst_o, worst_e, worst_r)
    assert not eng.read_state("err").any()
    m = eng.metrics()
    assert m[9] == N * n_steps and m[11] == N * ep


def device_generated_resets_match_host_statement(lib):
    """Generated mode (no staged episodes): start day/hour and the weather random walk come from the device
    Philox generator; the serial host statement of the same generator (tests/hostsim) must agree, and the
    two engines must then produce the same trajectory under the same actions."""
    import hostsim_build
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from replay import location_traces, scaled_err
    N, days = 96, 1
    kw = dict(months=np.arange(N) % 12, seeds=np.arange(N, dtype=np.uint64) * 7919 + 5, days_per_episode=days)
    gpu = Engine(N, [location_traces("ny")], [size_datacenter("ny")[0]], lib=lib, **kw)
    cpu = Engine(N, [location_traces("ny")], [size_datacenter("ny")[0]], lib=hostsim_build.load(), **kw)
    og, _ = gpu.reset_host(); oc, _ = cpu.reset_host()
    assert np.array_equal(gpu.read_state("t0"), cpu.read_state("t0"))
    assert scaled_err(gpu.read_state("weather"), cpu.read_state("weather")) <= 1e-4     # hardware log / sincos on the device
    assert scaled_err(og, oc) <= 1e-4
    lo, hi = gpu.day_lo.astype(int) * 96, gpu.day_hi.astype(int) * 96 + 23 * 4
    t0 = gpu.read_state("t0")
    assert ((t0 >= lo) & (t0 <= hi)).all() and len(np.unique(t0)) > 20
    rng = np.random.RandomState(3)
    worst = 0.0
    for s in range(2 * 96 + 5):                     # two auto-resets
        a = rng.randint(0, 3, size=(N, 3)).astype(np.int32)
        rg = gpu.step_host(a); rc = cpu.step_host(a)
        assert np.array_equal(rg[3], rc[3])
        worst = max(worst, scaled_err(rg[0], rc[0]), scaled_err(rg[2], rc[2]))
    # weather noise feeds discontinuous features (clip, first-peak index) through fp32 transcendentals that
    # differ in the last ulp between device and host libm, hence the looser bar of this generator test
    assert worst <= 2e-3, worst
    assert np.array_equal(gpu.read_state("episode"), cpu.read_state("episode")) and (gpu.read_state("episode") == 3).all()


def rolling_quartiles_with_ties_and_small_windows(lib):
    """Window sizes far below 10 000 wrap many times; heavy ties (quantised energies) stress the rank-contiguous
    bracket invariant.  After every few hundred steps the incremental brackets must equal a full sort."""
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from replay import location_traces
    N = 64
    for cap in (8, 64, 1000):
        eng = Engine(N, [location_traces("ny")], [size_datacenter("ny")[0]], months=np.arange(N) % 12,
                     days_per_episode=2, hist_cap=cap, lib=lib)
        eng.reset_host()
        rng = np.random.RandomState(cap)
        for s in range(700):
            eng.step_host(rng.randint(0, 3, size=(N, 3)).astype(np.int32), want_info=False, want_term=False)
            if s % 97 == 0 or s == 699:
                check_incremental_state(eng, tag=(cap, s))
        assert not eng.read_state("err").any()


def incremental_normaliser_under_drift(lib, steps=1300, N=48, cap=2000, names=("drift", "heavy", "ties")):
    """The incremental reward normaliser (brackets + moments + tail sets, sdc_core.h) against a direct numpy statement
    of utils/reward_creator.py:16-45 on the env's window, under conditions that force its refresh machinery: a
    pre-filled window far from the real energies (drift: re-centring, moving fences), heavy tails (tail sets
    overflow -> slack adapts / plain passes), heavy ties, and a wrapping window."""
    from dc_rl_b200 import info_layout
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from replay import location_traces
    col_e, col_ci = info_layout.COL["bat_total_energy_with_battery_KWh"], info_layout.COL["norm_CI"]
    stats = {}
    for name in names:
        eng = Engine(N, [location_traces("ny")], [size_datacenter("ny")[0]], months=np.arange(N) % 12, days_per_episode=3,
                     hist_cap=cap, lib=lib)
        rng = np.random.RandomState(hash(name) % 1000)
        if name == "drift":
            pre = 150.0 + 15.0 * rng.standard_normal((N, cap))
        elif name == "heavy":
            pre = 400.0 + 30.0 * rng.standard_t(1.5, size=(N, cap))
        else:
            pre = np.round(380.0 + 60.0 * rng.standard_normal((N, cap)) / 25.0) * 25.0
        eng.prefill_history(pre.astype(np.float32))
        eng.reset_host()
        worst, scans = 0.0, np.zeros(2, np.int64)
        for s in range(steps):
            obs, share, rew, done, info, _ = eng.step_host(rng.randint(0, 3, size=(N, 3)).astype(np.int32))
            scans += eng.read_state("pass_stats")[:2]
            if s % 61 == 0 or s >= steps - 3:
                check_incremental_state(eng, tag=(name, s))
                hist = eng.read_state("hist").astype(np.float64) + eng.read_state("hist_ref").reshape(-1, 1)   # window values are stored relative to the env's first sample
                e, nci = info[col_e].astype(np.float64), info[col_ci].astype(np.float64)
                for i in range(N):
                    w = hist[i]
                    q1, q3 = np.percentile(w, 25), np.percentile(w, 75)
                    cl = np.clip(w, q1 - 1.5 * (q3 - q1), q3 + 1.5 * (q3 - q1))
                    sd = cl.std()
                    z = (e[i] - cl.mean()) / (sd if sd > 0 else 1.0)
                    exp = -(nci[i] * z / 0.5)
                    worst = max(worst, abs(rew[i, 1] - exp) / max(1.0, abs(exp)))
        assert worst <= 1e-4, (name, worst)
        assert not eng.read_state("err").any(), name
        stats[name] = dict(worst=worst, plain=int(scans[0]), refresh=int(scans[1]), env_steps=N * steps,
                           valid=int((eng.read_state("tail_n").reshape(N, 2)[:, 0] >= 0).sum()))
    # the point of the design: in the well-behaved case almost no env-step needs a window pass
    if "drift" in stats:
        assert stats["drift"]["plain"] + stats["drift"]["refresh"] < 0.1 * stats["drift"]["env_steps"], stats
    return stats


def prefill_and_constant_history_branches(lib):
    """normalize_energy edge cases (utils/reward_creator.py:26-27,43-45): fewer than two samples -> 0;
    zero spread -> divide by 1.  Prefilled constant windows give q1 == q3 exactly."""
    from dc_rl_b200 import info_layout
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from replay import location_traces
    N = 40
    eng = Engine(N, [location_traces("ny")], [size_datacenter("ny")[0]], months=6, days_per_episode=1, lib=lib)
    eng.reset_host()
    a = np.ones((N, 3), np.int32)
    obs, share, rew, done, info, _ = eng.step_host(a)
    assert (rew[:, 1] == 0).all() and (rew[:, 2] == 0).all()          # first sample: z = 0
    eng.prefill_history(np.full(5000, 256.0, np.float32))
    obs, share, rew, done, info, _ = eng.step_host(a)
    e = info[info_layout.COL["bat_total_energy_with_battery_KWh"]].astype(np.float64)
    nci = info[info_layout.COL["norm_CI"]].astype(np.float64)
    expect = -(nci * (e - 256.0) / 0.5)                                  # sigma == 0 -> divide by 1
    assert np.max(np.abs(rew[:, 1] - expect) / np.maximum(1, np.abs(expect))) <= 1e-5
    assert not eng.read_state("err").any()


def state_blob_roundtrip(lib):
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from replay import location_traces
    N = 50
    eng = Engine(N, [location_traces("wa")], [size_datacenter("wa")[0]], months=9, days_per_episode=1, lib=lib)
    eng.reset_host()
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 3, size=(60, N, 3)).astype(np.int32)
    for s in range(20):
        eng.step_host(acts[s])
    blob = eng.get_state()
    a = [tuple(x.copy() for x in eng.step_host(acts[s])[:4]) for s in range(20, 60)]
    eng.set_state(blob)
    b = [tuple(x.copy() for x in eng.step_host(acts[s])[:4]) for s in range(20, 60)]
    for x, y in zip(a, b):
        for u, v in zip(x, y):
            assert np.array_equal(u, v)
