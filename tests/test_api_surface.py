"""Tests of the reference-facing Python surface (CudaShareVecEnv / InfoBatch / SustainDC / make_*_env / multi-process
sharding).  Every test that takes `lib` runs twice: on the serial hostsim build of the device logic (CPU suite) and, marked
`gpu`, on libsdc_b200.so itself.  Plus the check that the product library loads and exports every symbol
include/sdc_b200.h declares."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import sdc_oracle
from conftest import lib_params, resolve_lib
from helpers import load_traj, oracle_traces, traj_cfg

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", params=lib_params())
def lib(request):
    request.module._LIB_KIND = request.param
    return resolve_lib(request.param)


def test_library_exports_every_declared_symbol():
    """No compute calls (there is no GPU here): load libsdc_b200.so and resolve the whole header."""
    import ctypes
    from dc_rl_b200 import _lib
    header = open(os.path.join(REPO, "include", "sdc_b200.h")).read()
    declared = set(re.findall(r"\b(sdc_[a-z_0-9]+)\s*\(", header))
    declared -= {"sdc_env", "sdc_config", "sdc_location", "sdc_dc_params"}
    assert os.path.isfile(_lib.LIB_PATH), "build first: python __graft_entry__.py"
    so = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(so, name), name
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert so.sdc_abi_version() == _lib.ABI_VERSION


def test_missing_library_fails_loudly(tmp_path):
    from dc_rl_b200 import _lib
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "libsdc_b200.so"))


def test_product_package_never_imports_oracle_or_hostsim():
    pkg = os.path.join(REPO, "dc_rl_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".inc")):
                src = open(os.path.join(root, f)).read()
                assert "sdc_oracle" not in src and "import oracle" not in src, f
                if f.endswith(".py"):
                    assert "hostsim" not in src.replace("tests/hostsim", ""), f


def _vec_from_golden(lib, name, n):
    from dc_rl_b200.vec_env import CudaShareVecEnv
    from replay import location_traces, stage
    g = load_traj(name)
    cfg = traj_cfg(g)
    args = dict(cfg, traces=location_traces(cfg["location"]), nonoverlapping_shared_obs_space=True)
    v = CudaShareVecEnv(args, n, lib=lib)
    return g, cfg, v, stage


def test_vec_env_matches_oracle_harl_view_with_auto_reset(lib):
    """ShareVecEnv semantics (env_wrappers.py:173-192 over harlsustaindc_env.py:42-131) against the oracle's
    restatement: shapes, dtypes, share_obs, rewards [N,3,1], dones [N,3], original_obs on auto-reset."""
    n = 5
    g, cfg, v, stage = _vec_from_golden(lib, "wa_m9_s3", n)
    t_ep = cfg["days_per_episode"] * 96
    ora = sdc_oracle.OracleEnv(oracle_traces(cfg["location"]), cfg["location"], cfg["month"], cfg["days_per_episode"])

    def inject(k):
        ora.inject_episode(int(g["reset_day"][k]), int(g["reset_hour"][k]), g["reset_temp"][k], g["reset_wetb"][k],
                           float(g["reset_t_min30"][k]), float(g["reset_t_max30"][k]))
        stage(v.engine, g, k, np.arange(n, dtype=np.int32))
    inject(0)
    obs, share, avail = v.reset()
    rows, sh = sdc_oracle.harl_view(ora.reset())
    assert obs.shape == (n, 3, 26) and obs.dtype == np.float32 and share.shape == (n, 3, 29) and avail.shape == (n, 3, 3)
    assert np.array_equal(obs[2], rows) and np.array_equal(share[4], sh)
    k = 0
    for s in range(t_ep + 30):
        if s == t_ep - 1:
            inject(1)
        a = g["actions"][s].astype(np.int64)
        obs, share, rew, dones, infos, avail = v.step(np.broadcast_to(a.reshape(1, 3, 1), (n, 3, 1)))
        o, r, term, info = ora.step(*[int(x) for x in a])
        rows, sh = sdc_oracle.harl_view(o)
        assert rew.shape == (n, 3, 1) and dones.shape == (n, 3) and dones.dtype == bool and len(infos) == n
        assert np.allclose(rew[1, :, 0], r, rtol=0, atol=1e-4)
        assert dones.all() == term and dones.any() == term
        if term:
            assert np.array_equal(infos[3][0]["original_obs"], rows)
            assert np.array_equal(infos[3][0]["original_state"], sh)
            assert infos[3][0]["original_avail_actions"].shape == (3, 3)
            rows, sh = sdc_oracle.harl_view(ora.reset())
        else:
            assert "original_obs" not in infos[3][0].keys()
        assert np.array_equal(obs[0], rows) and np.array_equal(share[n - 1], sh)
        assert abs(infos[2][0]["bat_total_energy_with_battery_KWh"] - info["bat_total_energy_with_battery_KWh"]) < 1e-3
        assert infos[1][2].get("bat_a_t") == info["bat_a_t"] and infos[0][1].get("not_a_key", 7) == 7
        assert bool(infos[0][0]["isterminal"]) == term
    v.close()


def test_info_rows_expose_every_key_the_runners_read(lib):
    g, cfg, v, stage = _vec_from_golden(lib, "ny_m3_s1", 2)
    stage(v.engine, g, 0, np.arange(2, dtype=np.int32))
    v.reset()
    _, _, _, _, infos, _ = v.step(np.ones((2, 3, 1), np.int64))
    runner_keys = [   # harl/runners/on_policy_base_runner.py:617-638 and harl/envs/sustaindc/sustaindc_logger.py:87-99
        'ls_original_workload', 'ls_shifted_workload', 'ls_action', 'ls_norm_load_left', 'ls_unasigned_day_load_left',
        'ls_penalty_flag', 'ls_tasks_in_queue', 'ls_tasks_dropped', 'ls_current_hour', 'dc_ITE_total_power_kW',
        'dc_HVAC_total_power_kW', 'dc_total_power_kW', 'dc_power_lb_kW', 'dc_power_ub_kW', 'dc_crac_setpoint_delta',
        'dc_crac_setpoint', 'dc_cpu_workload_fraction', 'dc_int_temperature', 'dc_CW_pump_power_kW', 'dc_CT_pump_power_kW',
        'dc_water_usage', 'dc_exterior_ambient_temp', 'outside_temp', 'day', 'hour', 'bat_action', 'bat_SOC',
        'bat_CO2_footprint', 'bat_avg_CI', 'bat_total_energy_without_battery_KWh', 'bat_total_energy_with_battery_KWh',
        'bat_max_bat_cap', 'bat_dcload_min', 'bat_dcload_max', 'dc_CT_total_power_kW', 'dc_Compressor_total_power_kW']
    for agent in range(3):
        row = infos[1][agent]
        for key in runner_keys:
            assert key in row.keys() and row.get(key, None) is not None and np.isfinite(row[key]), key
    assert len(infos[0][0]["ls_task_age_histogram"]) == 5 and len(infos[0][0]["forecast_CI"]) == 8
    assert "bad_transition" not in infos[0][0].keys()
    assert infos.column("dc_water_usage").shape == (2,)
    v.close()


def test_hvac_histogram_matches_sample_percentile(lib):
    """the logger's mean / max / p90 of the positive HVAC power samples come from a fixed-bin device histogram."""
    g, cfg, v, stage = _vec_from_golden(lib, "ny_m3_s1", 8)
    stage(v.engine, g, 0, np.arange(8, dtype=np.int32))
    v.reset()
    rng = np.random.RandomState(3)
    samples = []
    for _ in range(60):
        _, _, _, _, infos, _ = v.step(rng.randint(0, 3, size=(8, 3, 1)))
        samples.append(np.array(infos.column("dc_HVAC_total_power_kW"), np.float64))
    samples = np.concatenate(samples)
    samples = samples[samples > 0]
    counts, range_kw = v.engine.hvac_histogram()
    width = range_kw / len(counts)
    assert int(counts.sum()) == len(samples) and range_kw > samples.max()
    st = v.hvac_power_stats()
    assert abs(st["p90"] - np.percentile(samples, 90)) <= 2 * width
    assert abs(st["mean"] - samples.mean()) <= width and 0 <= st["max"] - samples.max() <= width
    v.engine.hvac_histogram(clear=True)
    assert v.engine.hvac_histogram()[0].sum() == 0
    v.close()


def test_small_rack_geometry_trips_the_outlet_guard(lib):
    """BASELINE config 3 names 25 racks x 40 CPUs.  The reference refuses that geometry: racks that small heat the air by
    less than 2 C and `compute_datacenter_IT_load_outlet_temp` raises (envs/datacenter.py:295-300; reproduced with the live
    reference while minting the goldens).  Here the same condition sets SDC_F_OUTLET_DELTA on the env and the batch goes
    on; 25 racks x 200 CPUs is the non-default geometry pinned against the live reference (traj_ny_m6_dc25x200)."""
    from dc_rl_b200.dc_config import size_datacenter, synthetic_dc_config
    from dc_rl_b200.engine import Engine
    from replay import location_traces
    for cpus, flagged in ((40, True), (200, False)):
        params, _ = size_datacenter("ny", synthetic_dc_config(5, 5, cpus))
        eng = Engine(4, [location_traces("ny")], [params], months=6, days_per_episode=1, lib=lib)
        eng.reset_host()
        for _ in range(3):
            eng.step_host(np.ones((4, 3), np.int32))
        assert bool((eng.read_state("err") & 0x4).all()) == flagged
        eng.close()


def test_compact_async_lazy_and_view_paths_agree_with_the_plain_call(lib):
    """sdc_step_compact_host (the 29 distinct observation values per env) + sdc_expand_obs, the begin / end pair behind step_async /
    step_wait, the lazy info table (sdc_fetch_info) and the zero-copy `output_views` mode all return what the plain
    padded, eager, copying call returns."""
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from dc_rl_b200.vec_env import CudaShareVecEnv
    from replay import location_traces
    N, T = 70, 110
    kw = dict(months=np.arange(N) % 12, seeds=np.arange(N, dtype=np.uint64) + 3, days_per_episode=1, lib=lib)
    a_eng = Engine(N, [location_traces("az")], [size_datacenter("az")[0]], **kw)
    c_eng = Engine(N, [location_traces("az")], [size_datacenter("az")[0]], **kw)
    a_eng.reset_host(); c_eng.reset_host()
    rng = np.random.RandomState(2)
    n_done = 0
    for s in range(T):
        act = rng.randint(0, 3, size=(N, 3)).astype(np.int32)
        obs, share, rew, done, info, term = a_eng.step_host(act)
        oc, r2, d2, i2, tc = c_eng.step_compact_host(act, want_info=True)
        assert np.array_equal(rew, r2) and np.array_equal(done, d2) and np.array_equal(info, i2)
        eo, es = c_eng.expand_obs(oc)
        assert np.array_equal(eo, obs) and np.array_equal(es, share)
        fin = np.nonzero(done)[0]
        if len(fin):
            n_done += len(fin)
            assert np.array_equal(c_eng.expand_obs(tc[fin], want_share=False)[0], term[fin])
    assert n_done >= N
    a_eng.close(); c_eng.close()
    # vec-env level: eager + copies (reference-like) vs lazy info + views + step_async / step_wait
    args = {"location": "ny", "traces": location_traces("ny"), "days_per_episode": 1, "nonoverlapping_shared_obs_space": True}
    plain = CudaShareVecEnv(dict(args, info="eager"), 9, seed=1, lib=lib)
    fast = CudaShareVecEnv(dict(args, output_views=True), 9, seed=1, lib=lib)
    o1, s1, _ = plain.reset(); o2, s2, _ = fast.reset()
    assert np.array_equal(o1, o2) and np.array_equal(s1, s2)
    for s in range(100):
        act = rng.randint(0, 3, size=(9, 3, 1))
        r1 = plain.step(act)
        fast.step_async(act)
        r2 = fast.step_wait()
        for x, y in zip(r1[:4], r2[:4]):
            assert x.shape == y.shape and x.dtype == y.dtype and np.array_equal(x, y)
        assert np.array_equal(r1[4].column("dc_water_usage"), r2[4].column("dc_water_usage"))
        assert r1[4][3][1]["bat_SOC"] == r2[4][3][1]["bat_SOC"] and len(r2[4]) == 9
        if r1[3].any():
            assert np.array_equal(r1[4][2][0]["original_obs"], r2[4][2][0]["original_obs"])
    touched = fast.step(act)[4]
    touched.column("bat_SOC")                    # read while current: stays readable afterwards
    stale = fast.step(act)[4]                    # never read ...
    fast.step(act)
    assert touched.column("bat_SOC").shape == (9,)
    with pytest.raises(RuntimeError, match="earlier step"):
        stale.column("bat_SOC")                  # ... and its table has been overwritten on the device by now
    plain.close(); fast.close()


def test_make_env_month_and_seed_rules(lib):
    """harl/utils/envs_tools.py:56-67,95: month by rank unless pinned; seed + rank*1000 / seed*50000 + rank*10000."""
    from dc_rl_b200.dc_config import start_day_range
    from dc_rl_b200.harl_env import make_eval_env, make_train_env
    args = {"location": "az", "traces": "synthetic", "days_per_episode": 1, "_lib": lib}
    v = make_train_env("sustaindc", 3, 20, dict(args))
    months = [r % 12 if r < 12 else r % 3 + 5 for r in range(20)]
    assert list(v.engine.day_lo) == [start_day_range(m)[0] for m in months]
    assert list(v.engine.seeds) == [3 + r * 1000 for r in range(20)]
    v.close()
    v = make_train_env("sustaindc", 3, 4, dict(args, month=6))
    assert list(v.engine.day_lo) == [start_day_range(6)[0]] * 4
    v.close()
    v = make_eval_env("sustaindc", 2, 3, dict(args))
    assert list(v.engine.seeds) == [100000, 110000, 120000]
    v.close()
    with pytest.raises(NotImplementedError):
        make_train_env("smac", 0, 1, dict(args))


def test_single_env_gym_surface(lib):
    from dc_rl_b200.sustaindc_env import SustainDC
    from replay import location_traces
    with pytest.raises(TypeError):
        SustainDC({"location": "ny", "traces": "synthetic"}, lib=lib)
    env = SustainDC({"location": "ny", "month": 3, "days_per_episode": 1, "traces": location_traces("ny")}, lib=lib)
    obs = env.reset()
    assert set(obs) == {"agent_ls", "agent_dc", "agent_bat"} and obs["agent_dc"].shape == (14,)
    for s in range(96):
        obs, rew, term, trunc, info = env.step({"agent_ls": 0, "agent_dc": 2, "agent_bat": 0})
        assert not term["__all__"] and trunc["__all__"] == (s == 95)
        assert set(info) == {"agent_ls", "agent_dc", "agent_bat", "__common__"}
    sp = info["__common__"]["dc_crac_setpoint"]
    assert sp == pytest.approx(21.6)                      # accelerating increase run clipped at the maximum
    nxt = env.reset()
    assert nxt["agent_bat"][-1] == 0.0 and nxt["agent_ls"][12] == 0.0     # SoC and queue cleared
    _, _, _, _, info = env.step({"agent_ls": 1, "agent_dc": 1, "agent_bat": 2})
    assert info["__common__"]["dc_crac_setpoint"] == pytest.approx(21.6)   # set-point survives reset
    env.close()


def test_pettingzoo_parallel_env_surface(lib):
    """harl/envs/sustaindc/sustaindc_ptzoo.py:5-101: dict-keyed spaces, reset -> obs dict, step -> five agent-keyed dicts."""
    from dc_rl_b200 import SustainDCPettingZooEnv
    from replay import location_traces
    cfg = {"location": "wa", "month": 9, "days_per_episode": 1, "traces": location_traces("wa"), "partial_obs": True,
           "nonoverlapping_shared_obs_space": True}
    with pytest.raises(NotImplementedError):
        SustainDCPettingZooEnv(dict(cfg, partial_obs=False), lib=lib)
    env = SustainDCPettingZooEnv(cfg, lib=lib)
    assert env.possible_agents == ["agent_ls", "agent_dc", "agent_bat"] and env.share_observation_space["agent_dc"].shape == (29,)
    assert env.observation_space("agent_bat").shape == (13,) and env.action_space("agent_ls").n == 3
    obs = env.reset(seed=3)
    assert set(obs) == set(env.possible_agents) and obs["agent_ls"].shape == (26,)
    for s in range(96):
        o, r, d, t, info = env.step({"agent_ls": 1, "agent_dc": 1, "agent_bat": 2})
        assert set(o) == set(r) == set(d) == set(t) == set(info) == set(env.possible_agents)
        assert not any(d.values()) and all(t.values()) == (s == 95)
        assert info["agent_dc"]["dc_crac_setpoint"] == 18.0           # do-nothing actions keep the set-point
    env.close()


def test_requires_data_or_explicit_synthetic(lib):
    from dc_rl_b200.vec_env import CudaShareVecEnv
    old = os.environ.pop("SDC_DATA_ROOT", None)
    try:
        with pytest.raises(FileNotFoundError):
            CudaShareVecEnv({"location": "ny"}, 1, lib=lib)
    finally:
        if old:
            os.environ["SDC_DATA_ROOT"] = old
    with pytest.raises(AssertionError):       # utils/reward_creator.py:346 asserts on unknown names
        CudaShareVecEnv({"location": "ny", "traces": "synthetic", "ls_reward": "no_such_reward"}, 1, lib=lib)
    with pytest.raises(AssertionError):       # :217 -- needs a key the env never provides
        CudaShareVecEnv({"location": "ny", "traces": "synthetic", "dc_reward": "renewable_energy_reward"}, 1, lib=lib)


_WORKER = r"""
import os, sys
sys.path[:0] = [{repo!r}, {repo!r} + "/tests"]
import numpy as np, torch.distributed as dist
from conftest import resolve_lib
from dc_rl_b200.distributed import gather_metrics, make_sharded_env, reduce_hvac_histogram, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
args = {{"location": ["ny", "az"], "traces": "synthetic", "days_per_episode": 1, "nonoverlapping_shared_obs_space": True}}
env = make_sharded_env(args, 13, seed=5, device=0, lib=resolve_lib({kind!r}))
lo, hi = shard_range(13, rank, 2)
obs, _, _ = env.reset()
rng = np.random.RandomState(0)
acts = rng.randint(0, 3, size=(40, 13, 3))
outs = []
for s in range(40):
    o, _, r, d, _, _ = env.step(acts[s, lo:hi])
    outs.append((o.copy(), r.copy()))
allm = gather_metrics(env.engine.metrics())
hist = reduce_hvac_histogram(env.engine.hvac_histogram()[0])
np.savez({out!r} + "/rank%d.npz" % rank, obs=np.stack([o for o, _ in outs]), rew=np.stack([r for _, r in outs]), metrics=allm, lo=lo, hi=hi,
         hist=hist)
dist.destroy_process_group()
"""


def test_two_rank_sharding_is_equivalent_to_one_process(lib, tmp_path):
    """N envs on one handle == the same env ids split over 2 ranks (gloo): bit-identical per env, and the one
    collective (metric all-gather) sums to the single-process metrics.  Locations alternate per GLOBAL env id, so the
    shards must pick up the right (location, sizing) per env."""
    from dc_rl_b200.distributed import shard_range
    from dc_rl_b200.vec_env import CudaShareVecEnv
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(repo=REPO, port=port, out=str(tmp_path), kind=_LIB_KIND))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    args = {"location": ["ny", "az"], "traces": "synthetic", "days_per_episode": 1, "nonoverlapping_shared_obs_space": True}
    env = CudaShareVecEnv(args, 13, seed=5, lib=lib)
    env.reset()
    rng = np.random.RandomState(0)
    acts = rng.randint(0, 3, size=(40, 13, 3))
    obs, rew = [], []
    for s in range(40):
        o, _, r, _, _, _ = env.step(acts[s])
        obs.append(o.copy()); rew.append(r.copy())
    obs, rew = np.stack(obs), np.stack(rew)
    total = np.zeros(16)
    for r in range(2):
        z = np.load(tmp_path / ("rank%d.npz" % r))
        lo, hi = int(z["lo"]), int(z["hi"])
        assert (lo, hi) == shard_range(13, r, 2)
        assert np.array_equal(z["obs"], obs[:, lo:hi]) and np.array_equal(z["rew"], rew[:, lo:hi])
        assert z["metrics"].shape == (2, 16)
        total = z["metrics"].sum(0)
        assert np.array_equal(z["hist"], env.engine.hvac_histogram()[0])       # all-reduced histogram == single process
    assert np.allclose(total, env.engine.metrics(), rtol=1e-12)
    assert shard_range(10, 0, 3) == (0, 4) and shard_range(10, 2, 3) == (7, 10)
