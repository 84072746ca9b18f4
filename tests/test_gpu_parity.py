"""GPU parity tests (run with -m gpu on a B200): everything goes through the C ABI of libsdc_b200.so.

Checker = oracle/sdc_oracle.py (numpy fp64 restatement pinned to the live reference) and the golden
live-reference trajectories under tests/golden/.  Tolerance |a-b| <= tol*max(1,|b|): observations and info
1e-6 (fp64 physics on device, fp32 outputs), rewards 1e-4 (north_star bar; fp32 reward window).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from conftest import cuda_lib_or_skip
    return cuda_lib_or_skip()


@pytest.mark.parametrize("name", ["ny_m0_s0", "ny_m3_s1", "az_m6_s2", "wa_m9_s3", "ny_m6_dc25x200",
                                  "ny_m6_tz5", "ny_m2_altA", "az_m8_altB", "wa_m4_altC"])
def test_golden_replay_single_env(lib, name):
    from replay import replay
    w = replay(name, lib)
    assert w["err_flags"] == 0
    assert w["reset_obs"] <= 1e-6 and w["obs"] <= 1e-6 and w["term_obs"] <= 1e-6, w
    assert w["info"] <= 1e-6 and w["share"] == 0.0, w
    assert w["rew"] <= 1e-4, w


@pytest.mark.parametrize("n_envs,unit", [(70, 32), (33, 8), (256, 16)])
def test_golden_replay_batched_ragged(lib, n_envs, unit):
    """Ragged batch sizes (not a multiple of the warp unit) and all unit sizes: every env replays the
    same golden trajectory and must agree with it."""
    from replay import replay
    w = replay("wa_m9_s3", lib, n_envs=n_envs, unit_envs=unit)
    assert w["err_flags"] == 0
    assert w["obs"] <= 1e-6 and w["term_obs"] <= 1e-6 and w["info"] <= 1e-6 and w["rew"] <= 1e-4, w


def test_golden_replay_more_units_than_one_resident_round(lib):
    """20 000 envs at 8 envs per warp are 2 500 units: more than the 2 072 unit warps of one resident grid, so the first
    round takes its units from the grid slots and the rest from the ticket counter.  Every env replays the same golden
    trajectory (first episode boundary included) and must agree with it."""
    from replay import replay
    w = replay("wa_m9_s3", lib, n_envs=20000, unit_envs=8, max_steps=200)
    assert w["err_flags"] == 0
    assert w["obs"] <= 1e-6 and w["term_obs"] <= 1e-6 and w["info"] <= 1e-6 and w["rew"] <= 1e-4, w


def test_long_replay_window_saturates(lib):
    """11 000 steps: the 10 000-sample reward window fills, wraps, and the rolling quartile brackets stay exact."""
    from replay import replay
    w = replay("ny_m6_long", lib, n_envs=3, compact=True)
    assert w["err_flags"] == 0
    assert w["rew"] <= 1e-4 and w["info"] <= 1e-6, w
    eng = w["engine"]
    assert (eng.read_state("hist_len") == 10000).all()
    # brackets maintained incrementally == brackets from a full sort of the window
    from helpers import check_incremental_state
    valid, _ = check_incremental_state(eng, tag="long replay")
    assert valid == 3                       # the tail sets are in use at the end of the run
    hist = np.sort(eng.read_state("hist"), axis=1)
    eng.rebuild_brackets()
    ql2, a2, m2 = eng.read_state("qlist").reshape(3, 2, -1), eng.read_state("q_a"), eng.read_state("q_m")
    for e in range(3):
        for j in range(2):
            assert np.array_equal(ql2[e, j, :m2[e, j]], hist[e, a2[e, j]:a2[e, j] + m2[e, j]])


def test_batched_mixed_locations_vs_oracle(lib):
    """N = 4096 envs (BASELINE config 2 scale) over {ny, az, wa} x months, device tensors through sdc_step: env i
    replays oracle rollout i mod K (K seeded oracle envs stepped on the CPU), with auto-resets."""
    import scenarios
    scenarios.batched_mixed_locations_vs_oracle(lib, cuda=True, N=4096)


def test_device_generated_resets_match_host_statement(lib):
    import scenarios
    scenarios.device_generated_resets_match_host_statement(lib)


def test_rolling_quartiles_with_ties_and_small_windows(lib):
    import scenarios
    scenarios.rolling_quartiles_with_ties_and_small_windows(lib)


def test_incremental_normaliser_under_drift(lib):
    import scenarios
    print(scenarios.incremental_normaliser_under_drift(lib))


def test_incremental_normaliser_full_window_heavy_tails_and_ties(lib):
    """The same check at the full window length (H = 10 000) for the two distributions that flood a window pass: heavy
    tails (thousands of values beyond the band thresholds) and heavy ties (thousands of equal values inside a re-centring
    interval) -- more hits than a warp's slice of the parked-hit list holds, i.e. the pass's second classification loop."""
    from scenarios import incremental_normaliser_under_drift
    stats = incremental_normaliser_under_drift(lib, steps=200, N=16, cap=10000, names=("heavy", "ties"))
    assert stats["heavy"]["refresh"] + stats["heavy"]["plain"] > 0 and stats["ties"]["refresh"] + stats["ties"]["plain"] > 0, stats


def test_prefill_and_constant_history_branches(lib):
    import scenarios
    scenarios.prefill_and_constant_history_branches(lib)


def test_state_blob_roundtrip(lib):
    import scenarios
    scenarios.state_blob_roundtrip(lib)


def test_host_buffer_modes_agree(lib):
    """sdc_step_host on the handle's pinned buffers: every output stored by the kernel straight into host memory and the
    actions read from there (direct_host = 31, default), observations only (3), rewards / dones too (15), staged device
    buffers + copies (0), and caller-owned pageable arrays all return the same step results."""
    import ctypes as C
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine, _ptr
    from replay import location_traces
    N = 700
    outs = []
    for mode in (31, 3, 15, 0, "own"):
        eng = Engine(N, [location_traces("az")], [size_datacenter("az")[0]], months=np.arange(N) % 12,
                     seeds=np.arange(N, dtype=np.uint64) + 3, days_per_episode=1, lib=lib)
        if mode != "own":
            eng.set_tuning(direct_host=mode)
        eng.reset_host()
        rng = np.random.RandomState(9)
        rec = []
        own = dict(obs=np.full((N, 3, 26), 7.0, np.float32), share=np.zeros((N, 29), np.float32), rew=np.zeros((N, 3), np.float32),
                   done=np.zeros(N, np.uint8))
        for s in range(100):
            a = rng.randint(0, 3, size=(N, 3)).astype(np.int32)
            if mode == "own":
                rc = lib.sdc_step_host(eng._h, _ptr(a), _ptr(own["obs"]), _ptr(own["share"]), _ptr(own["rew"]), _ptr(own["done"]), None, None)
                assert rc == 0
                o, sh, r, d = own["obs"], own["share"], own["rew"], own["done"]
            else:
                o, sh, r, d, _, _ = eng.step_host(a, want_info=False, want_term=False)
            if s % 9 == 0 or s > 95:
                rec.append([x.copy() for x in (o, sh, r, d)])
        outs.append(rec)
        eng.close()
    for other in outs[1:]:
        for ra, rb in zip(outs[0], other):
            for xa, xb in zip(ra, rb):
                assert np.array_equal(xa, xb)


def test_run_to_run_bit_reproducible(lib):
    """Two handles built and driven identically (generated-mode resets, look-ahead generation, maintenance passes and the
    passes that price a step served by whichever CTA is free) return bit-identical observations, rewards, dones and info at
    every step: nothing a step returns depends on which CTA ran a job or when."""
    import torch
    import bench
    n, steps = 8192, 260
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev); g.manual_seed(7)
    acts = [torch.randint(0, 3, (n, 3), dtype=torch.int32, device=dev, generator=g) for _ in range(16)]
    st = torch.cuda.current_stream().cuda_stream
    runs = []
    for rep in range(2):
        eng, _ = bench.build_engine(n, 0)
        bench.prepare(eng, n, 0)            # pre-filled windows: the first step runs 8 192 window passes that price their step
        obs = torch.zeros(n, 3, 26, device=dev); share = torch.zeros(n, 29, device=dev); rew = torch.zeros(n, 3, device=dev)
        done = torch.zeros(n, dtype=torch.uint8, device=dev); info = torch.zeros(64, n, device=dev)
        rec = []
        for s in range(steps):
            eng.step_device(acts[s % 16], obs, share, rew, done, info, None, st)
            if s < 4 or s % 13 == 0 or s > steps - 4:
                rec.append([x.clone() for x in (obs, share, rew, done, info)])
        torch.cuda.synchronize()
        assert int(np.bitwise_or.reduce(eng.read_state("err"))) == 0
        runs.append(rec)
        eng.close()
    for ra, rb in zip(*runs):
        for xa, xb in zip(ra, rb):
            assert torch.equal(xa, xb)


def test_full_size_properties(lib):
    """BASELINE size (N = 65 536, H = 10 000 pre-filled) through size-independent properties: envs j and j + N/2 are given
    the same seed, month, window and actions, so they must produce bit-identical outputs although different warps / CTAs /
    worker CTAs handle them; the device-side logger sums equal the sums of the per-env outputs; the incremental normaliser
    state of a sample of envs equals what their windows say."""
    import torch
    from helpers import check_incremental_state
    from dc_rl_b200.dc_config import size_datacenter
    from dc_rl_b200.engine import Engine
    from dc_rl_b200.traces import LocationTraces
    N, H, steps = 65536, 10000, 40
    half = N // 2
    ids = np.arange(N) % half
    eng = Engine(N, [LocationTraces.synthetic("ny", 1234)], [size_datacenter("ny")[0]], months=ids % 12,
                 seeds=ids.astype(np.uint64) * 1000 + 17, days_per_episode=1, lib=lib)
    rng = np.random.default_rng(5)
    pool = (330.0 + 40.0 * rng.standard_normal((128, H), dtype=np.float32)).astype(np.float32)
    hist = pool[rng.integers(0, 128, half)] + rng.standard_normal((half, 1), dtype=np.float32)
    eng.prefill_history(np.concatenate([hist, hist]))
    del hist
    dev = torch.device("cuda:0")
    obs = torch.zeros(N, 3, 26, device=dev); share = torch.zeros(N, 29, device=dev); rew = torch.zeros(N, 3, device=dev)
    done = torch.zeros(N, dtype=torch.uint8, device=dev)
    eng.reset_device(obs, share)
    phase = np.tile(rng.integers(0, 90, half).astype(np.int32), 2)            # de-synchronised episodes: resets in every step
    eng.write_state("step_in_ep", phase)
    eng.write_state("t", eng.read_state("t0").astype(np.int32) + phase)
    eng.metrics(clear=True)
    g = torch.Generator(device=dev); g.manual_seed(3)
    rew_sum, n_done = 0.0, 0
    for s in range(steps):
        a = torch.randint(0, 3, (half, 3), dtype=torch.int32, device=dev, generator=g).repeat(2, 1).contiguous()
        eng.step_device(a, obs, share, rew, done)
        torch.cuda.synchronize()
        assert torch.equal(obs[:half], obs[half:]) and torch.equal(rew[:half], rew[half:]) and torch.equal(done[:half], done[half:])
        assert torch.equal(share[:half], share[half:])
        rew_sum += float(rew.double().sum()); n_done += int(done.sum())
    m = eng.metrics()
    assert m[9] == N * steps and m[11] == n_done and n_done > 0
    assert abs(m[10] - rew_sum) <= 1e-9 * max(1.0, abs(rew_sum))
    assert not eng.read_state("err").any()
    valid, checked = check_incremental_state(eng, envs=list(range(0, N, 4099)), tag="full size")
    assert valid == checked
