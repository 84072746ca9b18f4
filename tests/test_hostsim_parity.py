"""CPU tests of the device logic (dc_rl_b200/csrc/sdc_core.h + sdc_api.inc) through the serial hostsim
build: the same golden live-reference trajectories that pin the oracle, replayed through the C ABI.
Tolerance: |a-b| <= tol*max(1,|b|); observations 1e-6 (fp64 physics, fp32 outputs), info 1e-6,
rewards 1e-4 (north_star bar; the reward window is fp32)."""
import numpy as np
import pytest

import hostsim_build
from replay import replay


@pytest.fixture(scope="module")
def lib():
    return hostsim_build.load()


@pytest.mark.parametrize("name", ["ny_m0_s0", "ny_m3_s1", "az_m6_s2", "wa_m9_s3", "ny_m6_dc25x200",
                                  "ny_m6_tz5",                                   # timezone_shift = 5
                                  "ny_m2_altA", "az_m8_altB", "wa_m4_altC"])     # alternate reward methods
def test_replay_matches_live_reference(lib, name):
    w = replay(name, lib)
    assert w["err_flags"] == 0
    assert w["reset_obs"] <= 1e-6 and w["obs"] <= 1e-6 and w["term_obs"] <= 1e-6, w
    assert w["info"] <= 1e-6, w
    assert w["share"] == 0.0
    assert w["rew"] <= 1e-4, w


def test_long_replay_window_saturates(lib):
    w = replay("ny_m6_long", lib, compact=True)
    assert w["err_flags"] == 0
    assert w["rew"] <= 1e-4 and w["info"] <= 1e-6, w
    assert int(w["engine"].read_state("hist_len")[0]) == 10000
    from helpers import check_incremental_state
    valid, n = check_incremental_state(w["engine"], tag="long replay")
    assert valid == n


def test_incremental_normaliser_under_drift(lib):
    import scenarios
    print(scenarios.incremental_normaliser_under_drift(lib))


def test_batched_mixed_locations_vs_oracle(lib):
    import scenarios
    scenarios.batched_mixed_locations_vs_oracle(lib, cuda=False, N=24)


def test_rolling_quartiles_with_ties_and_small_windows(lib):
    import scenarios
    scenarios.rolling_quartiles_with_ties_and_small_windows(lib)


def test_prefill_and_constant_history_branches(lib):
    import scenarios
    scenarios.prefill_and_constant_history_branches(lib)


def test_state_blob_roundtrip(lib):
    import scenarios
    scenarios.state_blob_roundtrip(lib)


def test_generated_resets_cover_the_start_range(lib):
    import scenarios
    scenarios.device_generated_resets_match_host_statement(lib)
