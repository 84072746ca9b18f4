"""Shared helpers for the test-suite (oracle side). TEST INFRASTRUCTURE."""
import functools
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@functools.lru_cache(maxsize=None)
def load_traj(name):
    z = np.load(os.path.join(GOLDEN, f"traj_{name}.npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


@functools.lru_cache(maxsize=None)
def oracle_traces(loc, timezone_shift=0):
    import sdc_oracle
    from dc_rl_b200 import psychro
    z = np.load(os.path.join(GOLDEN, f"loc_{loc}.npz"), allow_pickle=False)
    return sdc_oracle.Traces.from_golden(z, psychro.wet_bulb_from_rel_hum, timezone_shift)


def reward_methods_of(cfg):
    """(ls, dc, bat) reward method names of a trajectory config (defaults as sustaindc_env.py:63-65)."""
    return tuple(cfg.get(k, "default_%s" % k) for k in ("ls_reward", "dc_reward", "bat_reward"))


@functools.lru_cache(maxsize=None)
def dc_configs():
    """utils/dc_config_dc{1,2,3}.json of the reference (fixture minted by oracle/make_golden_r2.py)."""
    with open(os.path.join(GOLDEN, "dc_configs.json")) as f:
        return json.load(f)


@functools.lru_cache(maxsize=None)
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def traj_cfg(g):
    return json.loads(str(g["cfg_json"][0]))


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))


def check_incremental_state(eng, envs=None, tag=""):
    """The incremental reward-normaliser state of every env must describe its window EXACTLY: quartile brackets ==
    the corresponding slice of the fully sorted window, tail bands == the multisets between the inner and outer thresholds, far-tail aggregates == direct sums, window
    moments == the direct sums (fp64 rounding only).  Returns (#envs with valid tail sets, #envs checked)."""
    from dc_rl_b200 import _lib
    n = eng.n_envs
    L = _lib.LIST_CAP
    q_a, q_m, ql = eng.read_state("q_a").reshape(n, 2), eng.read_state("q_m").reshape(n, 2), eng.read_state("qlist").reshape(n, 2, L)
    hl = eng.read_state("hist_len").reshape(n)
    hist = eng.read_state("hist").reshape(n, -1)
    tn, thr, tails = eng.read_state("tail_n").reshape(n, 2), eng.read_state("tail_thr").reshape(n, 4), eng.read_state("tails")
    an, ag = eng.read_state("agg_n").reshape(n, 2), eng.read_state("agg_s").reshape(n, 2, 2)
    nb, bs = eng.read_state("tail_nb").reshape(n, 2), eng.read_state("tail_bs").reshape(n, 2, 2)
    s1, s2, c0 = (eng.read_state(k).reshape(n) for k in ("mom_s1", "mom_s2", "mom_c0"))
    valid = 0
    for e in (range(n) if envs is None else envs):
        w = hist[e, :hl[e]]
        srt = np.sort(w)
        for j in range(2):
            a, m = int(q_a[e, j]), int(q_m[e, j])
            assert np.array_equal(ql[e, j, :m], srt[a:a + m]), (tag, "bracket", e, j, a, m)
            if hl[e] >= 2:
                k = ((1, 3)[j] * (hl[e] - 1)) // 4
                assert a <= k and k + 1 < a + m, (tag, "rank outside bracket", e, j, a, m, k)
        if tn[e, 0] < 0:
            continue
        valid += 1
        lo_set, hi_set = tails[e, 0, :tn[e, 0]], tails[e, 1, :tn[e, 1]]            # stored sorted
        assert np.array_equal(lo_set, srt[(srt < thr[e, 0]) & (srt >= thr[e, 2])]), (tag, "low band", e)
        assert np.array_equal(hi_set, srt[(srt > thr[e, 1]) & (srt <= thr[e, 3])]), (tag, "high band", e)
        # split of each band at the fences of the last step (np.percentile 'linear' in fp64, as the device computes them)
        w64 = w.astype(np.float64)
        q1, q3 = np.percentile(w64, 25), np.percentile(w64, 75)
        f_lo, f_hi = q1 - 1.5 * (q3 - q1), q3 + 1.5 * (q3 - q1)
        for side, beyond in enumerate((lo_set[lo_set.astype(np.float64) < f_lo], hi_set[hi_set.astype(np.float64) > f_hi])):
            yb = beyond.astype(np.float64) - c0[e]
            assert nb[e, side] == len(beyond), (tag, "band split", e, side, nb[e, side], len(beyond))
            assert abs(yb.sum() - bs[e, side, 0]) <= 1e-9 * max(1.0, float(np.abs(yb).sum())), (tag, "band s1", e, side)
            assert abs((yb * yb).sum() - bs[e, side, 1]) <= 1e-9 * max(1.0, float((yb * yb).sum())), (tag, "band s2", e, side)
        for side, far in enumerate((srt[srt < thr[e, 2]], srt[srt > thr[e, 3]])):
            yf = far.astype(np.float64) - c0[e]
            assert an[e, side] == len(far), (tag, "far count", e, side)
            assert abs(yf.sum() - ag[e, side, 0]) <= 1e-9 * max(1.0, float(np.abs(yf).sum())), (tag, "far s1", e, side)
            assert abs((yf * yf).sum() - ag[e, side, 1]) <= 1e-9 * max(1.0, float((yf * yf).sum())), (tag, "far s2", e, side)
        y = w.astype(np.float64) - c0[e]
        scale = max(1.0, float(np.sum(np.abs(y))))
        assert abs(y.sum() - s1[e]) <= 1e-9 * scale, (tag, "s1", e, y.sum(), s1[e])
        assert abs((y * y).sum() - s2[e]) <= 1e-9 * max(1.0, float((y * y).sum())), (tag, "s2", e)
    return valid, (n if envs is None else len(envs))
