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
def oracle_traces(loc):
    import sdc_oracle
    from dc_rl_b200 import psychro
    z = np.load(os.path.join(GOLDEN, f"loc_{loc}.npz"), allow_pickle=False)
    return sdc_oracle.Traces.from_golden(z, psychro.wet_bulb_from_rel_hum)


@functools.lru_cache(maxsize=None)
def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def traj_cfg(g):
    return json.loads(str(g["cfg_json"][0]))


def rel_err(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))
