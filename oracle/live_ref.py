"""TEST INFRASTRUCTURE ONLY -- harness that imports and runs the LIVE reference (HewlettPackard/dc-rl,
read-only under /root/reference) inside this container.  It exists to (a) pin the restatement in
oracle/sdc_oracle.py and (b) mint the golden vectors under tests/golden/ (see oracle/make_golden.py).
Nothing here may be imported by the product path, by `-m gpu` tests, smoke() or bench.py:
/root/reference does not exist on the GPU box.

Shims: gymnasium, matplotlib, psychrolib (oracle/ref_shims/) and a dummy
`harl.envs.sustaindc.dashboard_v2` module (sustaindc_env.py:32 imports it unconditionally).
Nondeterminism neutralised: rack order is pinned to JSON order by replacing
`as_completed` in utils/dc_config_reader.py:100-105 with the identity.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("SDC_REFERENCE_ROOT", "/root/reference")
_HERE = os.path.dirname(os.path.abspath(__file__))
_REPO = os.path.dirname(_HERE)


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "sustaindc_env.py"))


def import_reference():
    """Returns the live `sustaindc_env` module of the reference, with shims installed."""
    if not available():
        raise RuntimeError("live reference not present at %s" % REFERENCE_ROOT)
    for p in (_REPO, REFERENCE_ROOT, os.path.join(_HERE, "ref_shims")):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    if "harl.envs.sustaindc.dashboard_v2" not in sys.modules:
        # Namespace stubs so that `from harl.envs.sustaindc.dashboard_v2 import Dashboard` resolves
        # without executing harl/envs/__init__.py (absl flags, tensorboardX logger).
        for name in ("harl", "harl.envs", "harl.envs.sustaindc"):
            if name not in sys.modules:
                m = types.ModuleType(name)
                m.__path__ = []
                sys.modules[name] = m
        dash = types.ModuleType("harl.envs.sustaindc.dashboard_v2")
        dash.Dashboard = type("Dashboard", (), {})
        sys.modules["harl.envs.sustaindc.dashboard_v2"] = dash
    import utils.dc_config_reader as reader  # noqa: E402  (reference module)
    reader.as_completed = lambda futures: list(futures)  # pin rack order to JSON order
    import sustaindc_env  # noqa: E402  (reference module)
    return sustaindc_env


def fresh_env(env_config):
    """One live reference env with a private (cleared) reward history
    (utils/reward_creator.py:5 is a module global shared by all envs of a process)."""
    mod = import_reference()
    from utils import reward_creator
    reward_creator.energy_history.clear()
    return mod.SustainDC(env_config)
