"""Import shim (TEST INFRASTRUCTURE ONLY): the minimum of `gymnasium` the live
reference needs to import and run in this container.
The reference only uses Env/spaces as containers
(envs/carbon_ls.py:40-45, utils/make_envs_pyenv.py:114-132, envs/bat_env_fwd_view.py:23-27)."""
from . import spaces  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass

    def reset(self, *, seed=None, options=None):
        return None

    def close(self):
        pass
