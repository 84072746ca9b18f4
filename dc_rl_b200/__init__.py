"""dc_rl_b200: B200-native vectorised SustainDC per-timestep simulation (`sustaindc_env.step`).

Submodules are imported lazily so that host-only helpers (config, sizing, traces) work without torch or the
CUDA library.
"""
__version__ = "0.1.0"

_LAZY = {
    "CudaShareVecEnv": "vec_env",
    "InfoBatch": "vec_env",
    "SustainDC": "sustaindc_env",
    "HARLSustainDCEnv": "harl_env",
    "make_train_env": "harl_env",
    "make_eval_env": "harl_env",
    "SustainDCPettingZooEnv": "ptzoo_env",
    "SustainDCLogger": "logger",
}


def __getattr__(name):
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(__name__ + "." + _LAZY[name]), name)
    raise AttributeError(name)
