import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (REPO, os.path.join(REPO, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def cuda_lib_or_skip():
    """The product library bound on a CUDA device, or a skip (CPU runs select `-m "not gpu"`; a GPU test that is selected
    on a machine without a device is skipped, not an error)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dc_rl_b200 import _lib
    return _lib.load()


def lib_params():
    """Parametrisation of tests that run the same body on the serial hostsim build (CPU suite) and on libsdc_b200.so
    (`-m gpu`): the reference-facing Python surface is exercised on the CUDA library itself."""
    return ["hostsim", pytest.param("cuda", marks=pytest.mark.gpu)]


def resolve_lib(kind):
    if kind == "cuda":
        return cuda_lib_or_skip()
    import hostsim_build
    return hostsim_build.load()
