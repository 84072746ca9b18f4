"""TEST INFRASTRUCTURE: builds tests/hostsim/_build/libsdc_hostsim.so (serial host build of the device
logic, see tests/hostsim/hostsim.cpp) and binds it with the package's ctypes prototypes."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
SRC = os.path.join(HERE, "hostsim", "hostsim.cpp")
OUT = os.path.join(HERE, "hostsim", "_build", "libsdc_hostsim.so")
DEPS = [SRC] + [os.path.join(REPO, "dc_rl_b200", "csrc", f) for f in ("sdc_core.h", "sdc_api.inc")] + [
    os.path.join(REPO, "include", "sdc_b200.h")]


def build():
    if os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", OUT, SRC])
    return OUT


_lib = None


def load():
    global _lib
    if _lib is None:
        from dc_rl_b200 import _lib as binding
        _lib = binding.load(build())
    return _lib
