"""Import shim (TEST INFRASTRUCTURE ONLY) for psychrolib==2.5.0 (requirements.txt:53), which is
neither installed here nor vendored under /root/reference. utils/managers.py:10,530 needs
`SI`, `SetUnitSystem`, `GetTWetBulbFromRelHum`. The arithmetic is the ASHRAE-2017 restatement that the
product uses for EPW ingest (dc_rl_b200/psychro.py). PARITY UNPINNED for the wet-bulb trace: no
reference test or installed library pins it."""
import importlib

_psy = importlib.import_module("dc_rl_b200.psychro")

SI = "SI"


def SetUnitSystem(units):
    assert units == SI


GetTWetBulbFromRelHum = _psy.wet_bulb_from_rel_hum
