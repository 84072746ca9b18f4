"""Import shim (TEST INFRASTRUCTURE ONLY): sustaindc_env.py:22-26 imports matplotlib but never calls it."""


def use(*a, **k):
    pass
