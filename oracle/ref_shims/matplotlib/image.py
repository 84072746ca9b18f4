"""shim: imported, never called (sustaindc_env.py:25)."""
