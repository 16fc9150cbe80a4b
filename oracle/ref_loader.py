"""Load the reference's rla/srht.py by FILE PATH -- TEST INFRASTRUCTURE ONLY.

`import rla.srht` would execute rla/__init__.py, which imports pyMOR (absent),
so the module is loaded directly from its file.  Only
usable in the build container: /root/reference does not exist on the GPU box,
and nothing that runs there (gpu tests, smoke, bench) may call this.
"""
import importlib.util
import os

REFERENCE_SRHT = "/root/reference/rla/srht.py"


def reference_available():
    return os.path.exists(REFERENCE_SRHT)


def load_reference_srht():
    spec = importlib.util.spec_from_file_location("_rla4mor_reference_srht", REFERENCE_SRHT)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)      # ~9 s: numba compiles the four FWHT variants
    return mod
