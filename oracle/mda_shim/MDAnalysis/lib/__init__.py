"""TEST INFRASTRUCTURE ONLY -- see `oracle/mda_shim/MDAnalysis/__init__.py`."""
from . import distances  # noqa: F401
