"""Shadow of the reference's `run_nerf_helpers.py` (star-imported by run_nerf.py:20)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from indoor_nerf_b200 import run_nerf_helpers as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
