"""Shadow of the reference's `ray_utils.py`: put this directory ahead of PocketNeRF/ on sys.path and the
reference's flat `import ray_utils` / `from ray_utils import ...` resolve to the B200 implementation."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from indoor_nerf_b200.ray_utils import *  # noqa: F401,F403,E402
from indoor_nerf_b200 import ray_utils as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
