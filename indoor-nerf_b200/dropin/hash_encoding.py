"""Shadow of the reference's `hash_encoding.py`: put this directory ahead of PocketNeRF/ on sys.path and the
reference's flat `import hash_encoding` / `from hash_encoding import ...` resolve to the B200 implementation."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from indoor_nerf_b200.hash_encoding import *  # noqa: F401,F403,E402
from indoor_nerf_b200 import hash_encoding as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
