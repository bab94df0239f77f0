"""Shadow of the reference's `evaluation_utils.py`: put this directory ahead of PocketNeRF/ on sys.path and
`from evaluation_utils import ComprehensiveEvaluator` resolves to the B200 implementation (PSNR / SSIM reduced
on the GPU; the matplotlib figure helpers are not provided)."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))))
from indoor_nerf_b200.evaluation_utils import *  # noqa: F401,F403,E402
from indoor_nerf_b200 import evaluation_utils as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
