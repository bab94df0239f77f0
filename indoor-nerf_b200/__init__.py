"""indoor-nerf_b200 — B200-native (sm_100a) implementation of PocketNeRF's HashNeRF hot path
(HashEmbedder -> SHEncoder -> NeRFSmall -> raw2outputs -> sample_pdf inside render_rays) behind the
reference's Python API.  Import as ``indoor_nerf_b200`` (the importable alias of this directory)."""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from .ops import get_mlp_mode, set_mlp_mode  # noqa: F401
from .hash_encoding import HashEmbedder, SHEncoder  # noqa: F401
from .quantization import FakeQuantizer, LearnedBitwidthQuantizer, PassthroughQuantizer, calculate_fqr  # noqa: F401
from .run_nerf_helpers import (NeRFSmall, get_embedder, get_rays, get_rays_np, img2mse, mse2psnr, ndc_rays,  # noqa: F401
                               sample_pdf, to8b)
from .render import batchify, batchify_rays, patch, raw2outputs, render, render_path, render_rays, run_network  # noqa: F401
from .ray_bank import RayBank  # noqa: F401
from .evaluation_utils import ComprehensiveEvaluator  # noqa: F401
from . import quant_export  # noqa: F401

__version__ = "0.1.0"
