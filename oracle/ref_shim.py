"""TEST INFRASTRUCTURE — loader for the *unmodified* PocketNeRF reference.

Imports the reference from ``/root/reference/PocketNeRF`` (the build container) or, where that does not exist
(the GPU box), from the verbatim copy ``oracle/_ref/PocketNeRF`` that ``oracle/build_ref.py`` makes and that ships
with a gpurun snapshot (git-ignored, never in history).  It is used by ``oracle/make_golden.py`` to produce the
committed golden vectors under ``tests/golden/``, by the live cross-check tests and by ``oracle/ref_train_step.py``
(the reference arm of ``bench.py``).  Nothing in the product package imports this file.

The reference cannot be imported as-is: it needs kornia, creates a
CUDA tensor at import (utils.py:9-10), imports imageio/matplotlib/configargparse/
lpips/..., and ``NeRFSmall.__init__`` reads ``self.predict_normals`` before anything
sets it (run_nerf_helpers.py:258).  The shim supplies stand-ins for the absent
third-party modules and nothing else; every arithmetic line that runs is the
reference's own.
"""
import importlib
import os
import sys
import types
from unittest import mock

_HERE = os.path.dirname(os.path.abspath(__file__))


def _find_root():
    env = os.environ.get("POCKETNERF_REFERENCE")
    if env:
        return env
    for cand in ("/root/reference/PocketNeRF", os.path.join(_HERE, "_ref", "PocketNeRF")):
        if os.path.isfile(os.path.join(cand, "hash_encoding.py")):
            return cand
    return "/root/reference/PocketNeRF"


REF_ROOT = _find_root()


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "hash_encoding.py"))


def _install_stubs():
    import torch

    if "kornia" not in sys.modules:
        kornia = types.ModuleType("kornia")

        def create_meshgrid(height, width, normalized_coordinates=False, device=None, dtype=torch.float32):
            xs = torch.linspace(0, width - 1, width, dtype=dtype)
            ys = torch.linspace(0, height - 1, height, dtype=dtype)
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            return torch.stack([gx, gy], -1)[None]

        kornia.create_meshgrid = create_meshgrid
        sys.modules["kornia"] = kornia
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "seaborn", "lpips", "skimage",
                 "skimage.metrics", "pyvista", "configargparse"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = mock.MagicMock(name=name)


_loaded = {}


def load():
    """Return a namespace with the reference modules (hash_encoding, utils, run_nerf_helpers,
    quantization, run_nerf, loss, radam).  CPU-only hosts get the single ``device='cuda'``
    keyword at utils.py:9-10 neutralised while ``utils`` is imported."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference not present at %s" % REF_ROOT)
    import torch

    _install_stubs()
    # the reference uses flat imports; shadow nothing of ours: import under a private path entry
    sys.path.insert(0, REF_ROOT)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k in ("utils", "hash_encoding", "quantization", "run_nerf_helpers", "run_nerf",
                      "ray_utils", "loss", "radam", "optimizer", "structural_priors")}
    try:
        if not torch.cuda.is_available():
            real_tensor = torch.tensor

            def cpu_tensor(*a, **kw):
                if kw.get("device") == "cuda":
                    kw.pop("device")
                return real_tensor(*a, **kw)

            with mock.patch.object(torch, "tensor", cpu_tensor):
                utils = importlib.import_module("utils")
        else:
            utils = importlib.import_module("utils")
        for name in ("quantization", "hash_encoding", "run_nerf_helpers", "loss", "radam", "run_nerf"):
            _loaded[name] = importlib.import_module(name)
        _loaded["utils"] = utils
        _loaded["ray_utils"] = sys.modules["ray_utils"]
    finally:
        sys.path.remove(REF_ROOT)
        # keep the reference modules reachable only through the namespace we return
        for k in ("utils", "hash_encoding", "quantization", "run_nerf_helpers", "run_nerf",
                  "ray_utils", "loss", "radam", "optimizer", "structural_priors", "metric_logger",
                  "evaluation_utils", "load_llff", "load_blender", "load_deepvoxels",
                  "load_scannet", "load_LINEMOD"):
            sys.modules.pop(k, None)
        sys.modules.update(saved)
    return types.SimpleNamespace(**_loaded)


def make_nerf_small(ref, predict_normals=False, **kw):
    """Construct the reference NeRFSmall with the create_nerf shapes (run_nerf.py:240-247);
    ``predict_normals`` has to be a class attribute because of run_nerf_helpers.py:258."""
    cls = ref.run_nerf_helpers.NeRFSmall
    cls.predict_normals = bool(predict_normals)
    args = dict(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3,
                hidden_dim_color=64, input_ch=32, input_ch_views=16)
    args.update(kw)
    net = cls(**args)
    net.predict_normals = bool(predict_normals)
    return net
