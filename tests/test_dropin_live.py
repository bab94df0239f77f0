"""The drop-in boundary against the LIVE reference driver (build container only: needs /root/reference): with
indoor-nerf_b200/dropin ahead of PocketNeRF/ on sys.path, the UNMODIFIED run_nerf.py must import, resolve its flat
imports to the B200 modules, and its own create_nerf must build our objects (this is the call that raises TypeError
with the reference's own NeRFSmall); pn.patch must swap in the seven renderer functions."""
import os
import subprocess
import sys
import textwrap

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.live_reference
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import sys, types
    from unittest import mock
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "seaborn", "lpips", "skimage", "skimage.metrics", "pyvista",
                 "configargparse", "kornia", "cv2"):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = mock.MagicMock(name=name)
    sys.path[:0] = [sys.argv[1], sys.argv[2]]           # dropin first, then the reference
    import torch
    import run_nerf                                     # the reference's driver, unmodified
    import indoor_nerf_b200 as pn
    assert run_nerf.__file__.startswith(sys.argv[2]), run_nerf.__file__
    import hash_encoding, run_nerf_helpers, quantization, utils, radam, loss, evaluation_utils
    for mod in (hash_encoding, run_nerf_helpers, quantization, utils, radam, loss, evaluation_utils):
        assert mod.__file__.startswith(sys.argv[1]), mod.__file__
    assert run_nerf.NeRFSmall is pn.NeRFSmall and run_nerf.get_embedder is pn.get_embedder
    assert run_nerf.RAdam is radam.RAdam and run_nerf.ComprehensiveEvaluator is pn.ComprehensiveEvaluator
    assert run_nerf.sample_pdf is pn.sample_pdf and run_nerf.get_rays_np is pn.get_rays_np

    before = {n: getattr(run_nerf, n) for n in ("batchify", "run_network", "batchify_rays", "render", "render_path",
                                                 "raw2outputs", "render_rays")}
    pn.patch(run_nerf)
    for n, f in before.items():
        assert getattr(run_nerf, n) is getattr(pn, n) and getattr(run_nerf, n) is not f, n

    # the reference's own create_nerf, with the hot-path subset of its config (configs/chair.txt defaults)
    a = types.SimpleNamespace(multires=10, multires_views=4, i_embed=1, i_embed_views=2, use_viewdirs=True, N_samples=64,
                              N_importance=128, perturb=1.0, raw_noise_std=0.0, white_bkgd=True, lindisp=False, no_ndc=False,
                              dataset_type="blender", netchunk=65536, lrate=0.01, log2_hashmap_size=8, finest_res=512,
                              use_quantization=True, quantization_bits=8, predict_normals=True, basedir="/tmp/__none__",
                              expname="none", ft_path=None, no_reload=True, netdepth=8, netwidth=256, netdepth_fine=8,
                              netwidth_fine=256,
                              bounding_box=(torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3)))
    import os
    os.makedirs("/tmp/__none__/none", exist_ok=True)
    kw_train, kw_test, start, grad_vars, optimizer = run_nerf.create_nerf(a)
    assert isinstance(kw_train["embed_fn"], pn.HashEmbedder) and isinstance(kw_train["network_fine"], pn.NeRFSmall)
    assert kw_train["network_fine"].predict_normals and start == 0
    assert len(optimizer.param_groups) == 2 and optimizer.param_groups[1]["eps"] == 1e-15
    assert len(optimizer.param_groups[1]["params"]) == 16 + 48            # tables + table-quantiser scalars
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    # network_query_fn resolves run_network in run_nerf's globals at call time: the patched one
    assert kw_train["network_query_fn"].__globals__["run_network"] is pn.run_network
    print("DROPIN-OK")
''')


def test_reference_driver_runs_on_the_dropin_modules():
    if not ref_shim.available():
        pytest.skip("reference not present")
    dropin = os.path.join(ROOT, "indoor-nerf_b200", "dropin")
    env = dict(os.environ, PYTHONPATH="")
    r = subprocess.run([sys.executable, "-c", SCRIPT, dropin, ref_shim.REF_ROOT], capture_output=True, text=True, env=env,
                       cwd="/tmp", timeout=600)
    assert r.returncode == 0 and "DROPIN-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
