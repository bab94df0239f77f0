"""Oracle vs the LIVE, unmodified reference on fresh random inputs (build container only: needs /root/reference).
Complements test_oracle_golden.py, whose fixtures are frozen outputs of the same reference."""
import numpy as np
import pytest
import torch

from oracle import hashnerf_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.live_reference


@pytest.fixture(scope="module")
def ref():
    if not ref_shim.available():
        pytest.skip("reference not present")
    return ref_shim.load()


def test_hash_embedder_random(ref):
    torch.manual_seed(11)
    box = (torch.tensor([-2.0, -1.0, -3.0]), torch.tensor([1.5, 2.5, 0.5]))
    for finest in (512, 1024):
        emb = ref.hash_encoding.HashEmbedder(box, log2_hashmap_size=12, finest_resolution=finest).eval()
        with torch.no_grad():
            for e in emb.embeddings:
                e.weight.uniform_(-1, 1)
        x = (torch.rand(500, 3) - 0.5) * 8.0
        feat, keep = emb(x)
        res = O.level_resolutions(16, finest)
        of, ok = O.hash_embed(x, box[0], box[1], [e.weight for e in emb.embeddings], res, 12)
        assert torch.equal(of, feat) and torch.equal(ok, keep)


def test_nerf_small_and_sh_random(ref):
    torch.manual_seed(12)
    for normals in (False, True):
        net = ref_shim.make_nerf_small(ref, predict_normals=normals)
        d = torch.nn.functional.normalize(torch.randn(200, 3), dim=-1)
        sh = ref.hash_encoding.SHEncoder()(d)
        assert torch.equal(O.sh4(d), sh)
        x = torch.cat([torch.randn(200, 32), sh], -1)
        w = dict(s0=net.sigma_net[0].weight, s1=net.sigma_net[1].weight, c0=net.color_net[0].weight,
                 c1=net.color_net[1].weight, c2=net.color_net[2].weight)
        if normals:
            w.update(n0w=net.normal_net[0].weight, n0b=net.normal_net[0].bias, n2w=net.normal_net[2].weight,
                     n2b=net.normal_net[2].bias)
        assert torch.allclose(O.nerf_small(x, w), net(x), rtol=1e-6, atol=1e-7)


def test_raw2outputs_and_sample_pdf_random(ref):
    torch.manual_seed(13)
    RN, H = ref.run_nerf, ref.run_nerf_helpers
    raw = torch.randn(50, 64, 4) * torch.tensor([1.0, 1.0, 1.0, 3.0])
    z = torch.sort(2 + 4 * torch.rand(50, 64), -1)[0]
    d = torch.randn(50, 3)
    for white in (False, True):
        want = RN.raw2outputs(raw, z, d, 0, white)
        got = O.raw2outputs(raw, z, d, None, white)
        for a, b in zip(got, want):
            assert torch.allclose(a, b, rtol=1e-6, atol=1e-7, equal_nan=True)
    bins = .5 * (z[:, 1:] + z[:, :-1])
    wts = torch.rand(50, 62) ** 3
    assert torch.equal(O.sample_pdf(bins, wts, 128, det=True), H.sample_pdf(bins, wts, 128, det=True))
    for lindisp in (False, True):
        near, far = torch.full((50, 1), 2.0), torch.full((50, 1), 6.0)
        t = torch.linspace(0., 1., 64)
        want = (1. / (1. / near * (1. - t) + 1. / far * t)) if lindisp else near * (1. - t) + far * t
        assert torch.equal(O.coarse_z_vals(near, far, 64, lindisp), want.expand(50, 64))
