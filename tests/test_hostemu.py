"""CPU check of the CUDA sources' bit-exact arithmetic: the host/device functions in
indoor-nerf_b200/csrc/*_core.cuh, compiled for the host by tests/hostemu, against the golden vectors
from the live reference.  This is what lets index / interpolation / sampling bugs surface without a GPU;
the kernels themselves are checked by the -m gpu tests."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import hashnerf_oracle as O
from oracle.fixtures import synthetic_tables

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "hostemu")


class Grid(ctypes.Structure):
    _fields_ = [("box_min", ctypes.c_float * 3), ("box_max", ctypes.c_float * 3),
                ("resolution", ctypes.c_float * 16), ("n_levels", ctypes.c_int32),
                ("log2_hashmap_size", ctypes.c_int32)]


@pytest.fixture(scope="module")
def emu():
    from tests import emu_ops
    so = emu_ops.build()
    return ctypes.CDLL(so)


def fptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def make_grid(bmin, bmax, finest, log2T):
    g = Grid()
    g.box_min[:] = list(bmin)
    g.box_max[:] = list(bmax)
    g.resolution[:] = [float(r) for r in O.level_resolutions(16, finest)]
    g.n_levels, g.log2_hashmap_size = 16, log2T
    return g


def table_ptrs(tables):
    arr = (ctypes.c_void_p * 16)(*[t.ctypes.data for t in tables])
    return arr


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_hash_encode(emu, golden, tag):
    g = golden("hash_embed_" + tag)
    log2T = int(g["log2T"])
    tables = [np.ascontiguousarray(t) for t in synthetic_tables(16, log2T, salt=int(g["salt"]))]
    x = np.ascontiguousarray(g["x"])
    P = x.shape[0]
    grid = make_grid(g["box_min"], g["box_max"], int(g["finest"]), log2T)
    feat = np.zeros((P, 32), np.float32)
    keep = np.zeros(P, np.uint8)
    idx = np.zeros((P, 16, 8), np.int32)
    emu.emu_hash_encode(ctypes.byref(grid), table_ptrs(tables), None, fptr(x), ctypes.c_int64(P), fptr(feat),
                        fptr(keep), fptr(idx))
    assert (idx == g["idx"]).all()
    assert (keep.astype(bool) == g["keep"]).all()
    assert (feat == g["feat"]).all()
    # backward weights
    dt = [np.zeros_like(t) for t in tables]
    dfeat = np.ascontiguousarray(g["dfeat"])
    emu.emu_hash_bwd(ctypes.byref(grid), table_ptrs(dt), fptr(x), fptr(dfeat), ctypes.c_int64(P))
    np.testing.assert_allclose([np.abs(d).sum() for d in dt], g["grad_abs_sum"], rtol=1e-5)
    for l in (0, 7, 15):
        np.testing.assert_allclose(dt[l][g["grad_rows_%d" % l]], g["grad_vals_%d" % l], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_hash_encode_quant(emu, golden, mode):
    g = golden("hash_embed_quant_" + mode)
    tables = [np.ascontiguousarray(t) for t in synthetic_tables(16, 15, salt=5)]
    x = np.ascontiguousarray(g["x"])
    P = x.shape[0]
    grid = make_grid(g["box_min"], g["box_max"], 512, 15)
    q = np.zeros((16, 8), np.float32)
    for l in range(16):
        st = torch.from_numpy(g["q%d" % l])
        scale, zp, qmin, qmax = O.lbq_scalars(st[0], st[1], st[2], False, training=(mode == "train"))
        q[l] = [float(scale), float(scale + 1e-8), float(zp), qmin, qmax, 1.0, float(mode == "train"), 0]
    feat = np.zeros((P, 32), np.float32)
    emu.emu_hash_encode(ctypes.byref(grid), table_ptrs(tables), fptr(q), fptr(x), ctypes.c_int64(P), fptr(feat),
                        None, None)
    assert (feat == g["feat"]).all()


def test_sh4(emu, golden):
    g = golden("sh4")
    d = np.ascontiguousarray(g["dirs"])
    out = np.zeros((d.shape[0], 16), np.float32)
    emu.emu_sh4(fptr(d), ctypes.c_int64(d.shape[0]), fptr(out))
    assert (out == g["out"]).all()


def test_sample_from_cdf(emu, golden):
    g = golden("sample_pdf")
    bins, w = torch.from_numpy(g["bins"]), torch.from_numpy(g["weights"])
    for u, key in ((torch.from_numpy(g["u"]), "rnd"),
                   (torch.linspace(0., 1., 128).expand(bins.shape[0], 128).contiguous(), "det")):
        _, inds, cdf = O.sample_pdf(bins, w, 128, u=u, return_inds=True)
        cdf_n, bins_n, u_n = (np.ascontiguousarray(t.numpy()) for t in (cdf, bins, u))
        N = bins.shape[0]
        samples = np.zeros((N, 128), np.float32)
        ind_out = np.zeros((N, 128), np.int32)
        emu.emu_sample_from_cdf(fptr(cdf_n), fptr(bins_n), fptr(u_n), ctypes.c_int64(128), ctypes.c_int64(N), 63,
                                128, fptr(samples), fptr(ind_out))
        assert (ind_out == inds.numpy()).all()
        assert (samples == g[key]).all()


def test_rays_points_z(emu, golden):
    g = golden("rays")
    H, W = int(g["H"]), int(g["W"])
    K = np.ascontiguousarray(g["K"].astype(np.float32).reshape(-1))
    c2w = np.ascontiguousarray(g["c2w"].reshape(-1))
    ro, rd = np.zeros((H, W, 3), np.float32), np.zeros((H, W, 3), np.float32)
    emu.emu_gen_rays(H, W, fptr(K), fptr(c2w), fptr(ro), fptr(rd))
    assert (rd == g["rays_d"]).all() and (ro == g["rays_o"]).all()

    # NDC
    o_in = np.ascontiguousarray((g["rays_o"] - np.array([0, 0, 10.0], np.float32)).reshape(-1, 3))
    d_in = np.ascontiguousarray(g["rays_d"].reshape(-1, 3))
    oo, od = np.zeros_like(o_in), np.zeros_like(d_in)
    emu.emu_ndc_rays.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_void_p,
                                 ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p]
    emu.emu_ndc_rays(H, W, float(g["K"][0][0]), 1.0, fptr(o_in), fptr(d_in), o_in.shape[0], fptr(oo), fptr(od))
    assert (oo.reshape(H, W, 3) == g["ndc_o"]).all() and (od.reshape(H, W, 3) == g["ndc_d"]).all()

    r = golden("render_rays_blender")
    rays = r["rays"]
    N = rays.shape[0]
    t_vals = torch.linspace(0., 1., 64).numpy()
    for lindisp in (0, 1):
        for t_rand in (None, np.ascontiguousarray(r["t_rand"])):
            z = np.zeros((N, 64), np.float32)
            emu.emu_coarse_z(fptr(np.ascontiguousarray(rays[:, 6])), fptr(np.ascontiguousarray(rays[:, 7])),
                             fptr(t_vals), fptr(t_rand) if t_rand is not None else None, ctypes.c_int64(N), 64,
                             lindisp, fptr(z))
            want = O.coarse_z_vals(torch.from_numpy(rays[:, 6:7]), torch.from_numpy(rays[:, 7:8]), 64, bool(lindisp),
                                   torch.from_numpy(t_rand) if t_rand is not None else None)
            assert (z == want.numpy()).all()
    # points of the fine pass: pts = o + d * z_sorted
    pts = r["pts"]
    o, d = np.ascontiguousarray(rays[:, 0:3]), np.ascontiguousarray(rays[:, 3:6])
    # recover z from the golden points is not possible exactly; use the oracle's z
    from tests.test_oracle_golden import _render_case
    ret, _, _ = _render_case(r)
    zf = np.ascontiguousarray(ret["z_vals"].detach().numpy())
    out = np.zeros_like(pts)
    emu.emu_make_points(fptr(o), fptr(d), fptr(zf), ctypes.c_int64(N), zf.shape[1], fptr(out))
    assert (out == pts).all()


@pytest.mark.parametrize("finest,box", [(512, ((-1.5, -1.5, -1.5), (1.5, 1.5, 1.5))),
                                        (1024, ((-1.21, -0.93, -2.05), (1.37, 1.11, 0.49))),
                                        (512, ((-4.7, -3.1, -0.2), (5.9, 7.3, 3.3)))])
def test_fast_cell_index_equals_exact(emu, finest, box):
    """The bf16 paths take the voxel index from a reciprocal multiply with an exact-division fallback
    (hash_core.cuh point_cell<false>): it must equal the reference's floor((x - min) / g) (utils.py:108-110)
    everywhere — random points, points outside the box, and points ON voxel boundaries at every level, where
    the quotient sits within an ulp of an integer."""
    bmin, bmax = np.array(box[0], np.float32), np.array(box[1], np.float32)
    g = make_grid(bmin, bmax, finest, 19)
    rng = np.random.default_rng(7)
    pts = [rng.uniform(bmin - 0.05, bmax + 0.05, size=(200000, 3)).astype(np.float32)]
    res = [float(r) for r in O.level_resolutions(16, finest)]
    for r in res:                                   # knife edges: x = min + i * g (fp32), one ulp below, one ulp above
        gs = ((bmax - bmin) / np.float32(r)).astype(np.float32)
        i = rng.integers(0, int(r) + 1, size=(4000, 3)).astype(np.float32)
        edge = (i * gs + bmin).astype(np.float32)
        pts += [edge, np.nextafter(edge, np.float32(-np.inf)), np.nextafter(edge, np.float32(np.inf))]
        q = rng.integers(0, int(r) + 1, size=(4000, 3)).astype(np.float32)      # and via the quotient: x = min + q / (1/g)
        pts.append((q / (np.float32(1.0) / gs) + bmin).astype(np.float32))
    x = np.ascontiguousarray(np.concatenate(pts, 0))
    nfb = ctypes.c_int64(0)
    emu.emu_cell_index_mismatches.restype = ctypes.c_int64
    bad = emu.emu_cell_index_mismatches(ctypes.byref(g), fptr(x), ctypes.c_int64(x.shape[0]), ctypes.byref(nfb))
    assert bad == 0, "%d voxel indices differ between the fast and the exact form" % bad
    frac = nfb.value / (x.shape[0] * 16 * 3)
    assert 0 < frac < 0.25, "fallback rate %.4f (expected: rare on random points, common on the knife edges)" % frac


@pytest.mark.parametrize("train_form", [0, 1])
def test_division_free_fake_quant_equals_exact(emu, train_form):
    """The tensor-core MLP quantises its hidden activations with fake_quant_rcp (reciprocal multiply, IEEE division only
    near a rounding tie): it must return fake_quant's bits (quantization.py:177-187) everywhere — random activations,
    values ON the rounding ties of x / denom + zp (and one ulp either side), values beyond the clamp range, inf / NaN."""
    rng = np.random.default_rng(21)
    emu.emu_fake_quant_mismatches.restype = ctypes.c_int64
    f32 = ctypes.c_float
    total_fb, total_n = 0, 0
    for bits, vmax, zp_frac in ((8, 3.7, 0.0), (8, 0.0123, 0.37), (4, 11.0, 0.5), (12, 0.91, 0.123), (6, 250.0, 0.0),
                                (10, 1e-3, 0.77)):
        qmin, qmax = 0.0, float(2 ** bits - 1)
        denom = np.float32(vmax / qmax)                       # the divisor; scale deliberately a different rounding of it
        scale = np.float32(np.float32(vmax) / np.float32(qmax))
        zp = np.float32(np.round(zp_frac * qmax) if zp_frac in (0.0, 0.5) else zp_frac * qmax)
        xs = [rng.uniform(-0.2 * vmax, 1.3 * vmax, 300000).astype(np.float32),
              (rng.standard_normal(100000) * vmax).astype(np.float32)]
        k = rng.integers(-2, int(qmax) + 3, 20000).astype(np.float32)
        tie = ((k + np.float32(0.5) - zp) * denom).astype(np.float32)       # x with x / denom + zp ~ k + 0.5
        for d in range(-3, 4):
            t = tie.copy()
            for _ in range(abs(d)):
                t = np.nextafter(t, np.float32(np.inf if d > 0 else -np.inf))
            xs.append(t)
        xs.append(np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e30, -1e30, 1e-40], np.float32))
        x = np.ascontiguousarray(np.concatenate(xs))
        nfb = ctypes.c_int64(0)
        bad = emu.emu_fake_quant_mismatches(fptr(x), ctypes.c_int64(x.shape[0]), f32(scale), f32(denom), f32(zp), f32(qmin),
                                            f32(qmax), ctypes.c_int(train_form), ctypes.byref(nfb))
        assert bad == 0, "%d values differ between the division-free and the exact fake-quant (bits %d)" % (bad, bits)
        total_fb += nfb.value
        total_n += x.shape[0]
    frac = total_fb / total_n
    assert 0 < frac < 0.5, "fallback rate %.4f (expected: rare on random values, the rule on the ties)" % frac
