"""The oracle of the data formats either side of the path (oracle/dataio_oracle.py) against the golden vectors
the live reference produced (oracle/make_golden_io.py), and the host build of the kernels' per-element
arithmetic (csrc/io_core.cuh via tests/hostemu) against both.  No GPU needed."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import dataio_oracle as D

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "hostemu")
QROW = 8


@pytest.fixture(scope="module")
def emu():
    from tests import emu_ops
    so = emu_ops.build()
    lib = ctypes.CDLL(so)
    lib.emu_ssim_sum.restype = ctypes.c_double
    return lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def qrow_of(scale, zp, qmin, qmax):
    scale = np.float32(scale)
    return np.array([scale, np.float32(scale + np.float32(1e-8)), zp, qmin, qmax, 1, 0, 0], np.float32)


# ---- ray bank -------------------------------------------------------------------------------------------------------
def test_oracle_ray_bank_matches_reference(golden):
    g = golden("ray_bank")
    H, W, K = int(g["H"]), int(g["W"]), g["K"]
    bank = D.rays_rgb_bank(H, W, K, g["poses"], g["images"], list(g["i_train"]))
    assert bank.dtype == np.float32 and (bank == g["unshuffled"]).all()
    np.random.seed(0)
    order = D.shuffle_order(bank.shape[0])
    assert (order == g["order"]).all() and (bank[order] == g["shuffled"]).all()
    rays, tgt = D.batch_from_bank(g["shuffled"], 32, 16)
    assert rays.shape == (2, 16, 3) and (tgt == g["shuffled"][32:48, 2]).all()


@pytest.mark.parametrize("f64", [1, 0])
def test_hostemu_ray_bank(emu, golden, f64):
    g = golden("ray_bank")
    H, W, K = int(g["H"]), int(g["W"]), np.ascontiguousarray(g["K"], np.float64)
    poses, images = np.ascontiguousarray(g["poses"]), np.ascontiguousarray(g["images"])
    idx = np.ascontiguousarray(g["i_train"], np.int32)
    ids = np.ascontiguousarray(g["order"], np.int64)
    B = ids.size
    rays, tgt = np.zeros((2, B, 3), np.float32), np.zeros((B, 3), np.float32)
    emu.emu_ray_bank(ptr(ids), ctypes.c_int64(B), H, W, ptr(K), ptr(poses), ctypes.c_int64(16), ptr(idx), ptr(images),
                     f64, ptr(rays), ptr(tgt))
    assert (tgt == g["shuffled"][:, 2]).all() and (rays[0] == g["shuffled"][:, 0]).all()
    if f64:
        assert (rays[1] == g["shuffled"][:, 1]).all()                       # bit-exact get_rays_np + astype
    else:
        # the float32 arithmetic of get_rays: same rays to float32 rounding, and bit-equal to the oracle's get_rays
        assert np.abs(rays[1] - g["shuffled"][:, 1]).max() < 1e-6
        import torch
        from oracle import hashnerf_oracle as O
        for n, img in enumerate(idx):
            ro, rd = O.get_rays(H, W, K, torch.from_numpy(poses[img, :3, :4]))
            sel = np.nonzero(ids // (H * W) == n)[0]
            pix = ids[sel] % (H * W)
            assert (rays[1][sel] == rd.reshape(-1, 3).numpy()[pix]).all()


# ---- evaluation -----------------------------------------------------------------------------------------------------
def test_oracle_psnr_and_to8b(emu, golden):
    g = golden("eval_psnr")
    assert D.psnr_render_path(g["rgb"], g["gt"]) == pytest.approx(float(g["p_render_path"]), rel=1e-7)
    assert float(g["p_eval"]) == pytest.approx(float(g["p_render_path"]), rel=1e-5)
    assert (D.to8b(g["rgb"]) == g["rgb8"]).all()
    x = np.ascontiguousarray(np.concatenate([g["rgb"].reshape(-1), np.float32([-0.5, 0, 1, 1.5, 0.999999, 1 / 255, 0.00392])]))
    out = np.zeros(x.size, np.uint8)
    emu.emu_to8b(ptr(x), ctypes.c_int64(x.size), ptr(out))
    assert (out == D.to8b(x)).all()


def test_oracle_ssim_known_answers(emu):
    rs = np.random.RandomState(0)
    a = rs.rand(12, 11, 3).astype(np.float32)
    b = np.clip(a + rs.randn(12, 11, 3).astype(np.float32) * 0.1, 0, 1).astype(np.float32)
    assert D.ssim(a, a) == pytest.approx(1.0, abs=1e-6)
    brute = D.ssim_bruteforce(a, b)
    assert D.ssim(a, b) == pytest.approx(brute, abs=2e-5)                  # float32 uniform_filter vs float64 windows
    assert D.ssim(a.astype(np.float64), b.astype(np.float64)) == pytest.approx(brute, abs=1e-12)
    s = emu.emu_ssim_sum(ptr(a), ptr(b), 12, 11, 3, ctypes.c_double(1.0))
    assert s / (6 * 5 * 3) == pytest.approx(brute, abs=1e-12)
    const = np.full((9, 9, 1), 0.25, np.float32)
    assert D.ssim_bruteforce(const, const * 2) == pytest.approx((2 * .25 * .5 + 1e-4) / (.25 ** 2 + .5 ** 2 + 1e-4), rel=1e-9)


# ---- A-CAQ export ----------------------------------------------------------------------------------------------------
CASES = [("asym", k, False, "table") for k in range(6)] + [("sym", k, True, "w0") for k in range(3)]


@pytest.mark.parametrize("tag,k,sym,src", CASES)
def test_quant_codes_roundtrip(emu, golden, tag, k, sym, src):
    g = golden("quant_export")
    x = np.ascontiguousarray(g[src])
    pre = "%s_%d_" % (tag, k)
    bits, scale, zp, qmin, qmax = D.lbq_eval_params(g[pre + "bits"], g[pre + "range_scale"], g[pre + "v_max"], sym)
    assert bits == int(round(float(g[pre + "bits"])))
    codes = D.quant_codes(x, scale, zp, qmin, qmax)
    assert codes.min() >= 0 and codes.max() < 2 ** bits
    assert (D.dequant_codes(codes, scale, zp, qmin) == g[pre + "out"]).all()         # oracle == reference, bit for bit
    words = D.pack_bits(codes, bits)
    assert words.size == x.size * bits // 32
    assert (D.unpack_bits(words, bits, x.size) == codes).all()
    # the kernels' arithmetic (host build): same words, same values
    row = qrow_of(scale, zp, qmin, qmax)
    w2 = np.zeros_like(words)
    emu.emu_quant_pack(ptr(x), ctypes.c_int64(x.size), ptr(row), bits, ptr(w2))
    assert (w2 == words).all()
    y = np.zeros_like(x)
    emu.emu_quant_unpack(ptr(w2), ctypes.c_int64(x.size), ptr(row), bits, ptr(y))
    assert (y == g[pre + "out"]).all()


# ---- tables resident as integer codes: the gather must reproduce the reference's eval-mode quantised embedder -------------
class _Grid(ctypes.Structure):
    _fields_ = [("box_min", ctypes.c_float * 3), ("box_max", ctypes.c_float * 3), ("resolution", ctypes.c_float * 16),
                ("n_levels", ctypes.c_int32), ("log2_hashmap_size", ctypes.c_int32)]


class _Packed(ctypes.Structure):
    _fields_ = [("codes", ctypes.c_void_p * 16), ("entry_bytes", ctypes.c_int32 * 16), ("scale", ctypes.c_float * 16),
                ("zero_point", ctypes.c_float * 16), ("qmin", ctypes.c_float * 16)]


@pytest.mark.parametrize("force_bits", [None, 12.0, 20.0])
def test_hostemu_packed_gather_matches_reference_eval(emu, golden, force_bits):
    import torch
    from oracle import hashnerf_oracle as O
    from oracle.fixtures import synthetic_tables
    g = golden("hash_embed_quant_eval")
    tables = [np.ascontiguousarray(t) for t in synthetic_tables(16, 15, salt=5)]
    x = np.ascontiguousarray(g["x"])
    P = x.shape[0]
    grid = _Grid()
    grid.box_min[:] = list(g["box_min"]); grid.box_max[:] = list(g["box_max"])
    grid.resolution[:] = [float(r) for r in O.level_resolutions(16, 512)]
    grid.n_levels, grid.log2_hashmap_size = 16, 15
    pk, keepalive, rows = _Packed(), [], np.zeros((16, 8), np.float32)
    for l in range(16):
        st = g["q%d" % l]
        soft = float(st[0]) if force_bits is None else force_bits
        bits, scale, zp, qmin, qmax = D.lbq_eval_params(soft, st[1], st[2], False)
        rows[l] = qrow_of(scale, zp, qmin, qmax)
        if bits <= 16:
            cb = 1 if bits <= 8 else 2
            codes = np.zeros(tables[l].size, np.uint8 if cb == 1 else np.uint16)
            emu.emu_quant_codes(ptr(tables[l]), ctypes.c_int64(tables[l].size), ptr(rows[l]), cb, ptr(codes))
            assert (codes.astype(np.int64) == D.quant_codes(tables[l].reshape(-1), scale, zp, qmin, qmax)).all()
            # the same codes out of the bit stream
            words = D.pack_bits(codes, bits)
            c2 = np.zeros_like(codes)
            emu.emu_quant_unpack_codes(ptr(words), ctypes.c_int64(codes.size), bits, cb, ptr(c2))
            assert (c2 == codes).all()
            pk.entry_bytes[l] = 2 * cb
        else:
            codes = D.dequant_codes(D.quant_codes(tables[l].reshape(-1), scale, zp, qmin, qmax), scale, zp, qmin)
            pk.entry_bytes[l] = 8
        keepalive.append(codes)
        pk.codes[l] = codes.ctypes.data
        pk.scale[l], pk.zero_point[l], pk.qmin[l] = scale, zp, qmin
    feat = np.zeros((P, 32), np.float32)
    keep = np.zeros(P, np.uint8)
    emu.emu_hash_encode_packed(ctypes.byref(grid), ctypes.byref(pk), ptr(x), ctypes.c_int64(P), ptr(feat), ptr(keep))
    if force_bits is None:
        assert (feat == g["feat"]).all()                    # the live reference's eval-mode output, bit for bit
    # and always: equal to fake-quantising the fp32 entries in the gather (the training-time kernel in eval form)
    rows[:, 5] = 1.0
    feat_fq = np.zeros((P, 32), np.float32)
    tp = (ctypes.c_void_p * 16)(*[t.ctypes.data for t in tables])
    emu.emu_hash_encode(ctypes.byref(grid), tp, ptr(rows), ptr(x), ctypes.c_int64(P), ptr(feat_fq), None, None)
    assert (feat == feat_fq).all()


@pytest.mark.parametrize("bits", list(range(1, 25)))
def test_pack_unpack_every_width(emu, bits):
    """Bit stream round trip at every width the exporter can emit (1..24), asymmetric and symmetric ranges, with the
    extreme codes 0 and 2^bits-1 present: host build of the kernels' arithmetic == numpy oracle == input codes."""
    rs = np.random.RandomState(bits)
    n = 32 * 37
    for sym in (False, True):
        qmin = np.float32(-(2 ** (bits - 1)) if sym else 0)
        qmax = np.float32(2 ** (bits - 1) - 1 if sym else 2 ** bits - 1)
        zp = np.float32(0 if sym else rs.randint(0, 2 ** bits))
        scale = np.float32(rs.uniform(1e-7, 1e-2))
        codes = rs.randint(0, 2 ** bits, n).astype(np.int64)
        codes[:2] = [0, 2 ** bits - 1]
        x = D.dequant_codes(codes, scale, zp, qmin)                       # values that sit exactly on the lattice
        row = qrow_of(scale, zp, qmin, qmax)
        # quantising lattice values may move them by a step (the reference's 1e-8 in the divisor): pack the oracle's codes
        want_codes = D.quant_codes(x, scale, zp, qmin, qmax)
        words = np.zeros(n * bits // 32, np.uint32)
        emu.emu_quant_pack(ptr(x), ctypes.c_int64(n), ptr(row), bits, ptr(words))
        assert (words == D.pack_bits(want_codes, bits)).all()
        assert (D.unpack_bits(words, bits, n) == want_codes).all()
        y = np.zeros(n, np.float32)
        emu.emu_quant_unpack(ptr(words), ctypes.c_int64(n), ptr(row), bits, ptr(y))
        assert (y == D.dequant_codes(want_codes, scale, zp, qmin)).all()
        if bits <= 16:
            cb = 1 if bits <= 8 else 2
            c = np.zeros(n, np.uint8 if cb == 1 else np.uint16)
            emu.emu_quant_unpack_codes(ptr(words), ctypes.c_int64(n), bits, cb, ptr(c))
            assert (c.astype(np.int64) == want_codes).all()
