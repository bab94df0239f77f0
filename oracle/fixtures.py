"""TEST INFRASTRUCTURE — deterministic input generators shared by ``oracle/make_golden.py``
(which feeds them to the live reference) and the tests (which feed them to the oracle and to
the CUDA path).  No reference code is involved here."""
import numpy as np
import torch


def synthetic_tables(n_levels, log2T, amp=1e-4, salt=0):
    """Deterministic pseudo-random tables in (-amp, amp): a 32-bit multiplicative hash of
    (level, index, feature), evaluated in numpy uint64 so every platform agrees bit for bit."""
    T = 1 << log2T
    i = np.arange(T, dtype=np.uint64)[None, :, None]
    f = np.arange(2, dtype=np.uint64)[None, None, :]
    l = np.arange(n_levels, dtype=np.uint64)[:, None, None]
    h = (i * np.uint64(2654435761) + f * np.uint64(40503) + (l + np.uint64(salt)) * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    u = h.astype(np.float64) / 4294967296.0
    return ((u - 0.5) * 2.0 * amp).astype(np.float32)          # [L,T,2]


def synthetic_points(n, box_min, box_max, seed, outside_frac=0.1):
    rng = np.random.RandomState(seed)
    lo, hi = np.asarray(box_min, np.float32), np.asarray(box_max, np.float32)
    x = (lo + (hi - lo) * rng.rand(n, 3)).astype(np.float32)
    k = int(n * outside_frac)
    if k:
        x[:k] += ((rng.rand(k, 3) - 0.5) * 2.0 * (hi - lo) * (rng.rand(k, 3) < 0.4)).astype(np.float32)
    # exact faces / corners
    x[-1] = hi
    x[-2] = lo
    x[-3] = (lo + hi) / 2
    return x


def mlp_weights(seed, normals=False):
    g = torch.Generator().manual_seed(seed)
    def lin(o, i):
        bound = 1.0 / np.sqrt(i)
        return ((torch.rand(o, i, generator=g) * 2 - 1) * bound)
    w = dict(s0=lin(64, 32), s1=lin(16, 64), c0=lin(64, 31), c1=lin(64, 64), c2=lin(3, 64))
    if normals:
        w.update(n0w=lin(32, 15), n0b=(torch.rand(32, generator=g) - 0.5) * 0.2,
                 n2w=lin(3, 32), n2b=(torch.rand(3, generator=g) - 0.5) * 0.2)
    return w
