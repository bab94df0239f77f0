"""TEST INFRASTRUCTURE — CPU restatement (numpy / scipy) of the data formats either side of the hot path
(SURVEY.md §8f-2..4): the use_batching ray bank, render_path's PSNR, ComprehensiveEvaluator's PSNR / SSIM,
to8b, and the integer codes behind LearnedBitwidthQuantizer's eval form.  Only ``tests/``, ``smoke()`` and
``bench.py``'s baseline legs may import this file; the product package never does.

Parity status
  * ray bank, PSNR, to8b, quantiser codes: PINNED — ``oracle/make_golden_io.py`` runs the reference's own
    functions (get_rays_np, LearnedBitwidthQuantizer) in the build container and commits their outputs as
    tests/golden/{ray_bank,quant_export,eval_psnr}.npz; tests/test_oracle_io.py holds this file to them.
  * SSIM: parity UNPINNED.  The reference calls skimage.metrics.structural_similarity
    (evaluation_utils.py:33; scikit-image==0.25.2, requirements.txt:46), which is neither under
    /root/reference nor installed offline.  ``ssim`` below restates its published algorithm
    (skimage/metrics/_structural_similarity.py, Wang et al. 2004: 7x7 uniform window, sample covariance,
    K1=0.01, K2=0.03, mean of the map cropped by (win-1)//2) with scipy.ndimage.uniform_filter, the routine
    scikit-image itself calls; it is checked against a brute-force float64 evaluation and known answers only.

All paths relative to /root/reference/PocketNeRF.
"""
import numpy as np


# ---------------------------------------------------------------------------------------------------
# §8f-3: use_batching ray bank
# ---------------------------------------------------------------------------------------------------
def get_rays_np(H, W, K, c2w):
    """run_nerf_helpers.py:323-330.  With K a float64 ndarray (run_nerf.py:832-836) the pixel grid is promoted
    to float64 by the subtraction; the caller rounds to float32 afterwards (run_nerf.py:905)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    d0 = (i - K[0][2]) / K[0][0]
    d1 = -(j - K[1][2]) / K[1][1]
    dirs = np.stack([d0, d1, -np.ones_like(i)], -1)
    R = c2w[:3, :3]
    rays_d = np.stack([(dirs[..., 0] * R[c, 0] + dirs[..., 1] * R[c, 1]) + dirs[..., 2] * R[c, 2] for c in range(3)], -1)
    rays_o = np.broadcast_to(c2w[:3, -1], rays_d.shape)
    return rays_o, rays_d


def rays_rgb_bank(H, W, K, poses, images, i_train):
    """run_nerf.py:899-905: the UNSHUFFLED [n_train*H*W, 3(ro,rd,rgb), 3] float32 tensor."""
    rays = np.stack([np.stack(get_rays_np(H, W, K, p), 0) for p in poses[:, :3, :4]], 0)     # [N, ro+rd, H, W, 3]
    rays_rgb = np.concatenate([rays, images[:, None]], 1)                                     # [N, 3, H, W, 3]
    rays_rgb = np.transpose(rays_rgb, [0, 2, 3, 1, 4])
    rays_rgb = np.stack([rays_rgb[i] for i in i_train], 0)
    return np.reshape(rays_rgb, [-1, 3, 3]).astype(np.float32)


def shuffle_order(n, rng=np.random):
    """The permutation np.random.shuffle(rays_rgb) applies (run_nerf.py:907): shuffling the rows of an
    [n, ...] array and shuffling arange(n) draw the same Fisher-Yates indices from the same RNG state."""
    order = np.arange(n)
    rng.shuffle(order)
    return order


def batch_from_bank(rays_rgb, i_batch, N_rand):
    """run_nerf.py:962-966 -> batch_rays [2,B,3], target_s [B,3]."""
    batch = np.transpose(rays_rgb[i_batch:i_batch + N_rand], (1, 0, 2))
    return batch[:2], batch[2]


# ---------------------------------------------------------------------------------------------------
# §8f-2: evaluation
# ---------------------------------------------------------------------------------------------------
def to8b(x):
    """run_nerf_helpers.py:13."""
    return (255 * np.clip(x, 0, 1)).astype(np.uint8)


def psnr_render_path(rgb, gt):
    """run_nerf.py:186 (float32 numpy arithmetic)."""
    return -10. * np.log10(np.mean(np.square(rgb - gt)))


def ssim(im1, im2, data_range=1.0, win_size=7, K1=0.01, K2=0.03):
    """structural_similarity(im1, im2, channel_axis=2, data_range=data_range) of scikit-image 0.25.2 with its
    defaults (gaussian_weights=False, use_sample_covariance=True), as evaluation_utils.py:33 calls it.
    float32 inputs stay float32 (skimage's _supported_float_type); the final mean is taken in float64."""
    from scipy.ndimage import uniform_filter
    im1, im2 = np.asarray(im1), np.asarray(im2)
    ft = np.float32 if im1.dtype == np.float32 else np.float64
    vals = []
    for c in range(im1.shape[2]):
        x, y = im1[..., c].astype(ft, copy=False), im2[..., c].astype(ft, copy=False)
        NP = win_size ** 2
        cov_norm = NP / (NP - 1)
        ux, uy = uniform_filter(x, size=win_size), uniform_filter(y, size=win_size)
        uxx, uyy, uxy = uniform_filter(x * x, size=win_size), uniform_filter(y * y, size=win_size), uniform_filter(x * y, size=win_size)
        vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
        C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        pad = (win_size - 1) // 2
        vals.append(S[pad:-pad, pad:-pad].mean(dtype=np.float64))
    return float(np.mean(vals))


def ssim_bruteforce(im1, im2, data_range=1.0, win_size=7, K1=0.01, K2=0.03):
    """The same definition evaluated window by window in float64 (slow; tiny images only)."""
    a, b = np.asarray(im1, np.float64), np.asarray(im2, np.float64)
    H, W, C = a.shape
    NP = win_size ** 2
    cov = NP / (NP - 1)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    tot = 0.0
    for c in range(C):
        for y in range(H - win_size + 1):
            for x in range(W - win_size + 1):
                u, v = a[y:y + win_size, x:x + win_size, c], b[y:y + win_size, x:x + win_size, c]
                ux, uy = u.mean(), v.mean()
                vx, vy, vxy = cov * ((u * u).mean() - ux * ux), cov * ((v * v).mean() - uy * uy), cov * ((u * v).mean() - ux * uy)
                tot += ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux * ux + uy * uy + C1) * (vx + vy + C2))
    return tot / (C * (H - win_size + 1) * (W - win_size + 1))


# ---------------------------------------------------------------------------------------------------
# §8f-4: integer codes of the A-CAQ quantiser (quantization.py:126-187, eval form)
# ---------------------------------------------------------------------------------------------------
def lbq_eval_params(soft_bits, range_scale, v_max, symmetric, min_bits=2.0, max_bits=32.0):
    """(bits, scale, zp, qmin, qmax) as float32 scalars, eval form: B = round(clamp(soft_bits))."""
    import torch
    bw = torch.clamp(torch.tensor(float(soft_bits)), min_bits, max_bits)
    B = int(torch.round(bw).item())
    rs = torch.tensor(float(range_scale))
    if symmetric:
        qmin, qmax = -(2 ** (B - 1)), 2 ** (B - 1) - 1
        scale = rs / (2 ** (B - 1))
        zp = torch.tensor(0.0)
    else:
        qmin, qmax = 0, 2 ** B - 1
        scale = torch.clamp(rs, min=1e-8) / (2 ** B - 1)
        zp = torch.round(torch.clamp(torch.tensor(float(v_max)) / scale, qmin, qmax))
    return B, np.float32(scale.item()), np.float32(zp.item()), np.float32(qmin), np.float32(qmax)


def quant_codes(x, scale, zp, qmin, qmax):
    """Integer codes q - qmin of quantization.py:183-185 for a float32 array."""
    import torch
    xt = torch.from_numpy(np.ascontiguousarray(x, np.float32))
    q = torch.clamp(torch.round(xt / (torch.tensor(scale) + 1e-8) + torch.tensor(zp)), float(qmin), float(qmax))
    return (q - float(qmin)).numpy().astype(np.int64)


def dequant_codes(codes, scale, zp, qmin):
    """quantization.py:186: (q - zp) * scale in float32."""
    q = codes.astype(np.float32) + np.float32(qmin)
    return ((q - np.float32(zp)) * np.float32(scale)).astype(np.float32)


def pack_bits(codes, bits):
    """Little-endian bit stream: value e occupies bits [e*bits, (e+1)*bits); returned as uint32 words."""
    codes = np.asarray(codes, np.uint64).reshape(-1)
    n = codes.size
    assert n % 32 == 0
    shifts = np.arange(bits, dtype=np.uint64)
    bitmat = ((codes[:, None] >> shifts[None, :]) & np.uint64(1)).astype(np.uint8).reshape(-1)     # LSB first
    by = np.packbits(bitmat, bitorder="little")
    return by.view("<u4").copy()


def unpack_bits(words, bits, n):
    bitmat = np.unpackbits(np.asarray(words, "<u4").view(np.uint8), bitorder="little")[:n * bits].reshape(n, bits)
    weights = (np.uint64(1) << np.arange(bits, dtype=np.uint64))
    return (bitmat.astype(np.uint64) * weights[None, :]).sum(1).astype(np.int64)
