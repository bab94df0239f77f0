"""TEST INFRASTRUCTURE — golden vectors for the data formats either side of the hot path,
from the LIVE, unmodified reference.  Build container only (needs /root/reference):

    python -m oracle.make_golden_io

  ray_bank.npz      run_nerf.py:899-907 executed with the reference's own get_rays_np on a small scene:
                    the unshuffled and the np.random.seed(0)-shuffled rays_rgb tensor
  quant_export.npz  the reference's LearnedBitwidthQuantizer (asymmetric and symmetric, several learned
                    widths) calibrated on and applied to a table in eval mode
  eval_psnr.npz     render_path's PSNR expression (run_nerf.py:186) and compute_metrics' (evaluation_utils.py:24-25)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.make_golden import save  # noqa: E402


def small_scene(seed=3, n=5, H=7, W=5):
    from indoor_nerf_b200.synthetic import pose_spherical
    rs = np.random.RandomState(seed)
    poses = np.stack([pose_spherical(rs.uniform(-180, 180), rs.uniform(-90, -10), 4.0) for _ in range(n)]).astype(np.float32)
    focal = .5 * W / np.tan(.5 * 0.6911112070083618)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])                  # run_nerf.py:832-836 (float64)
    images = rs.rand(n, H, W, 3).astype(np.float32)
    return H, W, K, poses, images


def main():
    ref = ref_shim.load()
    RH, Q = ref.run_nerf_helpers, ref.quantization

    # ---- ray bank: run_nerf.py:899-907 verbatim in structure, calling the reference's get_rays_np ----------
    H, W, K, poses, images = small_scene()
    i_train = [0, 2, 3, 4]
    rays = np.stack([RH.get_rays_np(H, W, K, p) for p in poses[:, :3, :4]], 0)
    rays_rgb = np.concatenate([rays, images[:, None]], 1)
    rays_rgb = np.transpose(rays_rgb, [0, 2, 3, 1, 4])
    rays_rgb = np.stack([rays_rgb[i] for i in i_train], 0)
    rays_rgb = np.reshape(rays_rgb, [-1, 3, 3])
    assert rays_rgb.dtype == np.float64                     # the float64 K promoted everything
    rays_rgb = rays_rgb.astype(np.float32)
    unshuffled = rays_rgb.copy()
    np.random.seed(0)
    np.random.shuffle(rays_rgb)
    np.random.seed(0)
    order = np.arange(unshuffled.shape[0])
    np.random.shuffle(order)
    assert (unshuffled[order] == rays_rgb).all()            # row shuffle == arange shuffle
    save("ray_bank", H=H, W=W, K=K, poses=poses, images=images, i_train=np.array(i_train), unshuffled=unshuffled,
         shuffled=rays_rgb, order=order)

    # ---- quantiser eval form -------------------------------------------------------------------------------
    rs = np.random.RandomState(5)
    table = (rs.uniform(-1e-4, 1e-4, 4096)).astype(np.float32)
    table[:64] *= 40.0                                                          # a few trained-looking outliers
    w0 = (rs.randn(2048) * 0.2).astype(np.float32)
    out = {"table": table, "w0": w0}
    for tag, sym, x, bits_list in (("asym", False, table, (2.0, 4.7, 8.0, 11.5, 16.2, 24.0)),
                                   ("sym", True, w0, (3.4, 8.0, 12.0))):
        for k, b in enumerate(bits_list):
            q = Q.LearnedBitwidthQuantizer(init_bits=8.0, min_bits=2.0, max_bits=32.0, symmetric=sym)
            q.train()
            q(torch.from_numpy(x))                                              # calibrates (quantization.py:97-119)
            q.soft_bits.data.fill_(b)
            q.eval()
            with torch.no_grad():
                y = q(torch.from_numpy(x))
            out["%s_%d_bits" % (tag, k)] = np.float32(b)
            out["%s_%d_range_scale" % (tag, k)] = q.range_scale.detach().numpy()
            out["%s_%d_v_max" % (tag, k)] = q.v_max.detach().numpy() if q.v_max is not None else np.float32(0)
            out["%s_%d_out" % (tag, k)] = y.numpy()
    save("quant_export", **out)

    # ---- PSNR expressions --------------------------------------------------------------------------------------
    rs = np.random.RandomState(9)
    gt = rs.rand(40, 36, 3).astype(np.float32)
    rgb = np.clip(gt + rs.randn(40, 36, 3).astype(np.float32) * 0.05, 0, 1).astype(np.float32)
    p_render_path = -10. * np.log10(np.mean(np.square(rgb - gt)))                                   # run_nerf.py:186
    mse = torch.mean((torch.from_numpy(rgb) - torch.from_numpy(gt)) ** 2)                           # evaluation_utils.py:24
    p_eval = (-10. * torch.log(mse) / torch.log(torch.Tensor([10.]))).item()                       # :25
    save("eval_psnr", rgb=rgb, gt=gt, p_render_path=np.float64(p_render_path), p_eval=np.float64(p_eval),
         rgb8=RH.to8b(rgb))


if __name__ == "__main__":
    main()
