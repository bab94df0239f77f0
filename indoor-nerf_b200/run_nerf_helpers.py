"""Drop-in for the hot-path names of the reference's run_nerf_helpers.py: NeRFSmall, get_embedder,
sample_pdf, get_rays, get_rays_np, ndc_rays, img2mse, mse2psnr, to8b (run_nerf_helpers.py:11-13,
50-80, 169-397), backed by the sm_100a kernels.  The vanilla 8x256 NeRF / sin-cos Embedder
(`--i_embed 0`) are not on the HashNeRF path and are not provided."""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .hash_encoding import HashEmbedder, SHEncoder
from .quantization import FakeQuantizer, LearnedBitwidthQuantizer  # noqa: F401  (star-exported like the reference)

img2mse = lambda x, y: torch.mean((x - y) ** 2)
mse2psnr = lambda x: -10. * torch.log(x) / torch.log(torch.tensor([10.], device=x.device))
to8b = lambda x: (255 * np.clip(x, 0, 1)).astype(np.uint8)


def get_embedder(multires, args, i=0):
    """run_nerf_helpers.py:50-80 for the hash (i=1), SH (i=2) and identity (i=-1) encoders."""
    if i == -1:
        return nn.Identity(), 3
    if i == 1:
        embed = HashEmbedder(bounding_box=args.bounding_box, log2_hashmap_size=args.log2_hashmap_size,
                             finest_resolution=args.finest_res,
                             use_quantization=getattr(args, "use_quantization", False),
                             quantization_bits=getattr(args, "quantization_bits", 8))
        return embed, embed.out_dim
    if i == 2:
        embed = SHEncoder()
        return embed, embed.out_dim
    raise NotImplementedError("i_embed=0 (sin/cos positional encoding) is outside the HashNeRF path")


class NeRFSmall(nn.Module):
    """run_nerf_helpers.py:169-306.  Same constructor (plus the ``predict_normals`` kwarg that
    create_nerf passes at run_nerf.py:268 and the reference's __init__ forgets to accept), same
    sub-module names and therefore the same state_dict keys.  ``forward`` is one fused kernel."""

    def __init__(self, num_layers=3, hidden_dim=64, geo_feat_dim=15, num_layers_color=4, hidden_dim_color=64,
                 input_ch=3, input_ch_views=3, use_quantization=False, quantization_bits=8, predict_normals=False):
        super().__init__()
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.num_layers, self.num_layers_color = num_layers, num_layers_color
        self.hidden_dim, self.hidden_dim_color, self.geo_feat_dim = hidden_dim, hidden_dim_color, geo_feat_dim
        self.use_quantization = use_quantization
        self.predict_normals = bool(predict_normals)
        self.sigma_net = nn.ModuleList([
            nn.Linear(input_ch if l == 0 else hidden_dim, 1 + geo_feat_dim if l == num_layers - 1 else hidden_dim,
                      bias=False) for l in range(num_layers)])
        if use_quantization:
            self.sigma_act_quantizers = nn.ModuleList([
                LearnedBitwidthQuantizer(init_bits=float(quantization_bits), min_bits=2.0, max_bits=32.0,
                                         symmetric=False) for _ in range(num_layers - 1)])
            self.sigma_weight_quantizer = LearnedBitwidthQuantizer(init_bits=float(quantization_bits), min_bits=2.0,
                                                                   max_bits=32.0, symmetric=True)
        else:
            self.sigma_act_quantizers = None
            self.sigma_weight_quantizer = None
        self.color_net = nn.ModuleList([
            nn.Linear(input_ch_views + geo_feat_dim if l == 0 else hidden_dim,
                      3 if l == num_layers_color - 1 else hidden_dim, bias=False) for l in range(num_layers_color)])
        if self.predict_normals:
            self.normal_net = nn.Sequential(nn.Linear(geo_feat_dim, hidden_dim // 2), nn.ReLU(),
                                            nn.Linear(hidden_dim // 2, 3))

    # -- kernel-facing views ------------------------------------------------------------------------------
    def _check_shapes(self):
        ok = (self.num_layers == 2 and self.num_layers_color == 3 and self.hidden_dim == 64 and
              self.hidden_dim_color == 64 and self.geo_feat_dim == 15 and self.input_ch == 32 and
              self.input_ch_views == 16)
        if not ok:
            raise NotImplementedError(
                "the fused NeRFSmall kernel is built for the create_nerf shapes (run_nerf.py:240-247): "
                "num_layers=2, hidden 64, geo 15, num_layers_color=3, input_ch=32, input_ch_views=16")

    def kernel_weights(self):
        """(keys, tensors): the weights in kernel order; the first sigma layer passes through its weight
        fake-quantiser when quantisation is on (run_nerf_helpers.py:272-276)."""
        self._check_shapes()
        w0 = self.sigma_net[0].weight
        if self.use_quantization and self.sigma_weight_quantizer is not None:
            w0 = self.sigma_weight_quantizer(w0)
        keys = ["s0", "s1", "c0", "c1", "c2"]
        tensors = [w0, self.sigma_net[1].weight, self.color_net[0].weight, self.color_net[1].weight,
                   self.color_net[2].weight]
        if self.predict_normals:
            keys += ["n0w", "n0b", "n2w", "n2b"]
            tensors += [self.normal_net[0].weight, self.normal_net[0].bias, self.normal_net[2].weight,
                        self.normal_net[2].bias]
        return tuple(keys), tensors

    def act_qrow(self, feat_for_calibration=None, w0=None):
        """Device row for the fused activation fake-quant (run_nerf_helpers.py:281-284) or None."""
        if not (self.use_quantization and self.sigma_act_quantizers is not None):
            return None
        q = self.sigma_act_quantizers[0]
        if self.training and not q.calibrated:
            # one-off statistics pass (first quantised training call): min / max of relu(x W0^T)
            with torch.no_grad():
                h = torch.relu(feat_for_calibration.detach() @ w0.detach().t())
                q.calibrate_minmax(h.min(), h.max())
        return q.qrow(self.training).contiguous()

    def forward(self, x):
        keys, tensors = self.kernel_weights()
        x2 = x.reshape(-1, self.input_ch + self.input_ch_views)
        act_q = self.act_qrow(x2[:, :32], tensors[0])
        out = ops.MlpFn.apply(x2, act_q, keys, *tensors)
        return out.reshape(*x.shape[:-1], out.shape[-1])


def get_rays(H, W, K, c2w):
    """run_nerf_helpers.py:311-320 -> (rays_o, rays_d), each [H,W,3] on the GPU."""
    dev = c2w.device if (torch.is_tensor(c2w) and c2w.is_cuda) else torch.device("cuda", torch.cuda.current_device())
    return ops.gen_rays(H, W, K, c2w, dev)


def get_rays_np(H, W, K, c2w):
    """run_nerf_helpers.py:323-330 (host-side, numpy; used while building the ray bank)."""
    i, j = np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32), indexing="xy")
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1)
    rays_d = np.sum(dirs[..., np.newaxis, :] * c2w[:3, :3], -1)
    return np.broadcast_to(c2w[:3, -1], np.shape(rays_d)), rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """run_nerf_helpers.py:333-350."""
    return ops.ndc_rays(H, W, focal, near, rays_o, rays_d)


def sample_pdf(bins, weights, N_samples, det=False, pytest=False):
    """run_nerf_helpers.py:354-397.  Draws u from torch's global RNG exactly where the reference does
    (same shape, same order) so that seeded runs sample the same bins."""
    N = bins.shape[0]
    dev = bins.device
    if det:
        u = torch.linspace(0., 1., steps=N_samples, device=dev)
    else:
        u = torch.rand([N, N_samples], device=dev)
    if pytest:
        np.random.seed(0)
        if det:
            u = torch.tensor(np.linspace(0., 1., N_samples), dtype=torch.float32, device=dev)
        else:
            u = torch.tensor(np.random.rand(N, N_samples), dtype=torch.float32, device=dev)
    return ops.sample_pdf(bins, weights, u)
