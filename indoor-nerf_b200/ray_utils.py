"""Drop-in for the reference's ray_utils.py (ray_utils.py:5-98).  These run once at set-up, on the host,
to size the scene bounding box (utils.py:27-92); they are not kernels.  kornia's create_meshgrid is
replaced by a two-line torch equivalent (kornia is not a dependency of this package)."""
import torch


def _meshgrid_xy(H, W):
    xs = torch.linspace(0, W - 1, W)
    ys = torch.linspace(0, H - 1, H)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")
    return gx, gy


def get_ray_directions(H, W, focal):
    """ray_utils.py:5-28: camera-frame directions [(i-W/2)/f, -(j-H/2)/f, -1], shape (H, W, 3)."""
    i, j = _meshgrid_xy(H, W)
    return torch.stack([(i - W / 2) / focal, -(j - H / 2) / focal, -torch.ones_like(i)], -1)


def get_rays(directions, c2w):
    """ray_utils.py:31-54: world-frame origins and NORMALISED directions, each (H*W, 3)."""
    rays_d = directions @ c2w[:3, :3].T
    rays_d = rays_d / torch.norm(rays_d, dim=-1, keepdim=True)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def get_ndc_rays(H, W, focal, near, rays_o, rays_d):
    """ray_utils.py:57-98."""
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz = rays_o[..., 0] / rays_o[..., 2]
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    o0 = -1. / (W / (2. * focal)) * ox_oz
    o1 = -1. / (H / (2. * focal)) * oy_oz
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz)
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz)
    d2 = 1 - o2
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)
