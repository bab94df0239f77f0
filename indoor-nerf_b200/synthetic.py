"""Synthetic scenes of the shapes BASELINE.json names (no dataset is available offline): Blender-style
inward-looking cameras on the upper hemisphere (load_blender.py:30-35), the chair config's intrinsics
(configs/chair.txt + load_blender.py), an analytic target image, and the scene bounding box computed the
way utils.get_bbox3d_for_blenderobj does it (utils.py:27-58)."""
import numpy as np
import torch

from . import utils


def pose_spherical(theta, phi, radius):
    """load_blender.py:10-35 (trans_t, rot_phi, rot_theta composition)."""
    t = np.eye(4, dtype=np.float32); t[2, 3] = radius
    p = phi / 180. * np.pi
    rp = np.array([[1, 0, 0, 0], [0, np.cos(p), -np.sin(p), 0], [0, np.sin(p), np.cos(p), 0], [0, 0, 0, 1]], np.float32)
    th = theta / 180. * np.pi
    rt = np.array([[np.cos(th), 0, -np.sin(th), 0], [0, 1, 0, 0], [np.sin(th), 0, np.cos(th), 0], [0, 0, 0, 1]], np.float32)
    c2w = rt @ rp @ t
    return (np.array([[-1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], np.float32) @ c2w).astype(np.float32)


def blender_scene(H=400, W=400, n_views=100, camera_angle_x=0.6911112070083618, seed=0):
    """Chair-shaped synthetic scene: returns dict(H, W, focal, K, poses[n,4,4], near, far, bounding_box)."""
    rs = np.random.RandomState(seed)
    focal = .5 * W / np.tan(.5 * camera_angle_x)
    poses = np.stack([pose_spherical(rs.uniform(-180, 180), rs.uniform(-90, -10), 4.0) for _ in range(n_views)])
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    frames = {"camera_angle_x": camera_angle_x, "frames": [{"transform_matrix": p.tolist()} for p in poses]}
    bbox = utils.get_bbox3d_for_blenderobj(frames, H, W, near=2.0, far=6.0)
    return dict(H=H, W=W, focal=focal, K=K, poses=poses, near=2.0, far=6.0, bounding_box=bbox)


def analytic_rgb(rays_o, rays_d):
    """A cheap analytic 'photo': a shaded unit sphere at the origin on a white background."""
    o, d = rays_o, rays_d / rays_d.norm(dim=-1, keepdim=True)
    b = (o * d).sum(-1)
    c = (o * o).sum(-1) - 1.0
    disc = b * b - c
    hit = disc > 0
    t = -b - torch.sqrt(disc.clamp(min=0))
    n = torch.nn.functional.normalize(o + t[..., None] * d, dim=-1)
    shade = (0.5 + 0.5 * n) * (0.3 + 0.7 * n[..., 2:3].clamp(min=0))
    return torch.where(hit[..., None], shade, torch.ones_like(shade))


def ray_batch(scene, n_rays, seed, device="cpu", pin=False):
    """A training batch like run_nerf.py:962-1004 produces: batch_rays [2,N,3] and target_s [N,3], rays drawn
    from random pixels of random training views (host tensors; pinned on request)."""
    rs = np.random.RandomState(seed)
    H, W, K = scene["H"], scene["W"], scene["K"]
    view = rs.randint(0, len(scene["poses"]), n_rays)
    i = rs.randint(0, W, n_rays).astype(np.float32)
    j = rs.randint(0, H, n_rays).astype(np.float32)
    dirs = np.stack([(i - K[0][2]) / K[0][0], -(j - K[1][2]) / K[1][1], -np.ones_like(i)], -1).astype(np.float32)
    R = scene["poses"][view, :3, :3]
    rays_d = np.einsum("nk,nck->nc", dirs, R).astype(np.float32)
    rays_o = scene["poses"][view, :3, 3].astype(np.float32)
    rays = torch.from_numpy(np.stack([rays_o, rays_d], 0))
    target = analytic_rgb(rays[0], rays[1])
    if pin and torch.cuda.is_available():
        rays, target = rays.pin_memory(), target.pin_memory()
    return rays.to(device), target.to(device)
