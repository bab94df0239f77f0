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
    target = scene.get("target_fn", analytic_rgb)(rays[0], rays[1])
    if pin and torch.cuda.is_available():
        rays, target = rays.pin_memory(), target.pin_memory()
    return rays.to(device), target.to(device)


def _look_at(eye, target, up=(0.0, 0.0, 1.0)):
    """camera-to-world [4,4] in the reference's camera convention (x right, y up, camera looks down -z)."""
    eye, target, up = (np.asarray(v, np.float32) for v in (eye, target, up))
    f = target - eye
    f /= np.linalg.norm(f)
    r = np.cross(f, up)
    r /= np.linalg.norm(r)
    u = np.cross(r, f)
    c2w = np.eye(4, dtype=np.float32)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = r, u, -f, eye
    return c2w


def room_rgb(rays_o, rays_d, lo=(-3.0, -2.5, 0.0), hi=(3.0, 2.5, 2.8)):
    """Analytic 'photo' of a room seen from inside: the exit point of the ray from the axis-aligned box, textured with
    a checker per wall.  Also returns the hit depth (along the un-normalised ray) and the inward wall normal — the
    priors a structural loss would be fed."""
    lo, hi = torch.tensor(lo, device=rays_o.device), torch.tensor(hi, device=rays_o.device)
    inv = 1.0 / torch.where(rays_d.abs() < 1e-9, torch.full_like(rays_d, 1e-9), rays_d)
    t_far = torch.maximum((lo - rays_o) * inv, (hi - rays_o) * inv)
    t, axis = t_far.min(-1)
    hit = rays_o + t[..., None] * rays_d
    chk = (torch.floor(hit * 2.0).sum(-1) % 2.0)
    base = torch.stack([0.55 + 0.15 * axis.float(), 0.6 - 0.1 * axis.float(), 0.5 + 0.2 * (axis == 2).float()], -1)
    rgb = base * (0.7 + 0.3 * chk[..., None])
    normal = -torch.sign(torch.gather(rays_d, -1, axis[..., None])) * torch.nn.functional.one_hot(axis, 3).float()
    return rgb, t, normal


def scannet_scene(H=968, W=1296, n_views=30, focal=1170.0, seed=0):
    """ScanNet-shaped indoor scene (BASELINE configs[3]; configs/scannet_scene0000.txt, load_scannet.py): 1296 x 968
    colour frames, ~30 inward-looking cameras inside a room, near 0.1 / far 10 (run_nerf.py:782-783), bounding box =
    the room's bounds +- 1 (load_scannet.py:105)."""
    rs = np.random.RandomState(seed)
    lo, hi = np.array([-3.0, -2.5, 0.0], np.float32), np.array([3.0, 2.5, 2.8], np.float32)
    poses = []
    for _ in range(n_views):
        eye = np.array([rs.uniform(-1.5, 1.5), rs.uniform(-1.2, 1.2), rs.uniform(1.2, 1.7)], np.float32)
        tgt = np.array([rs.uniform(-3, 3), rs.uniform(-2.5, 2.5), rs.uniform(0.3, 2.3)], np.float32)
        if np.linalg.norm(tgt[:2] - eye[:2]) < 0.5:
            tgt[:2] += 1.0
        poses.append(_look_at(eye, tgt))
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    bbox = (torch.from_numpy(lo - 1.0), torch.from_numpy(hi + 1.0))
    return dict(H=H, W=W, focal=focal, K=K, poses=np.stack(poses), near=0.1, far=10.0, bounding_box=bbox,
                target_fn=lambda o, d: room_rgb(o, d)[0], prior_fn=room_rgb, ndc=False, white_bkgd=False)


def llff_scene(H=756, W=1008, n_views=20, focal=815.0, seed=0):
    """LLFF-shaped forward-facing scene (BASELINE configs[4]; configs/fern.txt at factor 4: 1008 x 756): a ~20-pose rig
    on a small planar patch looking down -z, rendered in NDC (near 0, far 1, run_nerf.py:757-759), bounding box by
    utils.get_bbox3d_for_llff (utils.py:61-92)."""
    rs = np.random.RandomState(seed)
    poses = []
    for _ in range(n_views):
        eye = np.array([rs.uniform(-0.6, 0.6), rs.uniform(-0.4, 0.4), rs.uniform(-0.05, 0.05)], np.float32)
        tgt = np.array([eye[0] * 0.3, eye[1] * 0.3, -4.0], np.float32)
        poses.append(_look_at(eye, tgt, up=(0.0, 1.0, 0.0)))
    poses = np.stack(poses)
    K = np.array([[focal, 0, 0.5 * W], [0, focal, 0.5 * H], [0, 0, 1]])
    bbox = utils.get_bbox3d_for_llff(poses[:, :3, :4], [H, W, focal], near=0.0, far=1.0)

    def target(o, d):
        dn = d / d.norm(dim=-1, keepdim=True)
        return torch.stack([0.5 + 0.5 * torch.sin(6 * dn[..., 0]), 0.5 + 0.5 * torch.cos(5 * dn[..., 1]),
                            0.5 + 0.4 * torch.sin(9 * dn[..., 0] * dn[..., 1])], -1)
    return dict(H=H, W=W, focal=focal, K=K, poses=poses, near=0.0, far=1.0, bounding_box=bbox, target_fn=target,
                ndc=True, white_bkgd=False)
