"""Drop-in for the hot-path names of the reference's utils.py (hash, get_voxel_vertices, BOX_OFFSETS,
bbox helpers; utils.py:9-117)."""
import numpy as np
import torch

from . import ops
from .ray_utils import get_ndc_rays, get_ray_directions, get_rays

# corner order of the 8 voxel vertices, id = 4dx + 2dy + dz (utils.py:9); created lazily on the device asked for
_BOX_OFFSETS_LIST = [[[i, j, k] for i in (0, 1) for j in (0, 1) for k in (0, 1)]]
BOX_OFFSETS = torch.tensor(_BOX_OFFSETS_LIST)

PRIMES = (1, 2654435761, 805459861, 3674653429, 2097192037, 1434869437, 2165219737)


def hash(coords, log2_hashmap_size):
    """utils.py:13-24.  CUDA tensors with <= 3 coordinates (every call on the training path: the TV loss,
    loss.py:29) go through the kernel.  The defining integer formula below is only reached by host-side set-up
    code and unit tests holding CPU tensors, or by the reference's unused 4..7-D case."""
    if coords.is_cuda and coords.shape[-1] <= 3:
        return ops.hash_coords(coords, log2_hashmap_size)
    acc = torch.zeros_like(coords)[..., 0]
    for i in range(coords.shape[-1]):
        acc ^= coords[..., i] * PRIMES[i]
    return torch.tensor((1 << log2_hashmap_size) - 1).to(acc.device) & acc


def get_voxel_vertices(xyz, bounding_box, resolution, log2_hashmap_size):
    """utils.py:95-117 -> (voxel_min_vertex, voxel_max_vertex, hashed_voxel_indices[B,8], keep_mask[B,3]).
    The encoder itself never calls this (the kernel fuses it); it exists for callers of the reference API."""
    box_min, box_max = bounding_box
    box_min, box_max = box_min.to(xyz.device), box_max.to(xyz.device)
    keep_mask = xyz == torch.max(torch.min(xyz, box_max), box_min)
    xc = torch.clamp(xyz, min=box_min, max=box_max)
    grid_size = (box_max - box_min) / resolution
    bottom_left_idx = torch.floor((xc - box_min) / grid_size).int()
    voxel_min_vertex = bottom_left_idx * grid_size + box_min
    voxel_max_vertex = voxel_min_vertex + grid_size
    if xyz.is_cuda:
        grid = ops.make_grid(box_min.tolist(), box_max.tolist(), [float(resolution)], log2_hashmap_size)
        hashed = ops.hash_indices(grid, xyz)[:, 0, :].long()
    else:
        hashed = hash(bottom_left_idx.unsqueeze(1) + BOX_OFFSETS, log2_hashmap_size)
    return voxel_min_vertex, voxel_max_vertex, hashed, keep_mask


def _bbox_from_rays(ray_sets, near, far, pad):
    lo, hi = torch.full((3,), 100.0), torch.full((3,), -100.0)
    for rays_o, rays_d, H, W in ray_sets:
        idx = torch.tensor([0, W - 1, H * W - W, H * W - 1])
        pts = torch.cat([rays_o[idx] + near * rays_d[idx], rays_o[idx] + far * rays_d[idx]], 0)
        lo, hi = torch.minimum(lo, pts.min(0)[0]), torch.maximum(hi, pts.max(0)[0])
    pad = torch.tensor(pad)
    return lo - pad, hi + pad


def get_bbox3d_for_blenderobj(camera_transforms, H, W, near=2.0, far=6.0):
    """utils.py:27-58: AABB of the near/far points of the four corner rays of every training camera, +-1."""
    focal = 0.5 * W / np.tan(0.5 * float(camera_transforms["camera_angle_x"]))
    directions = get_ray_directions(H, W, focal)
    sets = []
    for frame in camera_transforms["frames"]:
        c2w = torch.tensor(frame["transform_matrix"], dtype=torch.float32)
        rays_o, rays_d = get_rays(directions, c2w)
        sets.append((rays_o, rays_d, H, W))
    return _bbox_from_rays(sets, near, far, [1.0, 1.0, 1.0])


def get_bbox3d_for_llff(poses, hwf, near=0.0, far=1.0):
    """utils.py:61-92."""
    H, W, focal = hwf
    H, W = int(H), int(W)
    directions = get_ray_directions(H, W, focal)
    sets = []
    for pose in torch.tensor(np.asarray(poses), dtype=torch.float32):
        rays_o, rays_d = get_rays(directions, pose)
        rays_o, rays_d = get_ndc_rays(H, W, focal, 1.0, rays_o, rays_d)
        sets.append((rays_o, rays_d, H, W))
    return _bbox_from_rays(sets, near, far, [0.1, 0.1, 0.0001])
