"""Drop-in for the reference's loss.py (loss.py:11-47): the TV regulariser over a random cube of one hash
level and the (unused) Cauchy sparsity loss.  These stay torch code by design (BASELINE.json north_star:
losses and optimiser stay in PyTorch); the hash of the cube corners goes through the kernel."""
from math import exp, floor, log

import torch

from .utils import hash


def total_variation_loss(embeddings, min_resolution, max_resolution, level, log2_hashmap_size, n_levels=16):
    """loss.py:11-43 — same float64 resolution formula, same torch.randint draw (one per level, [3])."""
    min_resolution, max_resolution = int(min_resolution), int(max_resolution)
    b = exp((log(max_resolution) - log(min_resolution)) / (n_levels - 1))
    resolution = floor(min_resolution * b ** level)
    cube_size = int(floor(min(max(resolution / 10.0, min_resolution - 1), 50)))
    dev = embeddings.weight.device
    min_vertex = torch.randint(0, resolution - cube_size, (3,), device=dev)
    idx = min_vertex + torch.arange(cube_size + 1, device=dev)[:, None]
    cube = torch.stack(torch.meshgrid(idx[:, 0], idx[:, 1], idx[:, 2], indexing="ij"), dim=-1)
    e = embeddings(hash(cube, log2_hashmap_size))
    tv = (torch.pow(e[1:] - e[:-1], 2).sum() + torch.pow(e[:, 1:] - e[:, :-1], 2).sum()
          + torch.pow(e[:, :, 1:] - e[:, :, :-1], 2).sum())
    return tv / cube_size


def tv_level_geometry(min_resolution, max_resolution, level, n_levels=16):
    """loss.py:13-22 -> (resolution, cube_size) of one level (float64 formula, as the reference)."""
    min_resolution, max_resolution = int(min_resolution), int(max_resolution)
    b = exp((log(max_resolution) - log(min_resolution)) / (n_levels - 1))
    resolution = floor(min_resolution * b ** level)
    return resolution, int(floor(min(max(resolution / 10.0, min_resolution - 1), 50)))


def total_variation_loss_all(embed_fn):
    """sum_l total_variation_loss(embeddings[l], ...) as the training loop computes it (run_nerf.py:1027-1034),
    in two kernel launches instead of ~60 framework launches per level.  The random cube origins are drawn
    with the same per-level torch.randint calls, in the same order, as the reference."""
    from . import ops
    dev = embed_fn.embeddings[0].weight.device
    if not dev.type == "cuda":
        return sum(total_variation_loss(embed_fn.embeddings[i], embed_fn.base_resolution, embed_fn.finest_resolution, i,
                                        embed_fn.log2_hashmap_size, n_levels=embed_fn.n_levels)
                   for i in range(embed_fn.n_levels))
    mvs, cubes = [], []
    for l in range(embed_fn.n_levels):
        res, cube = tv_level_geometry(embed_fn.base_resolution, embed_fn.finest_resolution, l, embed_fn.n_levels)
        mvs.append(torch.randint(0, res - cube, (3,), device=dev))
        cubes.append(cube)
    losses = ops.TVLossFn.apply(torch.stack(mvs), tuple(cubes), embed_fn.log2_hashmap_size, *embed_fn.tables())
    return losses.sum()


def sigma_sparsity_loss(sigmas):
    """loss.py:45-47."""
    return torch.log(1.0 + 2 * sigmas ** 2).sum(dim=-1)
