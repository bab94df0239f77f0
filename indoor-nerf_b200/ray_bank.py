"""On-GPU ray batching: training batches generated on the fly from (image, pixel) ids.

The reference's ``use_batching`` branch precomputes every training ray on the host (run_nerf.py:896-907:
``get_rays_np`` per pose, concatenate with the images, reshape to ``[N*H*W, 3, 3]`` float32 = 36 B per ray,
``np.random.shuffle``), uploads that tensor (:915) and slices it every iteration (:962-966), re-permuting it
with ``torch.randperm`` after each epoch (:968-973).  At ScanNet size (1296x968) that is 45 MB per image on
top of the image itself.  Here the bank is only the permutation (8 B per ray); a batch is one kernel launch
that turns ids into origins / directions / target colours, bit-identical to slicing the reference's tensor
(ray directions in float64 arithmetic rounded to float32, exactly what get_rays_np + astype produce).

The ``no_batching`` branch (configs/chair.txt) generates the rays of a whole image every iteration and keeps
1024 of them (run_nerf.py:976-1004); ``sample_image`` draws the same pixels from numpy's global RNG and
generates only those rays (float32 arithmetic of ``get_rays``).

Both consume numpy's / torch's global RNGs exactly as the reference does, so seeded runs see the same batches.
With a process group every rank holds the same permutation and takes its contiguous shard of each global
batch (``parallel.shard_rays`` semantics) — no collective per step.
"""
import numpy as np
import torch

from . import ops, parallel


class RayBank:
    def __init__(self, H, W, K, poses, images, i_train, device=None, group=None):
        """poses [N,3,4] / [N,4,4] and images [N,H,W,3] (float32 in [0,1], or uint8) for ALL views, numpy or torch;
        i_train: indices of the training views (run_nerf.py:740-752).  Everything is moved to `device` once."""
        self.H, self.W = int(H), int(W)
        self.K = np.asarray(K, dtype=np.float64)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.poses = torch.as_tensor(np.asarray(poses) if not torch.is_tensor(poses) else poses).float().to(self.device).contiguous()
        img = torch.as_tensor(np.asarray(images) if not torch.is_tensor(images) else images)
        if img.dtype != torch.uint8:
            img = img.float()
        if tuple(img.shape[1:]) != (self.H, self.W, 3):
            raise ValueError("images must be [N,%d,%d,3] (blend RGBA onto the background first, run_nerf.py:765-768)" % (self.H, self.W))
        self.images = img.to(self.device).contiguous()
        self.i_train = [int(i) for i in i_train]
        self.image_index = torch.tensor(self.i_train, dtype=torch.int32, device=self.device)
        # group=None means "not distributed" here, even inside an initialised process group: a bank built by one rank
        # alone (a benchmark leg, an evaluation script) must not start a collective the other ranks never join
        self.group = group
        self.rank, self.world = (parallel.rank(group), parallel.world_size(group)) if group is not None else (0, 1)
        self.order = None
        self.i_batch = 0

    # ---- use_batching ---------------------------------------------------------------------------------------
    @property
    def n_rays(self):
        return len(self.i_train) * self.H * self.W

    def shuffle(self):
        """run_nerf.py:907 — np.random.shuffle(rays_rgb): the same Fisher-Yates draws applied to the ids."""
        order = np.arange(self.n_rays)
        np.random.shuffle(order)
        self.order = torch.from_numpy(order).to(self.device)
        self._sync_order()
        self.i_batch = 0
        return self

    def _sync_order(self):
        if self.world > 1:
            torch.distributed.broadcast(self.order, src=0, group=self.group)

    def next_batch(self, N_rand):
        """run_nerf.py:962-973 -> (batch_rays [2,B,3], target_s [B,3]) — this rank's shard when distributed."""
        if self.order is None:
            self.shuffle()
        ids = self.order[self.i_batch:self.i_batch + N_rand]
        if self.world > 1:
            n = ids.shape[0]
            ids = ids[(n * self.rank) // self.world:(n * (self.rank + 1)) // self.world]
        rays, target = ops.ray_bank_batch(ids, self.H, self.W, self.K, self.poses, self.image_index, self.images,
                                          f64_dirs=True)
        self.i_batch += N_rand
        if self.i_batch >= self.n_rays:
            rand_idx = torch.randperm(self.n_rays, device=self.device)          # :970 (default tensor type is CUDA there)
            self.order = self.order[rand_idx]
            self._sync_order()
            self.i_batch = 0
        return rays, target

    # ---- no_batching ------------------------------------------------------------------------------------------
    def sample_image(self, N_rand, precrop_frac=None, img_i=None):
        """run_nerf.py:976-1004: one random training view (np.random.choice(i_train)), N_rand distinct pixels of it
        (np.random.choice(..., replace=False)), optionally restricted to the central crop of the first
        precrop_iters iterations.  Returns (batch_rays, target_s, img_i)."""
        H, W = self.H, self.W
        if img_i is None:
            img_i = int(np.random.choice(self.i_train))
        if precrop_frac is not None:
            dH, dW = int(H // 2 * precrop_frac), int(W // 2 * precrop_frac)
            r0, c0, nr, nc = H // 2 - dH, W // 2 - dW, 2 * dH, 2 * dW
        else:
            r0, c0, nr, nc = 0, 0, H, W
        sel = np.random.choice(nr * nc, size=[N_rand], replace=False)
        if self.world > 1:
            sel = sel[(N_rand * self.rank) // self.world:(N_rand * (self.rank + 1)) // self.world]
        rows, cols = r0 + sel // nc, c0 + sel % nc
        ids = torch.from_numpy((rows * W + cols).astype(np.int64)).to(self.device, non_blocking=True)
        index = torch.tensor([img_i], dtype=torch.int32, device=self.device)
        rays, target = ops.ray_bank_batch(ids, H, W, self.K, self.poses, index, self.images, f64_dirs=False)
        return rays, target, img_i

    def bytes_resident(self):
        """Device bytes held for batching (permutation + index), next to what the reference's rays_rgb would take."""
        ours = (self.order.numel() * 8 if self.order is not None else self.n_rays * 8) + self.image_index.numel() * 4
        return {"ray_bank": ours, "reference_rays_rgb": self.n_rays * 36}
