"""One optimisation step of the reference's training loop (run_nerf.py:1007-1162) around the kernels:
render the ray batch, image + sparsity + TV losses, backward, (data-parallel gradient all-reduce),
RAdam.  The structural-prior losses are callers of this path (they consume depth_map / normal_map) and
stay in the reference's own torch code."""
import torch

from . import parallel
from .loss import total_variation_loss_all
from .render import render
from .run_nerf_helpers import img2mse, mse2psnr


class Trainer:
    def __init__(self, args, render_kwargs_train, optimizer, H, W, K, near, far, group=None):
        self.args, self.kw, self.opt = args, dict(render_kwargs_train), optimizer
        self.H, self.W, self.K, self.near, self.far = H, W, K, near, far
        self.group = group
        self.world = parallel.world_size(group)
        self.tv_weight = args.tv_loss_weight
        self.step_idx = 0
        self.embed_fn = self.kw["embed_fn"]
        self.nets = [self.kw["network_fn"]] + ([self.kw["network_fine"]] if self.kw.get("network_fine") is not None else [])

    def losses(self, rgb, extras, target_s):
        """run_nerf.py:1010-1037.  With a process group the terms are scaled so that the SUM of the ranks'
        gradients is the gradient of the single-process loss over the concatenated batch."""
        w = float(self.world)
        img_loss = img2mse(rgb, target_s)
        loss = img_loss / w
        if "rgb0" in extras:
            loss = loss + img2mse(extras["rgb0"], target_s) / w
        sparsity = self.args.sparse_loss_weight * (extras["sparsity_loss"].sum() + extras["sparsity_loss0"].sum())
        loss = loss + sparsity
        loss = loss + self.tv_weight * total_variation_loss_all(self.embed_fn) / w
        return loss, img_loss

    def step(self, batch_rays, target_s, chunk=None):
        """batch_rays [2,N,3], target_s [N,3] (this rank's shard).  Returns (loss, psnr) as 0-d device tensors."""
        rgb, depth, acc, extras = render(self.H, self.W, self.K, chunk=chunk or batch_rays.shape[1], rays=batch_rays,
                                         retraw=True, near=self.near, far=self.far, **self.kw)
        self.opt.zero_grad()
        loss, img_loss = self.losses(rgb, extras, target_s)
        loss.backward()
        if self.world > 1:
            parallel.allreduce_gradients(self.embed_fn, self.nets, self.group)
        self.opt.step()
        self.step_idx += 1
        if self.step_idx > 1000:                                   # run_nerf.py:1036-1037
            self.tv_weight = 0.0
        self.decay_learning_rate()
        return loss.detach(), mse2psnr(img_loss.detach())

    def decay_learning_rate(self):
        """run_nerf.py:1289-1293: lr = lrate * 0.1 ** (global_step / (lrate_decay * 1000)) on every group (a host scalar;
        RAdam folds it into its fused update).  The reference evaluates this before `global_step += 1` (:1475), i.e.
        with the number of iterations completed BEFORE the current one."""
        decay = getattr(self.args, "lrate_decay", None)
        if not decay:
            return
        new_lrate = self.args.lrate * (0.1 ** ((self.step_idx - 1) / (decay * 1000)))
        for g in self.opt.param_groups:
            g["lr"] = new_lrate
