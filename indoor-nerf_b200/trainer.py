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
    def __init__(self, args, render_kwargs_train, optimizer, H, W, K, near, far, group=None, start=0):
        """`start` is create_nerf's third return value (the checkpoint's global_step, run_nerf.py:299): the learning-rate
        decay and the TV cut-off continue from it after a resume.  `group=None` means a single-process trainer even
        inside an initialised process group (same convention as RayBank / render_path); pass the group to shard."""
        self.args, self.kw, self.opt = args, dict(render_kwargs_train), optimizer
        self.H, self.W, self.K, self.near, self.far = H, W, K, near, far
        self.group = group
        self.world = parallel.world_size(group) if group is not None else 1
        self.tv_weight = args.tv_loss_weight
        self.step_idx = int(start)                                 # iterations completed; TV is cut after > 1000 (:1036)
        self.embed_fn = self.kw["embed_fn"]
        self.nets = [self.kw["network_fn"]] + ([self.kw["network_fine"]] if self.kw.get("network_fine") is not None else [])
        self._uncalibrated = [q for q in parallel.model_quantizers(self.embed_fn, self.nets) if not q.calibrated]
        self.extra_loss_fn = None
        if self.step_idx > 0:
            self.decay_learning_rate(self.step_idx)                # the lr of the iteration about to run

    def losses(self, rgb, extras, target_s, depth=None):
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
        if self.extra_loss_fn is not None:
            # consumers of depth_map / normal_map (the reference adds its structural-prior losses here,
            # run_nerf.py:1043-1148; they stay the caller's torch code) — mean-type terms, hence / world
            loss = loss + self.extra_loss_fn(rgb, depth, extras) / w
        return loss, img_loss

    def step(self, batch_rays, target_s, chunk=None):
        """batch_rays [2,N,3], target_s [N,3] (this rank's shard).  Returns (loss, psnr) as 0-d device tensors."""
        rgb, depth, acc, extras = render(self.H, self.W, self.K, chunk=chunk or batch_rays.shape[1], rays=batch_rays,
                                         retraw=True, near=self.near, far=self.far, **self.kw)
        self._sync_fresh_quantizers()
        self.opt.zero_grad()
        loss, img_loss = self.losses(rgb, extras, target_s, depth)
        loss.backward()
        if self.world > 1:
            from . import ops
            if ops.KERNEL_EVENTS is not None:                      # bench.py's live timing of the step's one collective
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                parallel.allreduce_gradients(self.embed_fn, self.nets, self.group)
                e1.record()
                ops.KERNEL_EVENTS.append(("allreduce_gradients", batch_rays.shape[1], e0, e1))
            else:
                parallel.allreduce_gradients(self.embed_fn, self.nets, self.group)
        self.opt.step()
        self.step_idx += 1
        if self.step_idx > 1000:                                   # run_nerf.py:1036-1037
            self.tv_weight = 0.0
        self.decay_learning_rate(self.step_idx - 1)
        return loss.detach(), mse2psnr(img_loss.detach())

    def _sync_fresh_quantizers(self):
        """Quantisers calibrate from the LOCAL shard during the forward (quantization.py:97-119).  Which ones calibrate in
        a given step depends only on step counters, so every rank has the same `fresh` list: make their statistics the
        global min / max before any rank fake-quantises with them again (else the replicas diverge and the summed
        gradient is no longer the single-process gradient)."""
        if self.world <= 1 or not self._uncalibrated:
            return
        fresh = [q for q in self._uncalibrated if q.calibrated]
        if fresh:
            parallel.sync_quantizer_calibration(fresh, self.group)
            self._uncalibrated = [q for q in self._uncalibrated if not q.calibrated]

    def decay_learning_rate(self, global_step):
        """run_nerf.py:1289-1293: lr = lrate * 0.1 ** (global_step / (lrate_decay * 1000)) on every group (a host scalar;
        RAdam folds it into its fused update).  The reference evaluates this at the END of iteration i with
        global_step == i, before `global_step += 1` (:1475): the lr set here is the one the NEXT iteration uses."""
        decay = getattr(self.args, "lrate_decay", None)
        if not decay:
            return
        new_lrate = self.args.lrate * (0.1 ** (global_step / (decay * 1000)))
        for g in self.opt.param_groups:
            g["lr"] = new_lrate
