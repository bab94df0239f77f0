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
    GRAPH_WARMUP = 8        # eager steps before a capture: lazy allocations done, RAdam past N_sma >= 5 (step 6)

    def __init__(self, args, render_kwargs_train, optimizer, H, W, K, near, far, group=None, start=0, cuda_graph=False):
        """`start` is create_nerf's third return value (the checkpoint's global_step, run_nerf.py:299): the learning-rate
        decay and the TV cut-off continue from it after a resume.  `group=None` means a single-process trainer even
        inside an initialised process group (same convention as RayBank / render_path); pass the group to shard.

        `cuda_graph=True`: after GRAPH_WARMUP eager steps the whole iteration — render, losses, backward, the gradient
        all-reduce, RAdam — is recorded once per (batch shape, TV on/off) as a CUDA graph and replayed; per step the host
        copies the batch into the graph's input buffers, stores RAdam's two scalars per group and launches the graph
        (~0.1 ms instead of ~3 ms of Python and ~150 launches: what bounds the step at small per-GPU batches, i.e. strong
        scaling).  Same arithmetic and same use of torch's CUDA generator as the eager step.  Runs the eager step
        while quantisers are in play (their calibration and bit schedule are host logic) or when the optimiser is not
        this package's RAdam on CUDA; a consumer loss that synchronises with the host cannot be recorded and raises."""
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
        self.cuda_graph = bool(cuda_graph)
        self._graph = None                                         # dict: key, graph, static inputs / outputs, dyn
        self._eager_steps = 0
        self.graph_launches = 0                                    # kernels of this package launched by graph replays
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
        if self.cuda_graph and self._eager_steps >= self.GRAPH_WARMUP:
            out = self._step_graphed(batch_rays, target_s, chunk)
            if out is not None:
                return out
        self._eager_steps += 1
        return self._step_eager(batch_rays, target_s, chunk)

    # ---- the iteration as a CUDA graph ----------------------------------------------------------------------------
    def _graph_groups(self):
        """The optimiser's (group, live parameters) list when this iteration may run as a recorded graph, else None."""
        from . import ops
        if ops.KERNEL_EVENTS is not None or not hasattr(self.opt, "graph_capture_step"):
            return None
        if getattr(self.embed_fn, "use_quantization", False) or any(getattr(n, "use_quantization", False) for n in self.nets):
            return None
        return self.opt._graph_groups()

    def _capture(self, batch_rays, target_s, chunk, key):
        from . import _lib
        l0 = _lib.launch_count()
        g = {"key": key, "rays": batch_rays.detach().clone(), "target": target_s.detach().clone(),
             "dyn": torch.zeros((len(self.opt.param_groups), 2), dtype=torch.float32, device=batch_rays.device),
             "log10": torch.log(torch.tensor([10.], device=batch_rays.device))}   # mse2psnr's divisor, made outside the capture
        step_before = self.embed_fn.current_step
        graph = torch.cuda.CUDAGraph()
        # thread_local: a process group's watchdog thread may query its events while this thread captures
        try:
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                rgb, depth, acc, extras = render(self.H, self.W, self.K, chunk=chunk or batch_rays.shape[1], rays=g["rays"],
                                                 retraw=True, near=self.near, far=self.far, **self.kw)
                self.opt.zero_grad()
                loss, img_loss = self.losses(rgb, extras, g["target"], depth)
                loss.backward()
                if self.world > 1:
                    parallel.allreduce_gradients(self.embed_fn, self.nets, self.group)
                self.opt.graph_capture_step(g["dyn"])
                psnr = -10. * torch.log(img_loss.detach()) / g["log10"]         # mse2psnr, run_nerf_helpers.py:15
                g["out"] = torch.cat([loss.detach().reshape(1), psnr.reshape(1)])
            g["steps_per_iter"] = self.embed_fn.current_step - step_before  # hash_encoding.py:83-84, per network call
        finally:
            self.embed_fn.current_step = step_before                        # recording is not an iteration
        g["graph"] = graph
        g["launches"] = _lib.launch_count() - l0                           # this package's kernels in one replay
        return g

    def _step_graphed(self, batch_rays, target_s, chunk):
        from . import ops
        groups = self._graph_groups()
        if groups is None:
            return None
        key = (tuple(batch_rays.shape), tuple(target_s.shape), chunk, self.tv_weight, self.world, id(self.extra_loss_fn),
               ops.get_mlp_mode(), self.opt.graph_fingerprint(groups))
        if self._graph is None or self._graph["key"] != key:
            self._graph = None                                     # release the old graph's pool first
            try:
                g = self._capture(batch_rays, target_s, chunk, key)
            except Exception as ex:
                # e.g. a consumer loss that synchronises with the host.  No silent fallback: a capture that died half way
                # leaves torch's CUDA generator registered with it, so the process is not in a state to train on
                self.cuda_graph = False
                raise RuntimeError("Trainer(cuda_graph=True): the iteration could not be recorded as a CUDA graph (%r); "
                                   "construct the Trainer with cuda_graph=False for this configuration" % (ex,)) from ex
            # the recorded backward may have given parameters outside the gradient arena new .grad tensors: the key the
            # next step computes must be the one this graph is filed under
            groups = self._graph_groups()
            if groups is None:
                return None
            g["key"] = key[:-1] + (self.opt.graph_fingerprint(groups),)
            self._graph = g
        g = self._graph
        vals = self.opt.graph_advance(groups)
        if vals is None:
            return None
        g["rays"].copy_(batch_rays, non_blocking=True)
        g["target"].copy_(target_s, non_blocking=True)
        ops.store_floats(g["dyn"], vals)
        g["graph"].replay()
        self.graph_launches += g["launches"] + 1                   # + the scalar store above
        self.embed_fn.current_step += g["steps_per_iter"]
        out = g["out"].clone()
        self._end_of_iteration()
        return out[0], out[1:2]                                    # shapes of the eager step: loss 0-d, psnr [1]

    def release_graph(self):
        """Drop the recorded iteration (its memory pool and, with a process group, the collective it holds).  Call before
        destroying the process group: a communicator must outlive every graph that recorded work on it."""
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None

    def _end_of_iteration(self):
        self.step_idx += 1
        if self.step_idx > 1000:                                   # run_nerf.py:1036-1037
            self.tv_weight = 0.0
        self.decay_learning_rate(self.step_idx - 1)

    def _step_eager(self, batch_rays, target_s, chunk=None):
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
        self._end_of_iteration()
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
