"""Drop-in for the hot-path half of the reference's evaluation_utils.py (ComprehensiveEvaluator,
evaluation_utils.py:11-104): test-set rendering with PSNR and SSIM reduced on the GPU.

The reference moves every rendered frame to the host, computes SSIM there with scikit-image and PSNR with a
``.item()`` per frame (evaluation_utils.py:22-43).  Here both are one kernel each on the frame where it was
rendered (``pn_image_sqerr``, ``pn_image_ssim``); the per-frame scalars stay on the device and are read back
once per test set.  LPIPS needs the pretrained AlexNet weights of the ``lpips`` package, which are not part
of this repository: pass ``lpips_fn`` (any callable on two [1,3,H,W] tensors in [-1,1]) or get NaN.
The figure-generation helpers (matplotlib) are plotting and out of scope.
"""
import pickle
import time

import numpy as np
import torch

from . import ops, parallel


class ComprehensiveEvaluator:
    def __init__(self, device="cuda", lpips_fn=None):
        self.device = device
        if lpips_fn is None:
            try:                                                   # evaluation_utils.py:14
                import lpips
                lpips_fn = lpips.LPIPS(net="alex").to(device)
            except Exception:
                lpips_fn = None
        self.lpips_fn = lpips_fn
        self.metrics_history = {
            "train": {"iter": [], "psnr": [], "time": []},
            "test": {"iter": [], "psnr": [], "ssim": [], "lpips": [], "time": []},
            "memory": {"iter": [], "allocated_gb": [], "reserved_gb": []},
        }
        self.start_time = time.time()

    # -- metrics ----------------------------------------------------------------------------------------------
    def _metrics_device(self, pred, target):
        """(psnr, ssim, lpips) as 0-d device tensors — no host sync."""
        mse = ops.image_sqerr(pred, target) / float(pred.numel())
        psnr = -10. * torch.log(mse) / np.log(10.)                                         # evaluation_utils.py:24-25
        ssim = ops.image_ssim(target, pred, data_range=1.0)                                # :33
        if self.lpips_fn is not None:
            a = pred.permute(2, 0, 1).unsqueeze(0) * 2 - 1                                 # :36-38
            b = target.permute(2, 0, 1).unsqueeze(0) * 2 - 1
            lp = self.lpips_fn(a, b).reshape(()).double()
        else:
            lp = torch.full((), float("nan"), dtype=torch.float64, device=pred.device)
        return psnr, ssim, lp

    def compute_metrics(self, pred, target):
        """evaluation_utils.py:22-43 -> {'psnr', 'ssim', 'lpips'} (python floats)."""
        if not torch.is_tensor(target):
            target = torch.as_tensor(np.asarray(target), dtype=torch.float32)
        target = target.to(pred.device).float()
        vals = torch.stack(self._metrics_device(pred, target)).cpu().tolist()
        return {"psnr": vals[0], "ssim": vals[1], "lpips": vals[2]}

    def evaluate_test_set(self, model_fn, poses, hwf, K, chunk, render_kwargs, images_gt, group=None):
        """evaluation_utils.py:45-74 -> (avg_metrics, all_preds, all_metrics).  ``model_fn`` is ``render``.
        With a process group each frame is rendered pixel-sharded and the metrics are computed on the gathered
        frame (identical on every rank)."""
        H, W, focal = hwf
        H, W = int(H), int(W)
        rows, preds = [], []
        with torch.no_grad():
            for pose, gt_img in zip(poses, images_gt):
                c2w = torch.as_tensor(pose)[:3, :4]
                if group is not None and parallel.world_size(group) > 1:
                    from .run_nerf_helpers import get_rays
                    ro, rd = get_rays(H, W, K, c2w)
                    rgb, _, _ = parallel.render_sharded(lambda h, w, **a: model_fn(h, w, K, chunk=chunk, **a), H, W,
                                                        ro, rd, group=group, gather=True, **render_kwargs)
                else:
                    rgb, depth, acc, _ = model_fn(H, W, K, chunk=chunk, c2w=c2w, **render_kwargs)
                if not torch.is_tensor(gt_img):
                    gt_img = torch.as_tensor(np.asarray(gt_img), dtype=torch.float32)
                gt_img = gt_img.to(rgb.device, non_blocking=True).float()
                rows.append(torch.stack(self._metrics_device(rgb, gt_img)))
                host = torch.empty(rgb.shape, dtype=rgb.dtype, pin_memory=True)
                host.copy_(rgb, non_blocking=True)
                preds.append(host)
        table = torch.stack(rows).cpu().numpy() if rows else np.zeros((0, 3))        # the one read-back (syncs)
        torch.cuda.synchronize()
        all_metrics = [{"psnr": float(r[0]), "ssim": float(r[1]), "lpips": float(r[2])} for r in table]
        all_preds = [p.numpy() for p in preds]
        avg_metrics = {
            "psnr": np.mean([m["psnr"] for m in all_metrics]),
            "ssim": np.mean([m["ssim"] for m in all_metrics]),
            "lpips": np.mean([m["lpips"] for m in all_metrics]),
            "std_psnr": np.std([m["psnr"] for m in all_metrics]),
            "std_ssim": np.std([m["ssim"] for m in all_metrics]),
            "std_lpips": np.std([m["lpips"] for m in all_metrics]),
        }
        return avg_metrics, all_preds, all_metrics

    # -- bookkeeping (evaluation_utils.py:76-104) ---------------------------------------------------------------
    def record_test_metrics(self, iteration, test_metrics):
        t = self.metrics_history["test"]
        t["iter"].append(iteration)
        for k in ("psnr", "ssim", "lpips"):
            t[k].append(test_metrics[k])
        t["time"].append(time.time() - self.start_time)

    def record_memory_usage(self, iteration):
        if torch.cuda.is_available():
            m = self.metrics_history["memory"]
            m["iter"].append(iteration)
            m["allocated_gb"].append(torch.cuda.memory_allocated() / 1e9)
            m["reserved_gb"].append(torch.cuda.memory_reserved() / 1e9)

    def save_metrics(self, save_path):
        with open(save_path, "wb") as f:
            pickle.dump(self.metrics_history, f)
