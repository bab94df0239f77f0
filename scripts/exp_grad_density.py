"""How dense are the cotangents the fused backward sees in a real training step?  Hooks FieldFn.backward for a few steps
of the bench workload and prints, per launch: fraction of points with a non-zero cotangent row, fraction of 32-sample
warps with at least one such point, and the mean number of non-zero rows per active warp."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import model as pmodel, ops, synthetic  # noqa: E402
from indoor_nerf_b200.trainer import Trainer  # noqa: E402

pn.set_mlp_mode("bf16")
dev = torch.device("cuda", 0)
scene = synthetic.blender_scene(400, 400, n_views=100)
a = pmodel.default_args(bounding_box=scene["bounding_box"], lrate=0.01)
torch.manual_seed(0)
kw, _, _, _, opt = pmodel.create_nerf(a, device=dev)
tr = Trainer(a, kw, opt, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"])
stats = []
orig = ops.FieldFn.backward


def hooked(ctx, dout):
    nz = (dout != 0).any(-1).reshape(-1)
    w = nz[: (nz.numel() // 32) * 32].reshape(-1, 32)
    act = w.any(-1)
    stats.append({"points": int(nz.numel()), "row_density": float(nz.float().mean()), "warp_density": float(act.float().mean()),
                  "rows_per_active_warp": float(w[act].float().sum(-1).mean()) if act.any() else 0.0})
    return orig(ctx, dout)


ops.FieldFn.backward = staticmethod(hooked)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
for i in range(steps):
    r, t = synthetic.ray_batch(scene, 65536, seed=i, device=dev)
    loss, psnr = tr.step(r, t)
    if i in (0, 1, 5, 20, steps - 1):
        print(json.dumps({"step": i, "loss": float(loss), "launches": stats[-2:]}))
