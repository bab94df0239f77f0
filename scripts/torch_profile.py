"""Kernel / CPU-op breakdown of the training step with torch.profiler (cheap; ncu is for the per-kernel detail)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import model as pmodel, synthetic  # noqa: E402
from indoor_nerf_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
scene = synthetic.blender_scene(400, 400, n_views=100)
a = pmodel.default_args(bounding_box=scene["bounding_box"], lrate=0.01)
kw, _, _, _, opt = pmodel.create_nerf(a, device=dev)
tr = Trainer(a, kw, opt, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"])
rays, target = synthetic.ray_batch(scene, 65536, seed=1, device=dev)
for _ in range(4):
    tr.step(rays, target)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    tr.step(rays, target)
t_cpu = (time.perf_counter() - t0) / 5
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 5
print("mode %s: CPU issue time per step %.2f ms, wall per step %.2f ms" % (pn.get_mlp_mode(), t_cpu * 1e3, t_all * 1e3))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CPU, torch.profiler.ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.step(rays, target)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
