"""Host-side cost of one training step at a small per-GPU batch (strong scaling at 8 GPUs = 8192 rays per rank): wall time per
step against GPU time, and cProfile's top functions.  Usage: prof_host_step.py [rays] [--graph] (Trainer(cuda_graph=True))."""
import cProfile
import io
import os
import pstats
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import model as pmodel, synthetic  # noqa: E402
from indoor_nerf_b200.trainer import Trainer  # noqa: E402

GRAPH = "--graph" in sys.argv
argv = [x for x in sys.argv[1:] if x != "--graph"]
n = int(argv[0]) if argv else 8192
pn.set_mlp_mode("bf16")
dev = torch.device("cuda", 0)
scene = synthetic.blender_scene(400, 400, n_views=100)
a = pmodel.default_args(bounding_box=scene["bounding_box"], lrate=0.01)
torch.manual_seed(0)
kw, _, _, _, opt = pmodel.create_nerf(a, device=dev)
tr = Trainer(a, kw, opt, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"], cuda_graph=GRAPH)
pool = [synthetic.ray_batch(scene, n, seed=i, device=dev) for i in range(4)]
for i in range(12):
    tr.step(*pool[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for i in range(50):
    tr.step(*pool[i % 4])
e1.record()
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
print("graph" if GRAPH else "eager", end=" ")
print("rays %d: host issue %.3f ms/step, gpu %.3f ms/step" % (n, t_issue / 50 * 1e3, e0.elapsed_time(e1) / 50))
pr = cProfile.Profile()
pr.enable()
for i in range(50):
    tr.step(*pool[i % 4])
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print("\n".join(l[:150] for l in s.getvalue().splitlines()[:60]))
