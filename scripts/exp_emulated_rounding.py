"""Which forward rounding of a bf16 NeRFSmall costs training quality?  The oracle's training step (eager torch) with the
MLP forward rounded to bf16 at chosen places (straight-through backward): weights only, activations/inputs only, both.
Same scene / batches / steps as tests/test_gpu_training.py."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hashnerf_oracle as O  # noqa: E402
from oracle.train_step import OracleModel, train_step  # noqa: E402
from indoor_nerf_b200 import synthetic  # noqa: E402
from tests.test_gpu_training import _scene, N_RAYS, LOG2T  # noqa: E402

cfg = sys.argv[1]                       # none | w | a | wa | a16 (fp16 activations) | x (inputs only) | h (hidden only)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 600
dev = torch.device("cuda", 0)


def rnd(t, dtype=torch.bfloat16):
    return t + (t.to(dtype).float() - t).detach()


def nerf_small_rounded(x, w):
    rw = (lambda t: rnd(t)) if "w" in cfg else (lambda t: t)
    adt = torch.float16 if "16" in cfg else torch.bfloat16
    rx = (lambda t: rnd(t, adt)) if ("a" in cfg or "x" in cfg) else (lambda t: t)
    rh = (lambda t: rnd(t, adt)) if ("a" in cfg or "h" in cfg) else (lambda t: t)
    feat, views = torch.split(x, [32, 16], dim=-1)
    h = rh(F.relu(F.linear(rx(feat), rw(w["s0"]))))
    h = F.linear(h, rw(w["s1"]))
    sigma, geo = h[..., 0], h[..., 1:]
    c = torch.cat([rx(views), rh(geo)], dim=-1)
    c = rh(F.relu(F.linear(c, rw(w["c0"]))))
    c = rh(F.relu(F.linear(c, rw(w["c1"]))))
    color = F.linear(c, rw(w["c2"]))
    return torch.cat([color, sigma.unsqueeze(-1)], -1)


scene = _scene()
om = OracleModel(*scene["bounding_box"], log2T=LOG2T, finest=512, device=dev, lr=0.01)
if cfg != "none":
    om.query = lambda i: (lambda pts, vd: O.run_network(pts, vd, om.embed, lambda x: nerf_small_rounded(x, om.nets[i])))
torch.manual_seed(1)
ps = []
for i in range(steps):
    r, t = synthetic.ray_batch(scene, N_RAYS, seed=i, device=dev)
    # train_step returns the total loss; PSNR from a no-grad re-render would double the cost: use the loss's image term proxy
    loss = train_step(om, r, t, near=scene["near"], far=scene["far"], chunk=N_RAYS)
    ps.append(loss)
ps = np.array(ps)
w = [(20, 40), (40, 80), (80, 150), (150, 250), (250, 400), (400, 500), (500, 600)]
print(cfg, "mean loss per window:", " ".join("%.5f" % float(np.mean(ps[a:b])) for a, b in w if b <= steps))
