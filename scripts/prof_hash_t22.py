"""The stand-alone hash-encode kernels (K1 / K1b) at log2_hashmap 22 on 12.6 M ray-ordered points: the launch an ncu
capture profiles for DRAM bytes / L2 hit rate / sectors per request (profiles/r02_ncu_hash_t22_summary.txt)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
scene = synthetic.blender_scene(400, 400, n_views=100)
rays, _ = synthetic.ray_batch(scene, 65536, seed=5, device=dev)
z = torch.sort(2.0 + 4.0 * torch.rand(65536, 192, device=dev), -1)[0]
pts = ops.make_points(rays[0], rays[1], z).reshape(-1, 3).contiguous()
log2T = int(sys.argv[1]) if len(sys.argv) > 1 else 22
emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=log2T, finest_resolution=512).to(dev)
tables = [t.detach() for t in emb.tables()]
flat = torch.zeros(16, 1 << log2T, 2, device=dev)
dfeat = torch.randn(pts.shape[0], 32, device=dev)
with torch.no_grad():
    for _ in range(3):
        feat, keep = ops.hash_encode_fwd(emb.grid(), tables, pts)
        ops.hash_encode_bwd(emb.grid(), list(flat.unbind(0)), pts, dfeat)
torch.cuda.synchronize()
print("ok", float(feat.abs().sum()), float(flat.abs().sum()))
