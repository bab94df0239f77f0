"""Per-level cost of the hash-grid forward gather and backward scatter on ray-ordered points."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import ops, synthetic  # noqa: E402

dev = torch.device("cuda", 0)
scene = synthetic.blender_scene(400, 400, n_views=100)
rays, _ = synthetic.ray_batch(scene, 65536, seed=1, device=dev)
S = 192
z = torch.sort(2.0 + 4.0 * torch.rand(65536, S, device=dev), -1)[0]
pts = ops.make_points(rays[0], rays[1], z).reshape(-1, 3)
P = pts.shape[0]
emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=19, finest_resolution=512).to(dev)
res = [float(r) for r in emb.level_resolutions()]
bmin, bmax = [t.tolist() for t in scene["bounding_box"]]


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print("P = %d points (65536 rays x %d sorted samples)" % (P, S))
tot_f = tot_b = 0.0
for l in range(16):
    grid = ops.make_grid(bmin, bmax, [res[l]], 19)
    table = emb.embeddings[l].weight.detach()
    dfeat = torch.randn(P, 2, device=dev)
    dtab = torch.zeros_like(table)
    tf = timeit(lambda: ops.hash_encode_fwd(grid, [table], pts))
    tb = timeit(lambda: ops.hash_encode_bwd(grid, [dtab], pts, dfeat))
    tot_f += tf
    tot_b += tb
    print("level %2d res %4d: fwd %.3f ms  bwd %.3f ms" % (l, int(res[l]), tf, tb))
print("sum of single-level launches: fwd %.2f ms bwd %.2f ms" % (tot_f, tot_b))
grid = emb.grid()
tables = [t.detach() for t in emb.tables()]
dfeat = torch.randn(P, 32, device=dev)
flat = torch.zeros(16, 1 << 19, 2, device=dev)
print("all levels: fwd %.3f ms  bwd %.3f ms" % (timeit(lambda: ops.hash_encode_fwd(grid, tables, pts)),
                                                timeit(lambda: ops.hash_encode_bwd(grid, list(flat.unbind(0)), pts, dfeat))))
# random (unordered) points for contrast
perm = torch.randperm(P, device=dev)
pr = pts[perm].contiguous()
print("all levels, shuffled points: fwd %.3f ms  bwd %.3f ms" % (timeit(lambda: ops.hash_encode_fwd(grid, tables, pr)),
                                                                   timeit(lambda: ops.hash_encode_bwd(grid, list(flat.unbind(0)), pr, dfeat))))
