"""K1 micro-benchmark grid of SURVEY.md section 8d: hash-encode forward / backward over
P in {2^18, 2^22, 2^24} points x {uniform-random in the box, ray-ordered samples} x T in {2^19, 2^22} x finest in {512, 1024}.
Prints one JSON object per configuration (ms per launch, algorithmic GB/s at 1164 B/point, fraction of the measured HBM
peak).  One GPU, CUDA events, 3 warm-up + 5 timed launches; the table set (64 MiB / 512 MiB) is re-used across P.

    python scripts/bench_hash_micro.py > gpurun_out/hash_micro.jsonl
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import ops, synthetic  # noqa: E402

BYTES = 12 + 16 * 8 * 8 + 128
dev = torch.device("cuda", 0)
peak = 6534.1
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.isfile(p):
    peak = float(json.load(open(p))["hbm_gbs"])
scene = synthetic.blender_scene(400, 400, n_views=100)
bmin, bmax = scene["bounding_box"]


def timeit(fn, reps=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def points(P, order):
    if order == "uniform":
        return (bmin.to(dev) + (bmax - bmin).to(dev) * torch.rand(P, 3, device=dev)).contiguous()
    S = 192
    n = (P + S - 1) // S
    rays, _ = synthetic.ray_batch(scene, n, seed=3, device=dev)
    z = torch.sort(2.0 + 4.0 * torch.rand(n, S, device=dev), -1)[0]
    return ops.make_points(rays[0], rays[1], z).reshape(-1, 3)[:P].contiguous()


for log2T in (19, 22):
    for finest in (512, 1024):
        emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=log2T, finest_resolution=finest).to(dev)
        tables = [t.detach() for t in emb.tables()]
        flat = torch.zeros(16, 1 << log2T, 2, device=dev)
        grads = list(flat.unbind(0))
        for P in (1 << 18, 1 << 22, 1 << 24):
            for order in ("uniform", "ray_ordered"):
                x = points(P, order)
                dfeat = torch.randn(P, 32, device=dev)
                tf = timeit(lambda: ops.hash_encode_fwd(emb.grid(), tables, x))
                tb = timeit(lambda: ops.hash_encode_bwd(emb.grid(), grads, x, dfeat))
                row = {"log2T": log2T, "finest": finest, "P": P, "order": order,
                       "fwd_ms": tf, "fwd_GBs": BYTES * P / tf / 1e6, "fwd_frac_hbm": BYTES * P / tf / 1e6 / peak,
                       "bwd_ms": tb, "bwd_GBs": BYTES * P / tb / 1e6, "bwd_frac_hbm": BYTES * P / tb / 1e6 / peak}
                print(json.dumps(row), flush=True)
                del x, dfeat
        del emb, tables, flat, grads
        torch.cuda.empty_cache()
