"""Per-round latency table of the fused backward kernel (pn_field_bwd_bf16) from the clock64 marks that
pn_debug_timeline installs: where, inside one 128-point tile, do the ~20-40 k cycles go?

    PN_FIELD_BWD=v1|ws [PN_DEBUG_FLAGS=1] python scripts/timeline_rounds.py

Marks per tile (thread 0 = MMA issuer, thread 160 = plain epilogue thread), see mlp_tc.cu:
  0 tile start | 1 inputs loaded | per forward round R1..R4: issue-start, issue-end, mma-done, epilogue-done |
  18 cotangent tiles written | per backward round B1..B4: the same four | B5: issue-start, issue-end, mma-done | 38 tile end
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import _lib, ops, synthetic  # noqa: E402

MARKS = 39
NAMES = ["start", "loaded"]
for r in ("R1", "R2", "R3", "R4"):
    NAMES += [r + ".issue0", r + ".issued", r + ".done", r + ".epi"]
NAMES += ["B0.tiles"]
for r in ("B1", "B2", "B3", "B4"):
    NAMES += [r + ".issue0", r + ".issued", r + ".done", r + ".epi"]
NAMES += ["B5.issue0", "B5.issued", "B5.done", "end"]
assert len(NAMES) == MARKS


def main():
    pn.set_mlp_mode("bf16")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    scene = synthetic.blender_scene(400, 400, n_views=100)
    emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=19, finest_resolution=512).to(dev)
    with torch.no_grad():
        emb.table_storage.mul_(3000.0)
    net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16).to(dev)
    sh = pn.SHEncoder()
    rays, _ = synthetic.ray_batch(scene, 65536, seed=5, device=dev)
    vd = rays[1] / rays[1].norm(dim=-1, keepdim=True)
    z = torch.sort(2.0 + 4.0 * torch.rand(65536, 192, device=dev), -1)[0]
    pts = ops.make_points(rays[0], rays[1], z)
    tiles = 12
    buf = torch.zeros(2, MARKS * tiles, dtype=torch.int64, device=dev)
    for it in range(3):
        for prm in list(emb.parameters()) + list(net.parameters()):
            prm.grad = None
        o = pn.run_network(pts, vd, net, emb, sh)
        if it == 2:
            buf.zero_()
            _lib.call("pn_debug_timeline", ctypes.c_void_p(buf.data_ptr()), MARKS * tiles)
        o.backward(torch.ones_like(o))
        torch.cuda.synchronize()
    _lib.call("pn_debug_timeline", None, 0)
    t = buf.cpu().numpy().reshape(2, tiles, MARKS)
    out = {"variant": {k: os.environ.get(k) for k in ("PN_FIELD_BWD", "PN_DEBUG_FLAGS")}}
    for th, name in ((0, "thread0"), (1, "thread160")):
        d = np.diff(t[th, 2:], axis=1)                          # skip the first two tiles (cold)
        tile_total = t[th, 3:, 0] - t[th, 2:-1, 0]
        out[name] = {"cycles_per_tile": float(np.median(tile_total)),
                     "delta_to_next_mark": {NAMES[i]: float(np.median(d[:, i])) for i in range(MARKS - 1)}}
    print(json.dumps(out))
    th0 = out["thread0"]["delta_to_next_mark"]
    print("thread 0, cycles per tile %.0f" % out["thread0"]["cycles_per_tile"], file=sys.stderr)
    for k, v in th0.items():
        print("  %-12s %8.0f   (thread160 %8.0f)" % (k, v, out["thread160"]["delta_to_next_mark"][k]), file=sys.stderr)


if __name__ == "__main__":
    main()
