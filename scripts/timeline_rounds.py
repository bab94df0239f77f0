"""Per-round latency table of the fused backward kernel (pn_field_bwd_bf16) from the clock64 marks that
pn_debug_timeline installs: where, inside one 128-point tile, do the cycles go?

    PN_FIELD_BWD=v1 [PN_DEBUG_FLAGS=1] python scripts/timeline_rounds.py [--density 0.1]

Instrumented kernel: the single-role fused backward (mlp_tc_bwd_kernel<SRC_TILE>, PN_FIELD_BWD=v1).  Marks per tile
(thread 0 = MMA issuer, thread 160 = plain epilogue thread), see mlp_tc.cu:
  0 tile start | 1 inputs loaded | per forward round R1..R4: issue-start, issue-end, mma-done, epilogue-done |
  18 cotangent tiles written | per backward round B1..B4: the same four | B5: issue-start, issue-end, mma-done | 38 tile end
The tables this produced in round 2 (profiles/r02_timeline_*.txt) are what the three-role kernel was designed from.
"""
import argparse
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

ROUNDS = ["R1", "R2", "R3", "R4", "B1", "B2", "B3", "B4", "B5"]
V1 = ["start", "loaded"]
for r in ROUNDS[:4]:
    V1 += [r + ".issue0", r + ".issued", r + ".done", r + ".epi"]
V1 += ["B0.tiles"]
for r in ROUNDS[4:8]:
    V1 += [r + ".issue0", r + ".issued", r + ".done", r + ".epi"]
V1 += ["B5.issue0", "B5.issued", "B5.done", "end"]
V3E = ["start", "loaded"] + [x for r in ROUNDS[:8] for x in (r + ".done", r + ".epi")] + ["B5.done"]
V3M = ["start"] + [x for r in ROUNDS for x in (r + ".ready", r + ".issued")]


def table(t, names, skip=2):
    m = len(names)
    t = t[: (len(t) // m) * m].reshape(-1, m)
    t = t[(t != 0).all(1)]
    d = np.diff(t[skip:], axis=1)
    per_tile = np.diff(t[skip:, 0])
    wrap = t[skip + 1:, 0] - t[skip:-1, -1]
    out = {names[i]: float(np.median(d[:, i])) for i in range(m - 1)}
    out[names[-1] + "->next"] = float(np.median(wrap)) if len(wrap) else None
    return float(np.median(per_tile)) if len(per_tile) else None, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--density", type=float, default=1.0)
    a = ap.parse_args()
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
    if not os.environ.get("PN_FIELD_BWD", "").startswith("v1"):
        sys.exit("set PN_FIELD_BWD=v1: only the single-role kernel carries the clock64 marks")
    v3 = False
    tiles = 14
    cap = 40 * tiles
    buf = torch.zeros(3, cap, dtype=torch.int64, device=dev)
    for it in range(3):
        for prm in list(emb.parameters()) + list(net.parameters()):
            prm.grad = None
        o = pn.run_network(pts, vd, net, emb, sh)
        dout = torch.ones_like(o)
        if a.density < 1.0:
            blk = (torch.rand(65536, 192 // 16, 1, device=dev) < a.density).float()
            dout = dout * blk.repeat_interleave(16, 1)
        if it == 2:
            buf.zero_()
            _lib.call("pn_debug_timeline", ctypes.c_void_p(buf.data_ptr()), cap)
        o.backward(dout)
        torch.cuda.synchronize()
    _lib.call("pn_debug_timeline", None, 0)
    t = buf.cpu().numpy()
    out = {"variant": {k: os.environ.get(k) for k in ("PN_FIELD_BWD", "PN_DEBUG_FLAGS")}, "density": a.density}
    rows = [("thread0", 0, V3E if v3 else V1), ("thread160", 1, V3E if v3 else V1)] + ([("mma_warp", 2, V3M)] if v3 else [])
    for name, r, names in rows:
        per_tile, tab = table(t[r], names)
        out[name] = {"cycles_per_tile": per_tile, "delta_to_next_mark": tab}
    print(json.dumps(out))
    print("variant %s density %.2f" % (out["variant"], a.density), file=sys.stderr)
    for name, _, _ in rows:
        print(" %s: cycles per tile %s" % (name, out[name]["cycles_per_tile"]), file=sys.stderr)
        for k, v in out[name]["delta_to_next_mark"].items():
            print("    %-14s %8.0f" % (k, v if v is not None else -1), file=sys.stderr)


if __name__ == "__main__":
    main()
