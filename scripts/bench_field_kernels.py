"""Kernel-only timing of the two fused field kernels of the bf16 training step (pn_field_fwd_bf16 / pn_field_bwd_bf16)
on the step's own point sets: 65536 rays x 192 samples (fine pass) and x 64 (coarse pass), ray-ordered, T = 2^19 (or
--log2T 22).  CUDA events around back-to-back launches, 3 warm-up + 10 timed.  The kernel variant is chosen by the
environment (read once per process): PN_FIELD_BWD=v1 (single-role kernel) or unset (three-role kernel); PN_DEBUG_FLAGS=1
skips the scatter work, 2 the gather work (measurement only).

    python scripts/bench_field_kernels.py [--log2T 19] [--normals]            # one JSON line
    python scripts/bench_field_kernels.py --sweep                             # every variant, one subprocess each
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (PN_FIELD_FWD, PN_FWD_GW, PN_FIELD_BWD, PN_DEBUG_FLAGS, extra args)
VARIANTS = [("v1", "8", "v1", "0", []), ("v1", "8", "v4", "0", []),
            ("v1", "8", "v1", "0", ["--density", "0.1"]), ("v1", "8", "v4", "0", ["--density", "0.1"]),
            ("v1", "8", "v4", "1", [])]


def sweep(extra):
    for fwd, gw, bwd, dbg, more in VARIANTS:
        env = dict(os.environ, PN_FIELD_FWD=fwd, PN_FWD_GW=gw, PN_FIELD_BWD=bwd, PN_DEBUG_FLAGS=dbg)
        r = subprocess.run(["timeout", "300", sys.executable, os.path.abspath(__file__)] + list(extra) + more, env=env,
                           capture_output=True, text=True)
        line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else json.dumps({"error": r.stderr[-600:], "rc": r.returncode})
        print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log2T", type=int, default=19)
    ap.add_argument("--normals", action="store_true")
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--sweep", action="store_true")
    ap.add_argument("--density", type=float, default=1.0, help="fraction of 16-sample blocks with a non-zero cotangent")
    ap.add_argument("--iters", type=int, default=13)
    ap.add_argument("--only", type=int, default=0, help="run only this samples-per-ray value (192 or 64)")
    a = ap.parse_args()
    if a.sweep:
        return sweep([x for x in sys.argv[1:] if x != "--sweep"])
    import torch
    import indoor_nerf_b200 as pn
    from indoor_nerf_b200 import ops, synthetic
    pn.set_mlp_mode("bf16")
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    scene = synthetic.blender_scene(400, 400, n_views=100)
    emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=a.log2T, finest_resolution=512).to(dev)
    with torch.no_grad():
        emb.table_storage.mul_(3000.0)
    net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16,
                       predict_normals=a.normals).to(dev)
    sh = pn.SHEncoder()
    rays, _ = synthetic.ray_batch(scene, a.rays, seed=5, device=dev)
    vd = rays[1] / rays[1].norm(dim=-1, keepdim=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {"variant": {k: os.environ.get(k) for k in ("PN_FIELD_FWD", "PN_FWD_GW", "PN_FIELD_BWD", "PN_DEBUG_FLAGS")},
           "log2T": a.log2T, "normals": a.normals, "density": a.density}
    for S in ((a.only,) if a.only else (192, 64)):
        z = torch.sort(2.0 + 4.0 * torch.rand(a.rays, S, device=dev), -1)[0]
        pts = ops.make_points(rays[0], rays[1], z)
        fw, bw = [], []
        chk = None
        for it in range(a.iters):
            for prm in list(emb.parameters()) + list(net.parameters()):
                prm.grad = None
            ops.table_grad_buffer(list(emb.tables()))        # the zeroing of the flat gradient stays outside the timing
            torch.cuda.synchronize()
            e0.record()
            o = pn.run_network(pts, vd, net, emb, sh)
            e1.record()
            torch.cuda.synchronize()
            fw.append(e0.elapsed_time(e1))
            if it == 0:
                dout = torch.ones_like(o)
                if a.density < 1.0:
                    blk = (torch.rand(a.rays, S // 16, 1, device=dev) < a.density).float()
                    dout = dout * blk.repeat_interleave(16, 1)
            torch.cuda.synchronize()
            e0.record()
            o.backward(dout)
            e1.record()
            torch.cuda.synchronize()
            bw.append(e0.elapsed_time(e1))
            if chk is None:
                g = emb.embeddings[0].weight.grad
                chk = {"out_sum": float(o.double().sum()), "out_abs": float(o.double().abs().sum()),
                       "g0_abs": float(g.double().abs().sum()), "g15_abs": float(emb.embeddings[15].weight.grad.double().abs().sum()),
                       "ds0_abs": float(net.sigma_net[0].weight.grad.double().abs().sum())}
        import numpy as np
        out["S%d" % S] = {"points": a.rays * S, "fwd_ms": float(np.median(fw[min(3, a.iters - 1):])), "bwd_ms": float(np.median(bw[min(3, a.iters - 1):])), "check": chk}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
