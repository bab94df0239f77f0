"""Times the small per-ray kernels of a training step (compositing forward / backward, sample_pdf, sort-merge) at the
chair workload's shapes, CUDA events around `iters` back-to-back launches after a warm-up.  Prints one JSON line.
Usage: python scripts/bench_small_kernels.py [--rays 65536] [--iters 20]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import importlib
import torch
pn = importlib.import_module("indoor-nerf_b200")
ops = pn.ops


def timed(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=65536)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    N = a.rays
    gen = torch.Generator(device="cuda").manual_seed(0)
    out = {"rays": N, "lib": os.environ.get("POCKETNERF_LIB", "default")}
    rays_d = torch.randn(N, 3, device="cuda", generator=gen)
    for S in (64, 192):
        raw = torch.randn(N, S, 4, device="cuda", generator=gen, requires_grad=True)
        z = torch.sort(2 + 4 * torch.rand(N, S, device="cuda", generator=gen), -1)[0]
        res = ops.CompositeFn.apply(raw, z, rays_d, None, True)
        loss = lambda r: (r[0] ** 2).mean() + 1e-3 * r[5].mean()
        out["composite_fwd_%d_ms" % S] = timed(lambda: ops.CompositeFn.apply(raw.detach(), z, rays_d, None, True), a.iters)
        l = loss(res)
        def bwd():
            raw.grad = None
            l.backward(retain_graph=True)
        out["composite_bwd_%d_ms" % S] = timed(bwd, a.iters)      # includes the two tiny torch loss kernels
    z = torch.sort(2 + 4 * torch.rand(N, 64, device="cuda", generator=gen), -1)[0]
    w = torch.rand(N, 64, device="cuda", generator=gen) ** 6
    u = torch.rand(N, 128, device="cuda", generator=gen)
    mid = .5 * (z[:, 1:] + z[:, :-1])
    out["sample_pdf_ms"] = timed(lambda: ops.sample_pdf(mid, w[:, 1:-1], u), a.iters)
    s = ops.sample_pdf(mid, w[:, 1:-1], u)
    out["sort_merge_ms"] = timed(lambda: ops.sort_merge(z, s), a.iters)
    ss = torch.sort(s, -1)[0]
    out["sort_merge_sorted_ms"] = timed(lambda: ops.sort_merge(z, ss), a.iters)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
