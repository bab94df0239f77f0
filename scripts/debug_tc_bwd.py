import sys, torch
sys.path.insert(0, '.')
import indoor_nerf_b200 as pn
from oracle import hashnerf_oracle as O
from oracle.fixtures import mlp_weights

def rel(a, b):
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max() / (b.abs().max() + 1e-30))

for normals in (False, True):
    for P in (128, 1000):
        w = {k: v.cuda().contiguous() for k, v in mlp_weights(7 + normals, normals).items()}
        gen = torch.Generator(device="cuda").manual_seed(P)
        S = 8
        feat = torch.randn(P, 32, device="cuda", generator=gen) * 0.3
        dirs = torch.nn.functional.normalize(torch.randn(max(P // S, 1), 3, device="cuda", generator=gen), dim=-1)
        sh = pn.ops.sh_encode(dirs).repeat_interleave(S, 0)[:P]
        wo = {k: v.clone().requires_grad_(True) for k, v in w.items()}
        fo = feat.clone().requires_grad_(True)
        ref = O.nerf_small(torch.cat([fo, sh], -1), wo)
        dout = torch.randn(P, ref.shape[1], device="cuda", generator=gen)
        (ref * dout).sum().backward()
        for mode in ("fp32", "bf16"):
            dfeat, _, dw = pn.ops.mlp_bwd(w, feat, dout, dirs=dirs, samples_per_ray=S, mode=mode)
            print("normals=%s P=%d mode=%s dfeat L2/max rel: %.2e %.2e" % ((normals, P, mode) + rel(dfeat, fo.grad)))
            for k in w:
                print("     d%-4s %.2e %.2e" % ((k,) + rel(dw[k], wo[k].grad)))
        # isolate: only one output channel has a cotangent
        for ch in range(ref.shape[1]):
            d1 = torch.zeros_like(dout); d1[:, ch] = dout[:, ch]
            fo2 = feat.clone().requires_grad_(True)
            r2 = O.nerf_small(torch.cat([fo2, sh], -1), w)
            (r2 * d1).sum().backward()
            dfeat, _, dw = pn.ops.mlp_bwd(w, feat, d1, dirs=dirs, samples_per_ray=S, mode="bf16")
            print("   channel %d only: dfeat rel %.2e %.2e" % ((ch,) + rel(dfeat, fo2.grad)))
