"""TEST INFRASTRUCTURE — the reference's training step (run_nerf.py:1007-1162 with the chair config)
assembled from the oracle functions, for the CPU baseline / `bench.py --impl reference` and for the
"reference eager path on the same B200" denominator.  Device-agnostic; never imported by the product."""
import math

import torch

from . import hashnerf_oracle as O


class OracleModel:
    def __init__(self, box_min, box_max, log2T=19, finest=512, device="cpu", seed=0, lr=0.01):
        g = torch.Generator().manual_seed(seed)
        self.device = torch.device(device)
        self.log2T, self.finest = log2T, finest
        self.box_min, self.box_max = box_min.to(device), box_max.to(device)
        self.tables = [((torch.rand(1 << log2T, 2, generator=g) * 2 - 1) * 1e-4).to(device).requires_grad_(True)
                       for _ in range(16)]
        self.res = O.level_resolutions(16, finest, device=device)

        def lin(o, i):
            b = 1.0 / math.sqrt(i)
            return ((torch.rand(o, i, generator=g) * 2 - 1) * b).to(device).requires_grad_(True)
        self.nets = [dict(s0=lin(64, 32), s1=lin(16, 64), c0=lin(64, 31), c1=lin(64, 64), c2=lin(3, 64))
                     for _ in range(2)]
        mlp_params = [p for n in self.nets for p in n.values()]
        self.opt = O.RAdamState([dict(params=mlp_params, weight_decay=1e-6), dict(params=self.tables, eps=1e-15)],
                                lr=lr, betas=(0.9, 0.99))

    def params(self):
        return self.tables + [p for n in self.nets for p in n.values()]

    def embed(self, x):
        return O.hash_embed(x, self.box_min, self.box_max, self.tables, self.res, self.log2T)

    def query(self, i):
        w = self.nets[i]
        return lambda pts, vd: O.run_network(pts, vd, self.embed, lambda x: O.nerf_small(x, w))


def train_step(model, batch_rays, target_s, near=2.0, far=6.0, chunk=1024 * 32, N_samples=64, N_importance=128,
               white_bkgd=True, sparse_loss_weight=1e-10, tv_loss_weight=1e-6):
    """One iteration as the reference runs it: render in `chunk`-ray pieces, img/sparsity/TV losses,
    backward, RAdam.  Returns the loss value (python float)."""
    rays = O.pack_rays(batch_rays[0], batch_rays[1], near, far, use_viewdirs=True)
    dev = rays.device
    rets = []
    for i in range(0, rays.shape[0], chunk):
        rb = rays[i:i + chunk]
        t_rand = torch.rand(rb.shape[0], N_samples, device=dev)
        u = torch.rand(rb.shape[0], N_importance, device=dev)
        rets.append(O.render_rays(rb, model.query(0), model.query(1), N_samples, N_importance, t_rand=t_rand, u=u,
                                  white_bkgd=white_bkgd))
    cat = lambda k: torch.cat([r[k] for r in rets], 0)
    for p in model.params():
        p.grad = None
    loss = torch.mean((cat("rgb_map") - target_s) ** 2) + torch.mean((cat("rgb0") - target_s) ** 2)
    loss = loss + sparse_loss_weight * (cat("sparsity_loss").sum() + cat("sparsity_loss0").sum())
    tv = 0.
    for l in range(16):
        res, cube = O.tv_cube_size(l, 16, model.finest)
        mv = torch.randint(0, res - cube, (3,), device=dev)
        tv = tv + O.tv_loss_level(model.tables[l], l, model.log2T, mv, 16, model.finest)
    loss = loss + tv_loss_weight * tv
    loss.backward()
    model.opt.step()
    return float(loss.detach())
