"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the LIVE, unmodified reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Each fixture stores seeded inputs and the outputs the reference produced for them on this
container's CPU (torch version recorded inside the file).  Large inputs (hash tables) are not
stored: they are produced by ``synthetic_tables`` below, an integer formula that the tests
re-evaluate.  The committed fixtures are what pins ``oracle/hashnerf_oracle.py`` (and through
it the CUDA path) to the reference.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


from oracle.fixtures import synthetic_tables, synthetic_points, mlp_weights  # noqa: E402


def load_ref_mlp(ref, w, **kw):
    net = ref_shim.make_nerf_small(ref, predict_normals=("n0w" in w), **kw)
    with torch.no_grad():
        net.sigma_net[0].weight.copy_(w["s0"]); net.sigma_net[1].weight.copy_(w["s1"])
        net.color_net[0].weight.copy_(w["c0"]); net.color_net[1].weight.copy_(w["c1"])
        net.color_net[2].weight.copy_(w["c2"])
        if "n0w" in w:
            net.normal_net[0].weight.copy_(w["n0w"]); net.normal_net[0].bias.copy_(w["n0b"])
            net.normal_net[2].weight.copy_(w["n2w"]); net.normal_net[2].bias.copy_(w["n2b"])
    return net


def load_ref_embedder(ref, box, log2T, finest, tables, **kw):
    emb = ref.hash_encoding.HashEmbedder((torch.tensor(box[0]), torch.tensor(box[1])),
                                         log2_hashmap_size=log2T, finest_resolution=finest, **kw)
    with torch.no_grad():
        for l in range(16):
            emb.embeddings[l].weight.copy_(torch.from_numpy(tables[l]))
    return emb


def npy(d):
    return {k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    arrs["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **npy(arrs))
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in arrs.items()})


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)          # fixed reduction order for the GEMM goldens
    ref = ref_shim.load()
    U, HE, H = ref.utils, ref.hash_encoding, ref.run_nerf_helpers
    RN = ref.run_nerf

    # ---- hash primitives -------------------------------------------------------------------
    corners = torch.tensor([[0, 0, 0], [1, 2, 3], [511, 511, 511], [1023, 1000, 7], [17, 1024, 1024],
                            [1025, 1025, 1025], [3, 0, 1024]])
    res512 = [float(torch.floor(HE.HashEmbedder((torch.zeros(3), torch.ones(3)), finest_resolution=512).base_resolution
                                * HE.HashEmbedder((torch.zeros(3), torch.ones(3)), finest_resolution=512).b ** i)) for i in range(16)]
    e1024 = HE.HashEmbedder((torch.zeros(3), torch.ones(3)), finest_resolution=1024)
    res1024 = [float(torch.floor(e1024.base_resolution * e1024.b ** i)) for i in range(16)]
    save("hash_primitives", corners=corners, h19=U.hash(corners, 19), h22=U.hash(corners, 22),
         h12=U.hash(corners, 12), res512=np.array(res512, np.float32), res1024=np.array(res1024, np.float32))

    # ---- voxel vertices + embedding, several boxes -------------------------------------------
    for tag, box, log2T, finest, n in (("a", ([-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]), 19, 512, 257),
                                        ("b", ([-4.2, -3.9, -1.1], [4.4, 4.0, 5.3]), 14, 1024, 300),
                                        ("c", ([-1.6, -1.2, -1.0001], [1.7, 1.1, 1.0001]), 22, 512, 129)):
        x = torch.from_numpy(synthetic_points(n, box[0], box[1], seed=len(tag) + log2T))
        tables = synthetic_tables(16, log2T, salt=log2T)
        emb = load_ref_embedder(ref, box, log2T, finest, tables).eval()
        for p in emb.parameters():
            p.requires_grad_(True)
        feat, keep = emb(x)
        idx = []
        for l in range(16):
            resl = torch.floor(emb.base_resolution * emb.b ** l)
            _, _, h, _ = U.get_voxel_vertices(x, emb.bounding_box, resl, log2T)
            idx.append(h)
        idx = torch.stack(idx, 1)
        g = torch.from_numpy(np.random.RandomState(7).randn(*feat.shape).astype(np.float32))
        (feat * g).sum().backward()
        # table gradients are dense [T,2]; store per-level |g| sums and the touched rows of a probe level
        gsum = torch.stack([emb.embeddings[l].weight.grad.abs().sum() for l in range(16)])
        gsig = torch.stack([emb.embeddings[l].weight.grad.sum(0) for l in range(16)])
        probe = {}
        for l in (0, 7, 15):
            gr = emb.embeddings[l].weight.grad
            rows = torch.nonzero(gr.abs().sum(1)).flatten()[:512]
            probe["grad_rows_%d" % l] = rows
            probe["grad_vals_%d" % l] = gr[rows]
        save("hash_embed_" + tag, x=x, box_min=np.array(box[0], np.float32), box_max=np.array(box[1], np.float32),
             log2T=log2T, finest=finest, salt=log2T, feat=feat, keep=keep, idx=idx, dfeat=g,
             grad_abs_sum=gsum, grad_sum=gsig, **probe)

    # ---- quantised embedding (A-CAQ fake-quant on the gathered corners) -----------------------
    box = ([-1.5, -1.5, -1.5], [1.5, 1.5, 1.5])
    x = torch.from_numpy(synthetic_points(200, box[0], box[1], seed=3))
    tables = synthetic_tables(16, 15, salt=5)
    for mode in ("train", "eval"):
        emb = load_ref_embedder(ref, box, 15, 512, tables, use_quantization=True, quantization_bits=8)
        rs = np.random.RandomState(11)
        with torch.no_grad():
            for l, q in enumerate(emb.quantizers):
                q.soft_bits.fill_(float(rs.uniform(3.2, 11.7)))
        emb.current_step = 10_000        # past warm-up (hash_encoding.py:98)
        emb.train()
        feat_cal, _ = emb(x)             # first training call calibrates every quantizer
        qstate = {}
        for l, q in enumerate(emb.quantizers):
            qstate["q%d" % l] = torch.stack([q.soft_bits.data, q.range_scale.data, q.v_max.data,
                                             q.running_min, q.running_max])
        if mode == "eval":
            emb.eval()
        feat, keep = emb(x)
        save("hash_embed_quant_" + mode, x=x, box_min=np.array(box[0], np.float32),
             box_max=np.array(box[1], np.float32), log2T=15, finest=512, salt=5, feat=feat, keep=keep,
             feat_first_call=feat_cal, **qstate)

    # ---- SH ------------------------------------------------------------------------------------
    d = torch.from_numpy(np.random.RandomState(1).randn(64, 3).astype(np.float32))
    d = d / d.norm(dim=-1, keepdim=True)
    d[0] = torch.tensor([0.6, 0.0, 0.8])
    save("sh4", dirs=d, out=HE.SHEncoder()(d))

    # ---- NeRFSmall -------------------------------------------------------------------------------
    for tag, normals in (("plain", False), ("normals", True)):
        w = mlp_weights(5 + normals, normals)
        net = load_ref_mlp(ref, w)
        rs = np.random.RandomState(2)
        xin = torch.from_numpy(np.concatenate([rs.randn(300, 32).astype(np.float32) * 0.3,
                                               HE.SHEncoder()(d[rs.randint(0, 64, 300)]).numpy()], 1))
        xin.requires_grad_(True)
        out = net(xin)
        gout = torch.from_numpy(rs.randn(*out.shape).astype(np.float32))
        (out * gout).sum().backward()
        grads = {"g_" + k: p.grad for k, p in zip(
            ["s0", "s1", "c0", "c1", "c2"] + (["n0w", "n0b", "n2w", "n2b"] if normals else []),
            list(net.sigma_net.parameters()) + list(net.color_net.parameters())
            + (list(net.normal_net.parameters()) if normals else []))}
        save("nerf_small_" + tag, x=xin, out=out, gout=gout, gx=xin.grad,
             **{"w_" + k: v for k, v in w.items()}, **grads)

    # quantised NeRFSmall (weight fake-quant on layer 0, activation fake-quant after ReLU)
    w = mlp_weights(9)
    net = load_ref_mlp(ref, w, use_quantization=True, quantization_bits=8).train()
    xin = torch.from_numpy(np.random.RandomState(4).randn(128, 48).astype(np.float32) * 0.3)
    out_q = net(xin)                      # calibrates
    aq, wq = net.sigma_act_quantizers[0], net.sigma_weight_quantizer
    save("nerf_small_quant", x=xin, out=out_q, **{"w_" + k: v for k, v in w.items()},
         act_q=torch.stack([aq.soft_bits.data, aq.range_scale.data, aq.v_max.data]),
         w_q=torch.stack([wq.soft_bits.data, wq.range_scale.data]), out_eval=net.eval()(xin))

    # ---- raw2outputs ----------------------------------------------------------------------------
    RN.torch.set_default_dtype(torch.float32)
    raw_k = torch.tensor([[[0, 1, -1, .5], [2, -2, 0, 3], [.5, .5, .5, -1], [1, 0, 0, 10]]])
    outs = RN.raw2outputs(raw_k, torch.tensor([[2, 3, 4.5, 6]]), torch.tensor([[0, .6, .8]]), 0, True)
    save("raw2outputs_known", raw=raw_k, z=np.array([[2, 3, 4.5, 6]], np.float32), d=np.array([[0, .6, .8]], np.float32),
         rgb=outs[0], disp=outs[1], acc=outs[2], weights=outs[3], depth=outs[4], sparsity=outs[5])
    for tag, S, C, white in (("s64", 64, 4, True), ("s192n", 192, 7, False), ("s128", 128, 4, False)):
        rs = np.random.RandomState(S)
        N = 37
        raw = torch.from_numpy((rs.randn(N, S, C) * np.array([1, 1, 1, 4, 1, 1, 1][:C])).astype(np.float32))
        raw[3, :, 3] = -1.0                          # an empty ray: all weights 0, depth NaN
        raw[4, 5:, 3] = 50.0                         # an opaque ray
        z = torch.from_numpy(np.sort(2 + 4 * rs.rand(N, S), -1).astype(np.float32))
        dd = torch.from_numpy(rs.randn(N, 3).astype(np.float32))
        noise = torch.from_numpy(rs.randn(N, S).astype(np.float32)) if tag == "s128" else None
        raw.requires_grad_(True)
        # the reference draws the noise itself; feed ours by adding it to sigma beforehand is not
        # the same op order (sigma + noise happens inside) -> monkeypatch torch.randn for this call
        if noise is not None:
            real = torch.randn
            RN.torch.randn = lambda *a, **k: noise
            outs = RN.raw2outputs(raw, z, dd, 1.0, white, predict_normals=(C == 7))
            RN.torch.randn = real
        else:
            outs = RN.raw2outputs(raw, z, dd, 0, white, predict_normals=(C == 7))
        rgb, disp, acc, wts, depth, sp = outs[:6]
        cot = dict(rgb=rs.randn(N, 3), depth=rs.randn(N), acc=rs.randn(N), sp=rs.randn(N), w=rs.randn(N, S), disp=rs.randn(N))
        cot = {k: torch.from_numpy(v.astype(np.float32)) for k, v in cot.items()}
        ok = torch.isfinite(depth)
        loss = (rgb * cot["rgb"]).sum() + (depth[ok] * cot["depth"][ok]).sum() + (acc * cot["acc"]).sum() \
            + (sp * cot["sp"]).sum() + (wts * cot["w"]).sum() + (disp[ok] * cot["disp"][ok]).sum()
        extra = {}
        if C == 7:
            cn = torch.from_numpy(rs.randn(N, 3).astype(np.float32))
            loss = loss + (outs[6] * cn).sum()
            extra = dict(normal=outs[6], cot_normal=cn)
        loss.backward()
        save("raw2outputs_" + tag, raw=raw, z=z, d=dd, white=white, noise=(noise if noise is not None else np.zeros(0)),
             rgb=rgb, disp=disp, acc=acc, weights=wts, depth=depth, sparsity=sp, graw=raw.grad,
             **{"cot_" + k: v for k, v in cot.items()}, **extra)

    # ---- sample_pdf ------------------------------------------------------------------------------
    s = H.sample_pdf(torch.tensor([[2., 3, 4, 5]]), torch.tensor([[.1, .7, .2]]), 5, det=True)
    save("sample_pdf_known", out=s)
    rs = np.random.RandomState(8)
    N = 41
    z = np.sort(2 + 4 * rs.rand(N, 64), -1).astype(np.float32)
    bins = torch.from_numpy(.5 * (z[:, 1:] + z[:, :-1]))
    wts = rs.rand(N, 62).astype(np.float32) ** 4
    wts[2] = 0.0                                     # flat pdf
    wts[3, :] = 0.0; wts[3, 30] = 1.0                # single spike -> repeated cdf values
    wts = torch.from_numpy(wts)
    det = H.sample_pdf(bins, wts, 128, det=True)
    u = torch.from_numpy(rs.rand(N, 128).astype(np.float32))
    u[0, 0], u[0, 1] = 0.0, 0.99999994
    real = torch.rand
    H.torch.rand = lambda *a, **k: u
    rnd = H.sample_pdf(bins, wts, 128, det=False)
    H.torch.rand = real
    save("sample_pdf", bins=bins, weights=wts, det=det, u=u, rnd=rnd,
         merged=torch.sort(torch.cat([torch.from_numpy(z), rnd], -1), -1)[0], z=z)

    # ---- rays ------------------------------------------------------------------------------------
    Hh, Ww, focal = 20, 30, 27.5
    K = np.array([[focal, 0, 0.5 * Ww], [0, focal, 0.5 * Hh], [0, 0, 1]])
    c2w = torch.tensor([[0.8, -0.36, 0.48, 1.0], [0.6, 0.48, -0.64, -2.0], [0.0, 0.8, 0.6, 3.5]])
    ro, rd = H.get_rays(Hh, Ww, K, c2w)
    no, nd = H.ndc_rays(Hh, Ww, focal, 1., ro - torch.tensor([0, 0, 10.0]), rd)
    save("rays", H=Hh, W=Ww, K=K, c2w=c2w, rays_o=ro, rays_d=rd, ndc_o=no, ndc_d=nd)

    # ---- full render_rays through the reference modules ------------------------------------------
    for tag, normals, white, perturb, std, S_imp in (("blender", False, True, 1.0, 0.0, 128),
                                                       ("det", False, True, 0.0, 0.0, 128),
                                                       ("normals_noise", True, False, 1.0, 1.0, 64)):
        box = ([-3.1, -3.3, -2.9], [3.0, 3.2, 3.4])
        log2T = 14
        tables = synthetic_tables(16, log2T, amp=0.3, salt=21)      # "warm" magnitude so that sigma > 0 happens
        emb = load_ref_embedder(ref, box, log2T, 512, tables).eval()
        w0, w1 = mlp_weights(31, normals), mlp_weights(32, normals)
        net0, net1 = load_ref_mlp(ref, w0), load_ref_mlp(ref, w1)
        sh = HE.SHEncoder()
        rs = np.random.RandomState(5)
        N = 24
        o = torch.from_numpy((rs.randn(N, 3) * 0.2 + np.array([0, 0, 4.0])).astype(np.float32))
        tgt = rs.randn(N, 3) * 0.8
        dvec = torch.from_numpy((tgt - o.numpy()).astype(np.float32))
        dvec = dvec / dvec.norm(dim=-1, keepdim=True) * torch.from_numpy(rs.uniform(0.9, 1.1, (N, 1)).astype(np.float32))
        vd = dvec / dvec.norm(dim=-1, keepdim=True)
        rays = torch.cat([o, dvec, 2.0 * torch.ones(N, 1), 6.0 * torch.ones(N, 1), vd], -1)
        t_rand = torch.from_numpy(rs.rand(N, 64).astype(np.float32))
        u = torch.from_numpy(rs.rand(N, S_imp).astype(np.float32))
        n0 = torch.from_numpy(rs.randn(N, 64).astype(np.float32))
        n1 = torch.from_numpy(rs.randn(N, 64 + S_imp).astype(np.float32))
        rand_q, randn_q = [t_rand, u], [n0, n1]
        real_rand, real_randn = torch.rand, torch.randn
        RN.torch.rand = lambda *a, **k: rand_q.pop(0)
        RN.torch.randn = lambda *a, **k: randn_q.pop(0)
        H.torch.rand = RN.torch.rand
        query = lambda inputs, viewdirs, fn: RN.run_network(inputs, viewdirs, fn, embed_fn=emb, embeddirs_fn=sh)
        for p in list(emb.parameters()) + list(net0.parameters()) + list(net1.parameters()):
            p.requires_grad_(True); p.grad = None
        # record the depths raw2outputs is called with (second call = the sort-merged fine depths), so that a test can
        # feed the fine pass the reference's exact samples and hold it to the fp32 tolerance
        z_seen = []
        real_r2o = RN.raw2outputs

        def spy_r2o(raw, z_vals, *a, **k):
            z_seen.append(z_vals.detach().clone())
            return real_r2o(raw, z_vals, *a, **k)
        RN.raw2outputs = spy_r2o
        ret = RN.render_rays(rays, net0, query, 64, embed_fn=emb, retraw=True, perturb=perturb,
                             N_importance=S_imp, network_fine=net1, white_bkgd=white, raw_noise_std=std,
                             predict_normals=normals)
        RN.raw2outputs = real_r2o
        RN.torch.rand, RN.torch.randn = real_rand, real_randn
        target = torch.from_numpy(rs.rand(N, 3).astype(np.float32))
        loss = ((ret["rgb_map"] - target) ** 2).mean() + ((ret["rgb0"] - target) ** 2).mean() \
            + 1e-3 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())
        loss.backward()
        g = {}
        for nm, net in (("m0", net0), ("m1", net1)):
            for k, p in zip(["s0", "s1", "c0", "c1", "c2", "n0w", "n0b", "n2w", "n2b"],
                            list(net.sigma_net.parameters()) + list(net.color_net.parameters())
                            + (list(net.normal_net.parameters()) if normals else [])):
                g["g_%s_%s" % (nm, k)] = p.grad if p.grad is not None else torch.zeros_like(p)
        g["g_table_abs_sum"] = torch.stack([emb.embeddings[l].weight.grad.abs().sum() for l in range(16)])
        g["g_table_sum"] = torch.stack([emb.embeddings[l].weight.grad.sum(0) for l in range(16)])
        keys = ["rgb_map", "depth_map", "acc_map", "sparsity_loss", "rgb0", "depth0", "acc0",
                "sparsity_loss0", "z_std", "raw", "pts"] + (["normal_map", "normal0"] if normals else [])
        save("render_rays_" + tag, rays=rays, t_rand=t_rand, u=u, noise0=n0, noise1=n1, target=target,
             box_min=np.array(box[0], np.float32), box_max=np.array(box[1], np.float32), log2T=log2T,
             salt=21, amp=0.3, perturb=perturb, raw_noise_std=std, white=white, N_importance=S_imp, loss=loss,
             z_coarse=z_seen[0], z_fine=z_seen[1],
             **{"w0_" + k: v for k, v in w0.items()}, **{"w1_" + k: v for k, v in w1.items()},
             **{k: ret[k] for k in keys}, **g)

    # ---- TV loss + RAdam (train-step surroundings of the CPU baseline) ------------------------------
    table = torch.from_numpy(synthetic_tables(1, 19, salt=2)[0]).requires_grad_(True)
    tv = {}
    for level in (0, 5, 15):
        torch.manual_seed(100 + level)
        emb_mod = torch.nn.Embedding(1 << 19, 2, _weight=table)
        val = ref.loss.total_variation_loss(emb_mod, 16, 512, level, 19, n_levels=16)
        from oracle.hashnerf_oracle import tv_cube_size
        res, cube = tv_cube_size(level)
        torch.manual_seed(100 + level)
        mv = torch.randint(0, res - cube, (3,))
        tv["tv_%d" % level] = val.detach()
        tv["mv_%d" % level] = mv
    save("tv_loss", salt=2, **tv)

    rs = np.random.RandomState(12)
    p0 = [torch.from_numpy(rs.randn(5, 7).astype(np.float32)).requires_grad_(True),
          torch.from_numpy(rs.randn(33).astype(np.float32) * 1e-4).requires_grad_(True)]
    init = [p.detach().clone() for p in p0]
    import warnings
    warnings.simplefilter("ignore")
    opt = ref.radam.RAdam([{"params": [p0[0]], "weight_decay": 1e-6}, {"params": [p0[1]], "eps": 1e-15}],
                          lr=5e-4, betas=(0.9, 0.99))
    grads, hist = [], []
    for it in range(8):
        gs = [torch.from_numpy(rs.randn(*p.shape).astype(np.float32)) * (1e-3 if i else 1.0) for i, p in enumerate(p0)]
        for p, g_ in zip(p0, gs):
            p.grad = g_.clone()
        opt.step()
        grads.append(gs)
        hist.append([p.detach().clone() for p in p0])
    save("radam", init0=init[0], init1=init[1],
         **{"g%d_%d" % (it, i): grads[it][i] for it in range(8) for i in range(2)},
         **{"p%d_%d" % (it, i): hist[it][i] for it in range(8) for i in range(2)})


if __name__ == "__main__":
    main()
