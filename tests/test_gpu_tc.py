"""tcgen05 descriptor self-test: the K-major and MN-major readings of the shared-memory tile layout and the
M=128 / M=64 accumulator layouts in TMEM, checked on hardware against a bf16-rounded fp32 GEMM."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(cfg, A, B, N):
    from indoor_nerf_b200 import _lib
    D = torch.full((128, N), float("nan"), device="cuda")
    c = (ctypes.c_int32 * 15)(*cfg)
    _lib.call("pn_tc_selftest", c, _lib.dptr(A), _lib.dptr(B), _lib.dptr(D), _lib.stream())
    torch.cuda.synchronize()
    return D


def bf(x):
    return x.to(torch.bfloat16).float()


def test_k_major_m128():
    g = torch.Generator(device="cuda").manual_seed(0)
    A = torch.randn(128, 64, device="cuda", generator=g)
    B = torch.randn(64, 64, device="cuda", generator=g)
    D = _run([128, 64, 64, 64, 128, 64, 4, 0, 0, 128, 1024, 256, 128, 1024, 256], A, B, 64)
    want = bf(A) @ bf(B).t()
    assert torch.allclose(D, want, rtol=1e-4, atol=1e-3), float((D - want).abs().max())
    # K = 32, N = 16 (the shapes of the first / small layers)
    A = torch.randn(128, 32, device="cuda", generator=g)
    B = torch.randn(16, 32, device="cuda", generator=g)
    D = _run([128, 32, 16, 32, 128, 16, 2, 0, 0, 128, 512, 256, 128, 512, 256], A, B, 16)
    assert torch.allclose(D, bf(A) @ bf(B).t(), rtol=1e-4, atol=1e-3)


def test_mn_major_b():
    """dgrad form: D[128 x N] = A[128 x K] * Bt[K x N] with Bt stored as a [K x N] tile read MN-major."""
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(128, 64, device="cuda", generator=g)
    Bt = torch.randn(64, 32, device="cuda", generator=g)          # K = 64 rows, N = 32 cols
    rg = (32 // 8) * 128
    want = bf(A) @ bf(Bt)
    res = {}
    for name, lbo, sbo in (("lbo=rowgroup,sbo=128", rg, 128), ("lbo=128,sbo=rowgroup", 128, rg)):
        D = _run([128, 64, 64, 32, 128, 32, 4, 0, 1, 128, 1024, 256, lbo, sbo, 2 * rg], A, Bt, 32)
        res[name] = float((D - want).abs().max())
    print("MN-major B hypotheses:", res)
    assert res["lbo=rowgroup,sbo=128"] < 1e-2, res


def test_mn_major_both_m64():
    """wgrad form: D[64 x N] = At^T * Bt with At [K=128 x 64], Bt [K=128 x N] tiles, both read MN-major."""
    g = torch.Generator(device="cuda").manual_seed(2)
    At = torch.randn(128, 64, device="cuda", generator=g)
    Bt = torch.randn(128, 32, device="cuda", generator=g)
    rga, rgb = (64 // 8) * 128, (32 // 8) * 128
    want = bf(At).t() @ bf(Bt)                                       # [64, 32]
    res = {}
    for name, la, sa, lb, sb in (("lbo=rowgroup,sbo=128", rga, 128, rgb, 128), ("lbo=128,sbo=rowgroup", 128, rga, 128, rgb)):
        D = _run([128, 64, 128, 32, 64, 32, 8, 1, 1, la, sa, 2 * rga, lb, sb, 2 * rgb], At, Bt, 32)
        rows = torch.arange(64, device="cuda")
        lanes = (rows % 16) + 32 * (rows // 16)                      # M=64: row m -> lane (m%16) + 32*(m/16)
        res[name] = float((D[lanes] - want).abs().max())
        res[name + " [lanes 0..63]"] = float((D[:64] - want).abs().max())
    print("MN-major A,B (M=64) hypotheses:", res)
    assert res["lbo=rowgroup,sbo=128"] < 2e-2, res


# ---------------------------------------------------------------------------------------------------------
# NeRFSmall in the bf16 tensor-core mode
# ---------------------------------------------------------------------------------------------------------
def _rel(a, b):
    a, b = a.detach().float(), b.detach().float()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max() / (b.abs().max() + 1e-30))


def _emulated_bf16_mlp(w, feat, sh, dout, keep):
    """What the tensor-core kernels compute, restated in torch: operands rounded to bf16 exactly where the kernels
    round them (inputs, weights, hidden activations, gradient tiles), fp32 accumulation.  Forward AND backward."""
    W = {k: bf(v) for k, v in w.items() if v.dim() == 2}
    x = bf(feat)
    h1 = torch.relu(x @ W["s0"].t()); m1 = h1 > 0; h1 = bf(h1)
    h2 = h1 @ W["s1"].t()
    sigma, geo = h2[:, 0], bf(h2[:, 1:])
    cin = torch.cat([bf(sh), geo], -1)
    a1 = bf(torch.relu(cin @ W["c0"].t()))
    a2 = bf(torch.relu(a1 @ W["c1"].t()))
    rgb = a2 @ W["c2"].t()
    normals = "n0w" in w
    kept = keep.float()
    if normals:
        nh = bf(torch.relu(geo @ W["n0w"].t() + w["n0b"]))
        nraw = nh @ W["n2w"].t() + w["n2b"]
        nn = nraw.norm(dim=-1, keepdim=True).clamp_min(1e-12)
        n = nraw / nn
        out = torch.cat([rgb, sigma[:, None], n[:, :2], n[:, 2:] * kept[:, None]], -1)
        d = dout.clone(); d[:, 6] *= kept
    else:
        out = torch.cat([rgb, (sigma * kept)[:, None]], -1)
        d = dout.clone(); d[:, 3] *= kept
    # backward
    g = {}
    do = bf(d[:, :3])
    g["c2"] = do.t() @ a2
    da2 = bf((do @ W["c2"]) * (a2 > 0))
    g["c1"] = da2.t() @ a1
    da1 = bf((da2 @ W["c1"]) * (a1 > 0))
    g["c0"] = da1.t() @ cin
    dgeo = (da1 @ W["c0"])[:, 16:]
    if normals:
        dn = d[:, 4:7]
        dot = (n * dn).sum(-1, keepdim=True)
        dnr = bf((dn - n * dot) / nn)
        g["n2w"] = dnr.t() @ nh; g["n2b"] = dnr.sum(0)
        dnh = bf((dnr @ W["n2w"]) * (nh > 0))
        g["n0w"] = dnh.t() @ geo; g["n0b"] = dnh.sum(0)
        dgeo = dgeo + dnh @ W["n0w"]
    dh2 = bf(torch.cat([d[:, 3:4], dgeo], -1))
    g["s1"] = dh2.t() @ h1
    dh1 = bf((dh2 @ W["s1"]) * m1)
    g["s0"] = dh1.t() @ x
    return out, dh1 @ W["s0"], g


@pytest.mark.parametrize("normals", [False, True])
@pytest.mark.parametrize("P", [128, 1000, 40000])
def test_mlp_bf16_kernels(normals, P):
    """(1) kernel == bf16-emulating reference to fp32-accumulation accuracy (proves the kernels' logic);
    (2) forward within bf16 rounding of the fp32 oracle: five chained bf16 layers with random weights measure
        3.3e-3 in L2 norm / 4.4e-3 max (of scale); bar 5e-3 / 1e-2;
    (3) gradients vs the fp32 oracle: dominated by ReLU units whose pre-activation changes sign between a bf16 and
        an fp32 forward (a fraction f ~ 1.5e-3 of units; for the incoherent random cotangent used here that alone
        gives a relative error ~ sqrt(2f) ~ 5%, independent of the kernel) -> sanity bound only."""
    import indoor_nerf_b200 as pn
    from oracle import hashnerf_oracle as O
    from oracle.fixtures import mlp_weights
    w = {k: v.cuda().contiguous() for k, v in mlp_weights(7 + normals, normals).items()}
    gen = torch.Generator(device="cuda").manual_seed(P)
    S = 8
    feat = torch.randn(P, 32, device="cuda", generator=gen) * 0.3
    dirs = torch.nn.functional.normalize(torch.randn(P // S, 3, device="cuda", generator=gen), dim=-1)
    keep = torch.rand(P, device="cuda", generator=gen) > 0.2
    sh = pn.ops.sh_encode(dirs).repeat_interleave(S, 0)
    out = pn.ops.mlp_fwd(w, feat, dirs=dirs, samples_per_ray=S, keep=keep, mode="bf16")
    dout = torch.randn(P, out.shape[1], device="cuda", generator=gen)
    dfeat, _, dw = pn.ops.mlp_bwd(w, feat, dout, dirs=dirs, samples_per_ray=S, keep=keep, mode="bf16")
    # (1)
    e_out, e_dfeat, e_g = _emulated_bf16_mlp(w, feat, sh, dout, keep)
    assert _rel(out, e_out)[0] < 2e-3, ("forward vs emulation", _rel(out, e_out))
    assert _rel(dfeat, e_dfeat)[0] < 1.5e-2, ("dfeat vs emulation", _rel(dfeat, e_dfeat))
    for k in w:
        assert _rel(dw[k], e_g[k])[0] < 1.5e-2, ("d%s vs emulation" % k, _rel(dw[k], e_g[k]))
    # (2) + (3)
    wo = {k: v.clone().requires_grad_(True) for k, v in w.items()}
    fo = feat.clone().requires_grad_(True)
    ref = O.nerf_small(torch.cat([fo, sh], -1), wo)
    mask = torch.ones_like(ref); mask[~keep, -1] = 0
    ref = ref * mask
    l2, mx = _rel(out, ref)
    assert l2 < 5e-3 and mx < 1e-2, ("forward vs fp32 oracle", l2, mx)
    (ref * dout).sum().backward()
    assert _rel(dfeat, fo.grad)[0] < 0.15
    for k in w:
        assert _rel(dw[k], wo[k].grad)[0] < 0.15, (k, _rel(dw[k], wo[k].grad))
    out2 = pn.ops.mlp_fwd(w, feat, sh=sh, mode="bf16")
    ref2 = O.nerf_small(torch.cat([feat, sh], -1), w)
    assert _rel(out2, ref2)[0] < 5e-3


def _shape_case(kind, M, N, K, a_cols=None, a_col0=0, seed=0):
    """One GEMM in exactly the operand forms the MLP kernels use.
    kind 'fwd'  : D[128,N] = A[128,K] (K-major, optionally a column window of a wider tile) * W[N,K]^T (K-major)
    kind 'dgrad': D[128,N] = A[128,K] (K-major) * Wt[K,N] (MN-major)
    kind 'wgrad': D[64,N]  = At[128,64]^T * Bt[128,N]  (both MN-major, reduction over the 128 rows)"""
    g = torch.Generator(device="cuda").manual_seed(seed)
    if kind == "fwd":
        a_cols = a_cols or K
        A = torch.randn(128, a_cols, device="cuda", generator=g)
        B = torch.randn(N, K, device="cuda", generator=g)
        cfg = [128, a_cols, N, K, 128, N, K // 16, 0, 0, 128, (a_cols // 8) * 128, 256, 128, (K // 8) * 128, 256]
        # the kernel applies the column window by offsetting the start address; emulate by rolling the columns
        Ause = A[:, a_col0:a_col0 + K]
        if a_col0:
            A = torch.cat([A[:, a_col0:], A[:, :a_col0]], 1).contiguous()
        want = bf(Ause) @ bf(B).t()
    elif kind == "dgrad":
        A = torch.randn(128, K, device="cuda", generator=g)
        B = torch.randn(K, N, device="cuda", generator=g)
        rg = (N // 8) * 128
        cfg = [128, K, K, N, 128, N, K // 16, 0, 1, 128, (K // 8) * 128, 256, rg, 128, 2 * rg]
        want = bf(A) @ bf(B)
    else:
        A = torch.randn(128, 64, device="cuda", generator=g)
        B = torch.randn(128, N, device="cuda", generator=g)
        rga, rgb = 1024, (N // 8) * 128
        cfg = [128, 64, 128, N, 64, N, 8, 1, 1, rga, 128, 2 * rga, rgb, 128, 2 * rgb]
        want = bf(A).t() @ bf(B)
    D = _run(cfg, A, B, N)
    if kind == "wgrad":
        rows = torch.arange(64, device="cuda")
        D = D[(rows % 16) + 32 * (rows // 16)]
    return float((D - want).abs().max() / want.abs().max())


@pytest.mark.parametrize("case", [
    ("fwd", 128, 64, 32), ("fwd", 128, 16, 64), ("fwd", 128, 64, 64), ("fwd", 128, 16, 32), ("fwd", 128, 32, 16),
    ("dgrad", 128, 64, 16), ("dgrad", 128, 64, 64), ("dgrad", 128, 32, 64), ("dgrad", 128, 32, 16), ("dgrad", 128, 16, 32),
    ("wgrad", 64, 16, 128), ("wgrad", 64, 32, 128), ("wgrad", 64, 64, 128)])
def test_every_gemm_shape_of_the_mlp(case):
    err = _shape_case(*case)
    print(case, "rel err", err)
    assert err < 1e-3, (case, err)


@pytest.mark.parametrize("N", [700, 3101])
def test_fused_field_kernels_match_unfused(N):
    """pn_field_fwd/bwd_bf16 (hash + SH + MLP in one kernel; MLP backward + scatter in one kernel) against the
    unfused bf16 path (hash kernel -> fp32 features -> pn_mlp_*_bf16 -> dfeat -> hash scatter kernel).
    N = 3100 rays x 64 samples = 1550 tiles: every persistent CTA (296 of them) walks 5-6 tiles, so the
    warp-specialised kernels' double-buffer / mbarrier phases wrap several times; a ragged last tile in both."""
    import numpy as np
    import indoor_nerf_b200 as pn
    from oracle.fixtures import mlp_weights, synthetic_tables
    S = 64
    box = (torch.tensor([-3.0, -3.2, -2.9]), torch.tensor([3.1, 3.0, 3.3]))
    for normals in (False, True):
        emb = pn.HashEmbedder(box, log2_hashmap_size=15).cuda()
        with torch.no_grad():
            emb.table_storage.copy_(torch.from_numpy(synthetic_tables(16, 15, amp=0.3, salt=3)))
        w = mlp_weights(11, normals)
        net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32,
                           input_ch_views=16, predict_normals=normals).cuda()
        with torch.no_grad():
            net.sigma_net[0].weight.copy_(w["s0"]); net.sigma_net[1].weight.copy_(w["s1"])
            net.color_net[0].weight.copy_(w["c0"]); net.color_net[1].weight.copy_(w["c1"]); net.color_net[2].weight.copy_(w["c2"])
            if normals:
                net.normal_net[0].weight.copy_(w["n0w"]); net.normal_net[0].bias.copy_(w["n0b"])
                net.normal_net[2].weight.copy_(w["n2w"]); net.normal_net[2].bias.copy_(w["n2b"])
        gen = torch.Generator(device="cuda").manual_seed(4)
        pts = (torch.rand(N, S, 3, device="cuda", generator=gen) - 0.5) * 7.0      # some points fall outside the box
        dirs = torch.nn.functional.normalize(torch.randn(N, 3, device="cuda", generator=gen), dim=-1)
        dout = torch.randn(N, S, 7 if normals else 4, device="cuda", generator=gen)
        sh = pn.SHEncoder()
        results = []
        for fused in (True, False):
            for q in list(emb.parameters()) + list(net.parameters()):
                q.grad = None
            pn.set_mlp_mode("bf16")
            try:
                if fused:
                    out = pn.run_network(pts, dirs, net, emb, sh)
                else:
                    feat, keep = emb(pts.reshape(-1, 3))
                    keys, weights = net.kernel_weights()
                    x = torch.cat([feat, sh(dirs[:, None].expand(N, S, 3).reshape(-1, 3))], -1)
                    out = pn.ops.MlpFn.apply(x, None, keys, *weights)
                    mask = torch.ones_like(out); mask[~keep, -1] = 0
                    out = (out * mask).reshape(N, S, -1)
                (out * dout).sum().backward()
            finally:
                pn.set_mlp_mode("fp32")
            results.append((out.detach(), [e.weight.grad.clone() for e in emb.embeddings],
                            [q.grad.clone() for q in net.parameters()]))
        (o1, t1, w1), (o2, t2, w2) = results
        assert _rel(o1, o2)[0] < 2e-3, ("fused forward", _rel(o1, o2))        # features are bf16-rounded once in both
        for l in (0, 5, 10, 15):
            assert _rel(t1[l], t2[l])[0] < 2e-2, ("table grad level %d" % l, _rel(t1[l], t2[l]))
        # The fused backward takes its voxel indices from point_cell<false> (reciprocal multiply + exact fallback), the
        # unfused scatter from the reference form: the SETS of table rows that received a gradient must be identical at
        # every level — the indices are bit-exact in the bf16 mode too
        for l in range(16):
            s1, s2 = (t1[l] != 0).any(-1), (t2[l] != 0).any(-1)
            assert bool((s1 == s2).all()), ("gradient support differs at level %d: %d rows" % (l, int((s1 != s2).sum())))
        for a, b in zip(w1, w2):
            assert _rel(a, b)[0] < 2e-2, ("weight grad", _rel(a, b))


def test_render_bf16_mode_vs_fp32_mode():
    """End-to-end bar of the bf16 tensor-core mode (BASELINE north_star: 2e-3 for rgb / depth / weights): the same
    4096 rays rendered by the fused bf16 kernels and by the fp32 kernels (themselves 1e-5 from the reference), with
    identical parameters and random draws; then the training gradients of both modes.

    raw2outputs is discontinuous in sigma at 0 for the LAST sample of a ray (its interval is 1e10 long,
    run_nerf.py:365: alpha jumps from 0 to 1 as sigma crosses 0) and steep near 0 elsewhere, so rays whose density
    hovers around zero amplify ANY perturbation (the fp32 kernels vs the reference included, at their own scale).
    The strict bar is therefore asserted on the rays whose samples are clear of sigma = 0, the typical-ray (median)
    error on all rays, and the scalar training loss."""
    import numpy as np
    import indoor_nerf_b200 as pn
    from oracle.fixtures import mlp_weights, synthetic_tables
    from tests.test_gpu_parity import _Rng, embedder_from, mlp_from, mlp_grads
    box = (np.array([-3.2, -3.1, -3.3], np.float32), np.array([3.1, 3.3, 3.2], np.float32))
    emb = embedder_from(pn, box[0], box[1], 19, 512, synthetic_tables(16, 19, amp=0.3, salt=9)).train()
    nets = [mlp_from(pn, mlp_weights(41)), mlp_from(pn, mlp_weights(42))]
    N = 4096
    rs = np.random.RandomState(2)
    o = (rs.randn(N, 3) * 0.2 + np.array([0, 0, 4.0])).astype(np.float32)
    d = (rs.randn(N, 3) * 0.8 - o).astype(np.float32)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    rays = torch.from_numpy(np.concatenate([o, d, np.full((N, 1), 2.0, np.float32), np.full((N, 1), 6.0, np.float32), d], -1)).cuda()
    t_rand, u = torch.rand(N, 64, device="cuda"), torch.rand(N, 128, device="cuda")
    target = torch.rand(N, 3, device="cuda")
    sh = pn.SHEncoder()
    query = lambda inputs, viewdirs, fn: pn.run_network(inputs, viewdirs, fn, embed_fn=emb, embeddirs_fn=sh)

    def run(mode, n_imp):
        pn.set_mlp_mode(mode)
        try:
            for prm in list(emb.parameters()) + [q for n in nets for q in n.parameters()]:
                prm.grad = None
            with _Rng([t_rand, u], []):
                ret = pn.render_rays(rays, nets[0], query, 64, embed_fn=emb, retraw=True, perturb=1.0, N_importance=n_imp,
                                     network_fine=nets[1], white_bkgd=True)
            loss = ((ret["rgb_map"] - target) ** 2).mean()
            if n_imp:
                loss = loss + ((ret["rgb0"] - target) ** 2).mean()
            loss.backward()
            return ret, [e.weight.grad.clone() for e in emb.embeddings], [mlp_grads(n) for n in nets], float(loss.detach())
        finally:
            pn.set_mlp_mode("fp32")

    # (A) one pass at identical sample positions: MLP precision, then compositing on the rays clear of sigma = 0
    (a32, _, _, _), (a16, _, _, _) = run("fp32", 0), run("bf16", 0)
    raw_l2 = [float((a16["raw"][..., c] - a32["raw"][..., c]).norm() / a32["raw"][..., c].norm()) for c in range(4)]
    s32, s16 = a32["raw"][..., 3], a16["raw"][..., 3]
    tol = 0.05 * float(s32.abs().median())
    stable = (s32.abs() > tol).all(-1) & (torch.sign(s32) == torch.sign(s16)).all(-1)
    rep = {"stable_ray_fraction": float(stable.float().mean())}
    for k in ["rgb_map", "acc_map", "depth_map"]:
        a, b = a16[k].detach(), a32[k].detach()
        fin = torch.isfinite(b) & torch.isfinite(a)
        err = ((a - b).abs() / b[fin].abs().max()).reshape(N, -1).amax(-1)
        err = torch.where(torch.isfinite(err), err, torch.zeros_like(err))
        rep[k] = (float(err.median()), float(err[stable].max()) if stable.any() else 0.0, float(err.max()))
    wl2 = float((a16["rgb_map"] - a32["rgb_map"]).norm() / a32["rgb_map"].norm())
    print("bf16 vs fp32, single pass: raw per-channel L2", raw_l2, "| (median, max over stable rays, max) err/scale", rep,
          "| rgb_map L2", wl2)
    assert max(raw_l2) < 2e-2, raw_l2          # pre-sigmoid colour logits of a random net cancel heavily; see the maps below
    assert rep["stable_ray_fraction"] > 0.02
    for k in ["rgb_map", "acc_map", "depth_map"]:
        assert rep[k][0] < 2e-3 and rep[k][1] < 2e-3, (k, rep[k])
    # ALL rays: whatever exceeds 2e-3 must be one of the enumerated sigma ~ 0 rays (`~stable`: some sample within 5 % of the
    # median |sigma| of zero, or a sign flip between the modes), never a ray clear of the discontinuity; their number is
    # bounded by the size of that set, and the worst of them stays a bounded error (an alpha flipping between 0 and 1 on
    # one sample), not garbage
    for k in ["rgb_map", "acc_map", "depth_map"]:
        a, b = a16[k].detach(), a32[k].detach()
        fin = torch.isfinite(b) & torch.isfinite(a)
        err = ((a - b).abs() / b[fin].abs().max()).reshape(N, -1).amax(-1)
        err = torch.where(torch.isfinite(err), err, torch.zeros_like(err))
        over = err > 2e-3
        assert not bool((over & stable).any()), (k, "a stable ray exceeds 2e-3")
        assert int(over.sum()) <= int((~stable).sum()), (k, int(over.sum()), int((~stable).sum()))
        assert rep[k][2] <= 1.0 + 1e-6, (k, "all-ray max", rep[k][2])
    print("rays beyond 2e-3 (all within the %d sigma~0 rays):" % int((~stable).sum()),
          {k: int((((a16[k] - a32[k]).abs() / a32[k][torch.isfinite(a32[k])].abs().max()).reshape(N, -1).amax(-1) > 2e-3).sum())
           for k in ["rgb_map", "acc_map"]})

    # (B) the whole coarse + fine step: typical ray, loss, gradients
    (r32, t32, w32, l32), (r16, t16, w16, l16) = run("fp32", 128), run("bf16", 128)
    med = {}
    for k in ["rgb0", "acc0", "depth0", "rgb_map", "acc_map", "depth_map"]:
        a, b = r16[k].detach(), r32[k].detach()
        fin = torch.isfinite(b) & torch.isfinite(a)
        med[k] = float(((a - b).abs()[fin] / b[fin].abs().max()).median())
    g = {"table%d" % l: _rel(t16[l], t32[l])[0] for l in (0, 5, 10, 15)}
    for i in range(2):
        for k in w32[i]:
            g["net%d.%s" % (i, k)] = _rel(w16[i][k], w32[i][k])[0]
    print("bf16 vs fp32, full step: median err/scale", med, "| loss", l16, l32, "| gradient relative L2 error", g)
    assert max(med.values()) < 2e-3, med
    assert abs(l16 - l32) / l32 < 2e-3
    assert max(g.values()) < 0.25, g
