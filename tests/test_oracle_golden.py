"""The oracle (oracle/hashnerf_oracle.py) against the golden vectors produced by the live
reference (oracle/make_golden.py).  CPU only.  Integer results must be identical; float results
are compared at 1e-6 relative (identical ATen ops; the slack only covers GEMM blocking on a
different host CPU)."""
import numpy as np
import pytest
import torch

from oracle import hashnerf_oracle as O
from oracle.fixtures import synthetic_tables

T = torch.from_numpy


def close(a, b, rtol=1e-6, atol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    nan = np.isnan(b)
    assert (np.isnan(a) == nan).all()
    scale = np.abs(b[~nan]).max() if (~nan).any() else 1.0
    np.testing.assert_allclose(a[~nan], b[~nan], rtol=rtol, atol=atol + rtol * scale)


def test_hash_primitives(golden):
    g = golden("hash_primitives")
    c = T(g["corners"])
    for k in (12, 19, 22):
        assert (O.hash_coords(c, k).numpy() == g["h%d" % k]).all()
    assert [float(r) for r in O.level_resolutions(16, 512)] == list(g["res512"])
    assert [float(r) for r in O.level_resolutions(16, 1024)] == list(g["res1024"])
    # SURVEY.md §8c seed vectors
    assert list(g["h19"][:5]) == [0, 128476, 474075, 212356, 36881]
    assert list(g["h22"][:5]) == [0, 2749916, 1522651, 2309508, 3706897]


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_hash_embed(golden, tag):
    g = golden("hash_embed_" + tag)
    log2T, finest = int(g["log2T"]), int(g["finest"])
    tables = [T(t).requires_grad_(True) for t in synthetic_tables(16, log2T, salt=int(g["salt"]))]
    res = O.level_resolutions(16, finest)
    feat, keep, idx = O.hash_embed(T(g["x"]), T(g["box_min"]), T(g["box_max"]), tables, res, log2T,
                                   return_indices=True)
    assert (idx.numpy() == g["idx"]).all()
    assert (keep.numpy() == g["keep"]).all()
    assert (feat.detach().numpy() == g["feat"]).all()          # same ops, same order: bit-exact
    (feat * T(g["dfeat"])).sum().backward()
    close(torch.stack([t.grad.abs().sum() for t in tables]), g["grad_abs_sum"], 1e-5)
    for l in (0, 7, 15):
        close(tables[l].grad[T(g["grad_rows_%d" % l])], g["grad_vals_%d" % l], 1e-6)


def test_known_voxel_vertices():
    # SURVEY.md §8c
    bmin, bmax = torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3)
    x = torch.tensor([[0.1, -0.7, 1.2], [1.5, 1.5, 1.5], [-1.5, 0, 2.0]])
    vmin, vmax, h, keep = O.voxel_vertices(x, bmin, bmax, torch.tensor(512.0), 19)
    assert vmin[0].tolist() == [0.099609375, -0.703125, 1.1953125]
    assert h[0].tolist() == [391333, 462920, 344340, 502265, 391334, 462923, 344343, 502266]
    assert h[1].tolist() == [281088, 188821, 390065, 219172, 281089, 188820, 390064, 219173]
    assert h[2].tolist() == [432896, 12437, 393393, 43812, 432897, 12436, 393392, 43813]
    assert keep.all(-1).tolist() == [True, True, False]


@pytest.mark.parametrize("mode", ["train", "eval"])
def test_hash_embed_quant(golden, mode):
    g = golden("hash_embed_quant_" + mode)
    tables = [T(t) for t in synthetic_tables(16, 15, salt=5)]
    res = O.level_resolutions(16, 512)
    x, bmin, bmax = T(g["x"]), T(g["box_min"]), T(g["box_max"])
    # calibration as the first training call performs it (quantization.py:97-119)
    quant = []
    for l in range(16):
        q = T(g["q%d" % l])
        _, _, hidx, _ = O.voxel_vertices(x, bmin, bmax, res[l], 15)
        rs, vmax, rmin, rmax = O.lbq_calibrate(tables[l][hidx], symmetric=False)
        assert float(rs) == float(q[1]) and float(vmax) == float(q[2])
        assert float(rmin) == float(q[3]) and float(rmax) == float(q[4])
        quant.append(O.lbq_scalars(q[0], rs, vmax, False, training=(mode == "train")) + (mode == "train",))
    feat, keep = O.hash_embed(x, bmin, bmax, tables, res, 15, quant=quant)
    assert (feat.numpy() == g["feat"]).all()
    assert (keep.numpy() == g["keep"]).all()


def test_sh4(golden):
    g = golden("sh4")
    assert (O.sh4(T(g["dirs"])).numpy() == g["out"]).all()
    np.testing.assert_allclose(O.sh4(torch.tensor([[0.6, 0.0, 0.8]]))[0, [0, 2, 3, 6]].numpy(),
                               [0.282094806, 0.390882015, -0.293161511, 0.290160269], rtol=1e-6)


@pytest.mark.parametrize("tag", ["plain", "normals"])
def test_nerf_small(golden, tag):
    g = golden("nerf_small_" + tag)
    w = {k[2:]: T(g[k]).requires_grad_(True) for k in g.files if k.startswith("w_")}
    x = T(g["x"]).requires_grad_(True)
    out = O.nerf_small(x, w)
    close(out.detach(), g["out"], 1e-6)
    (out * T(g["gout"])).sum().backward()
    close(x.grad, g["gx"], 1e-5)
    for k in w:
        close(w[k].grad, g["g_" + k], 1e-5)


def test_nerf_small_quant(golden):
    g = golden("nerf_small_quant")
    w = {k[2:]: T(g[k]) for k in g.files if k.startswith("w_") and k != "w_q"}
    x = T(g["x"])
    aq, wq = T(g["act_q"]), T(g["w_q"])
    # calibration values: weight quantizer sees W0, activation quantizer sees relu(x W0q^T)
    rs_w, _, _, _ = O.lbq_calibrate(w["s0"], symmetric=True)
    assert float(rs_w) == float(wq[1])
    for training, key in ((True, "out"), (False, "out_eval")):
        qw = O.lbq_scalars(wq[0], wq[1], None, True, training) + (training,)
        qa = O.lbq_scalars(aq[0], aq[1], aq[2], False, training) + (training,)
        out = O.nerf_small(x, w, quant=dict(weight=qw, act=qa))
        close(out, g[key], 1e-6)


@pytest.mark.parametrize("tag", ["s64", "s192n", "s128"])
def test_raw2outputs(golden, tag):
    g = golden("raw2outputs_" + tag)
    raw = T(g["raw"]).requires_grad_(True)
    normals = raw.shape[-1] == 7
    noise = T(g["noise"]) * 1.0 if g["noise"].size else None
    outs = O.raw2outputs(raw, T(g["z"]), T(g["d"]), noise, bool(g["white"]), normals)
    for k, v in zip(["rgb", "disp", "acc", "weights", "depth", "sparsity"], outs):
        close(v.detach(), g[k], 1e-6)
    rgb, disp, acc, wts, depth, sp = outs[:6]
    ok = torch.isfinite(depth)
    loss = (rgb * T(g["cot_rgb"])).sum() + (depth[ok] * T(g["cot_depth"])[ok]).sum() + (acc * T(g["cot_acc"])).sum() \
        + (sp * T(g["cot_sp"])).sum() + (wts * T(g["cot_w"])).sum() + (disp[ok] * T(g["cot_disp"])[ok]).sum()
    if normals:
        close(outs[6].detach(), g["normal"], 1e-6)
        loss = loss + (outs[6] * T(g["cot_normal"])).sum()
    loss.backward()
    close(raw.grad, g["graw"], 1e-5)


def test_raw2outputs_known(golden):
    g = golden("raw2outputs_known")
    outs = O.raw2outputs(T(g["raw"]), T(g["z"]), T(g["d"]), None, True)
    np.testing.assert_allclose(outs[0][0].numpy(), [0.729956150, 0.362515152, 0.409085542], rtol=1e-6)
    np.testing.assert_allclose(outs[3][0].numpy(), [0.393469334, 0.599792719, 0, 0.006737946], rtol=1e-6)
    np.testing.assert_allclose(float(outs[5]), 0.707309961, rtol=1e-6)
    for k, v in zip(["rgb", "disp", "acc", "weights", "depth", "sparsity"], outs):
        close(v, g[k], 1e-6)


def test_sample_pdf(golden):
    g = golden("sample_pdf")
    bins, w = T(g["bins"]), T(g["weights"])
    assert (O.sample_pdf(bins, w, 128, det=True).numpy() == g["det"]).all()
    rnd = O.sample_pdf(bins, w, 128, u=T(g["u"]))
    assert (rnd.numpy() == g["rnd"]).all()
    merged = torch.sort(torch.cat([T(g["z"]), rnd], -1), -1)[0]
    assert (merged.numpy() == g["merged"]).all()
    k = golden("sample_pdf_known")
    out = O.sample_pdf(torch.tensor([[2., 3, 4, 5]]), torch.tensor([[.1, .7, .2]]), 5, det=True)
    assert (out.numpy() == k["out"]).all()
    np.testing.assert_allclose(out[0].numpy(), [2.0, 3.214279175, 3.571427584, 3.928575993, 5.0], rtol=1e-6)


def test_rays(golden):
    g = golden("rays")
    ro, rd = O.get_rays(int(g["H"]), int(g["W"]), g["K"], T(g["c2w"]))
    assert (rd.numpy() == g["rays_d"]).all() and (ro.numpy() == g["rays_o"]).all()
    no, nd = O.ndc_rays(int(g["H"]), int(g["W"]), float(g["K"][0][0]), 1., ro - torch.tensor([0, 0, 10.0]), rd)
    assert (no.numpy() == g["ndc_o"]).all() and (nd.numpy() == g["ndc_d"]).all()


def _render_case(g):
    log2T = int(g["log2T"])
    tables = [T(t).requires_grad_(True) for t in synthetic_tables(16, log2T, amp=float(g["amp"]), salt=int(g["salt"]))]
    res = O.level_resolutions(16, 512)
    bmin, bmax = T(g["box_min"]), T(g["box_max"])
    ws = [{k[3:]: T(g[k]).requires_grad_(True) for k in g.files if k.startswith(p)} for p in ("w0_", "w1_")]
    normals = "w0_n0w" in g.files
    embed = lambda x: O.hash_embed(x, bmin, bmax, tables, res, log2T)
    q = [lambda pts, vd, w=w: O.run_network(pts, vd, embed, lambda x: O.nerf_small(x, w)) for w in ws]
    std = float(g["raw_noise_std"])
    perturb = float(g["perturb"]) > 0
    ret = O.render_rays(T(g["rays"]), q[0], q[1], 64, int(g["N_importance"]),
                        t_rand=T(g["t_rand"]) if perturb else None, u=T(g["u"]) if perturb else None,
                        noise0=T(g["noise0"]) * std if std > 0 else None,
                        noise1=T(g["noise1"]) * std if std > 0 else None,
                        white_bkgd=bool(g["white"]), predict_normals=normals)
    return ret, tables, ws


@pytest.mark.parametrize("tag", ["blender", "det", "normals_noise"])
def test_render_rays(golden, tag):
    g = golden("render_rays_" + tag)
    ret, tables, ws = _render_case(g)
    assert (ret["pts"].detach().numpy() == g["pts"]).all()      # => identical sample bins and z order
    for k in ["rgb_map", "depth_map", "acc_map", "sparsity_loss", "rgb0", "depth0", "acc0",
              "sparsity_loss0", "z_std", "raw"] + (["normal_map", "normal0"] if "normal_map" in g.files else []):
        close(ret[k].detach(), g[k], 2e-6)
    target = T(g["target"])
    loss = ((ret["rgb_map"] - target) ** 2).mean() + ((ret["rgb0"] - target) ** 2).mean() \
        + 1e-3 * (ret["sparsity_loss"].sum() + ret["sparsity_loss0"].sum())
    close(loss.detach(), g["loss"], 1e-6)
    loss.backward()
    close(torch.stack([t.grad.abs().sum() for t in tables]), g["g_table_abs_sum"], 1e-4)
    for i, w in enumerate(ws):
        for k, p in w.items():
            ref = g["g_m%d_%s" % (i, k)]
            got = p.grad if p.grad is not None else torch.zeros_like(p)
            close(got, ref, 1e-4)


def test_tv_loss(golden):
    g = golden("tv_loss")
    table = T(synthetic_tables(1, 19, salt=int(g["salt"]))[0])
    for level in (0, 5, 15):
        v = O.tv_loss_level(table, level, 19, T(g["mv_%d" % level]))
        close(v, g["tv_%d" % level], 1e-6)


def test_radam(golden):
    g = golden("radam")
    p = [T(g["init0"].copy()), T(g["init1"].copy())]
    opt = O.RAdamState([dict(params=[p[0]], weight_decay=1e-6), dict(params=[p[1]], eps=1e-15)], lr=5e-4)
    for it in range(8):
        for i in range(2):
            p[i].grad = T(g["g%d_%d" % (it, i)])
        opt.step()
        for i in range(2):
            close(p[i], g["p%d_%d" % (it, i)], 1e-6)
