"""Parity of the §8f kernels (through the C ABI) with the reference: on-GPU ray batching, test-set evaluation
(PSNR / SSIM / to8b, render_path, ComprehensiveEvaluator) and the bit-packed A-CAQ export.
Bars: rays, targets, uint8 images, integer codes and dequantised values bit-exact; PSNR / SSIM (fp64 reductions,
order not fixed) 1e-6 / 1e-5 relative to the oracle."""
import os
import pickle
import zlib

import numpy as np
import pytest
import torch

from oracle import dataio_oracle as D
from oracle import hashnerf_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pn():
    import indoor_nerf_b200 as pkg
    assert torch.cuda.is_available()
    return pkg


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- ray bank ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("f64", [True, False])
def test_ray_bank_kernel_golden(pn, golden, f64):
    g = golden("ray_bank")
    H, W = int(g["H"]), int(g["W"])
    ids = cu(g["order"].astype(np.int64))
    idx = cu(g["i_train"].astype(np.int32))
    rays, tgt = pn.ops.ray_bank_batch(ids, H, W, g["K"], cu(g["poses"]), idx, cu(g["images"]), f64_dirs=f64)
    assert torch.equal(tgt.cpu(), torch.from_numpy(g["shuffled"][:, 2]))
    assert torch.equal(rays[0].cpu(), torch.from_numpy(g["shuffled"][:, 0]))
    if f64:
        assert torch.equal(rays[1].cpu(), torch.from_numpy(g["shuffled"][:, 1]))        # get_rays_np + astype, bit for bit
    else:
        for n, img in enumerate(g["i_train"]):                                           # get_rays (fp32), bit for bit
            ro, rd = pn.get_rays(H, W, g["K"], cu(g["poses"][img, :3, :4]))
            sel = torch.nonzero(ids // (H * W) == n)[:, 0]
            assert torch.equal(rays[1][sel], rd.reshape(-1, 3)[ids[sel] % (H * W)])
    # uint8 images, [N,3,4] poses, rays only
    img8 = (g["images"] * 255).astype(np.uint8)
    rays2, tgt2 = pn.ops.ray_bank_batch(ids, H, W, g["K"], cu(g["poses"][:, :3, :4]), idx, cu(img8), f64_dirs=f64)
    assert torch.equal(rays2, rays)
    want = np.stack([(img8 / 255.).astype(np.float32)[i] for i in g["i_train"]]).reshape(-1, 3)[g["order"]]
    assert torch.equal(tgt2.cpu(), torch.from_numpy(want))
    assert torch.equal(pn.ops.ray_bank_batch(ids, H, W, g["K"], cu(g["poses"]), idx, f64_dirs=f64), rays)
    assert pn.ops.ray_bank_batch(ids[:0], H, W, g["K"], cu(g["poses"]), idx, cu(g["images"]))[0].shape == (2, 0, 3)


def test_ray_bank_matches_reference_bank_over_epochs(pn):
    """A 40x30, 6-view scene: every batch of two epochs equals slicing the reference's rays_rgb tensor."""
    from indoor_nerf_b200 import synthetic
    sc = synthetic.blender_scene(30, 40, n_views=6)
    rs = np.random.RandomState(4)
    images = rs.rand(6, 30, 40, 3).astype(np.float32)
    i_train = [0, 1, 3, 5]
    ref = D.rays_rgb_bank(30, 40, sc["K"], sc["poses"], images, i_train)
    np.random.seed(11)
    order = D.shuffle_order(ref.shape[0])
    ref = torch.from_numpy(ref[order]).cuda()
    bank = pn.RayBank(30, 40, sc["K"], sc["poses"], images, i_train, device="cuda")
    np.random.seed(11)
    bank.shuffle()
    torch.manual_seed(5)
    state = torch.cuda.get_rng_state()
    i_batch, N_rand = 0, 1000
    for it in range(12):
        torch.cuda.set_rng_state(state)
        rays, tgt = bank.next_batch(N_rand)
        torch.cuda.set_rng_state(state)
        batch = torch.transpose(ref[i_batch:i_batch + N_rand], 0, 1)
        assert torch.equal(rays, batch[:2]) and torch.equal(tgt, batch[2]), "iteration %d" % it
        i_batch += N_rand
        if i_batch >= ref.shape[0]:
            ref = ref[torch.randperm(ref.shape[0], device="cuda")]
            i_batch = 0
        state = torch.cuda.get_rng_state()
    np.random.seed(2)
    rays, tgt, img_i = bank.sample_image(256, precrop_frac=0.5)
    ro, rd = pn.get_rays(30, 40, sc["K"], torch.from_numpy(sc["poses"][img_i, :3, :4]).cuda())
    np.random.seed(2)
    assert np.random.choice(i_train) == img_i
    sel = np.random.choice(14 * 20, size=[256], replace=False)
    rows, cols = torch.from_numpy(15 - 7 + sel // 20).cuda(), torch.from_numpy(20 - 10 + sel % 20).cuda()
    assert torch.equal(rays[1], rd[rows, cols]) and torch.equal(rays[0], ro[rows, cols])
    assert torch.equal(tgt, torch.from_numpy(images[img_i]).cuda()[rows, cols])


# ---- evaluation -----------------------------------------------------------------------------------------------------------
def test_psnr_ssim_to8b_kernels(pn, golden):
    g = golden("eval_psnr")
    rgb, gt = cu(g["rgb"]), cu(g["gt"])
    s = pn.ops.image_sqerr(rgb, gt)
    p = -10. * np.log10(float(s) / rgb.numel())
    assert p == pytest.approx(float(g["p_render_path"]), rel=1e-6)
    assert torch.equal(pn.ops.to8b(rgb).cpu(), torch.from_numpy(g["rgb8"]))
    edge = torch.tensor([-0.5, 0, 1, 1.5, 0.999999, 1 / 255, 0.00392, 0.5], device="cuda")
    assert torch.equal(pn.ops.to8b(edge).cpu(), torch.from_numpy(D.to8b(edge.cpu().numpy())))
    assert float(pn.ops.image_ssim(gt, rgb)) == pytest.approx(D.ssim(g["gt"], g["rgb"]), abs=2e-5)
    assert float(pn.ops.image_ssim(gt, rgb)) == pytest.approx(D.ssim(g["gt"].astype(np.float64), g["rgb"].astype(np.float64)), abs=1e-9)
    assert float(pn.ops.image_ssim(gt, gt)) == pytest.approx(1.0, abs=1e-12)
    # ragged sizes around the 32x8 tile and a single channel
    rs = np.random.RandomState(1)
    for (h, w, c) in ((7, 7, 1), (8, 39, 3), (41, 33, 2), (123, 77, 3)):
        a = rs.rand(h, w, c).astype(np.float32)
        b = np.clip(a + rs.randn(h, w, c).astype(np.float32) * 0.1, 0, 1).astype(np.float32)
        want = D.ssim(a.astype(np.float64), b.astype(np.float64))
        assert float(pn.ops.image_ssim(cu(a), cu(b))) == pytest.approx(want, abs=1e-9), (h, w, c)
        assert float(pn.ops.image_sqerr(cu(a), cu(b))) == pytest.approx(float(np.square((a - b).astype(np.float64)).sum()), rel=1e-6)
    with pytest.raises(pn._lib.PocketNerfError):
        pn.ops.image_ssim(torch.zeros(5, 9, 3, device="cuda"), torch.zeros(5, 9, 3, device="cuda"))


def small_model(pn, H=24, W=32, **over):
    from indoor_nerf_b200 import model as pmodel, synthetic
    sc = synthetic.blender_scene(H, W, n_views=4)
    torch.manual_seed(0)
    a = pmodel.default_args(bounding_box=sc["bounding_box"], log2_hashmap_size=12, **over)
    kw_train, kw_test, _, _, _ = pmodel.create_nerf(a, device="cuda")
    with torch.no_grad():
        kw_test["embed_fn"].table_storage.mul_(3000.0)
    kw_test = dict(kw_test, near=sc["near"], far=sc["far"])
    return sc, kw_train, kw_test


def test_render_path_and_evaluator(pn, tmp_path):
    sc, _, kw = small_model(pn)
    H, W = sc["H"], sc["W"]
    hwf = [H, W, sc["focal"]]
    poses = torch.from_numpy(sc["poses"][:3])
    rs = np.random.RandomState(0)
    gts = rs.rand(3, H, W, 3).astype(np.float32)
    savedir = str(tmp_path / "testset")
    rgbs, depths = pn.render_path(poses, hwf, sc["K"], 1024, kw, gt_imgs=gts, savedir=savedir)
    assert rgbs.shape == (3, H, W, 3) and depths.shape == (3, H, W)
    psnrs = []
    for i in range(3):
        with torch.no_grad():
            rgb, depth, acc, _ = pn.render(H, W, sc["K"], chunk=1024, c2w=poses[i, :3, :4].cuda(), **kw)
        assert np.array_equal(rgbs[i], rgb.cpu().numpy())
        assert np.allclose(depths[i], ((depth - sc["near"]) / (sc["far"] - sc["near"])).cpu().numpy(), equal_nan=True)
        psnrs.append(D.psnr_render_path(rgbs[i], gts[i]))
        # the PNG holds to8b(rgb): decode the single IDAT chunk by hand
        raw = open(os.path.join(savedir, "%03d.png" % i), "rb").read()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        k = raw.index(b"IDAT")
        n = int.from_bytes(raw[k - 4:k], "big")
        pix = np.frombuffer(zlib.decompress(raw[k + 4:k + 4 + n]), np.uint8).reshape(H, 1 + 3 * W)[:, 1:].reshape(H, W, 3)
        assert np.array_equal(pix, D.to8b(rgbs[i]))
    assert np.allclose(pn.render_path.last_psnrs, psnrs, rtol=1e-6)
    pk = [f for f in os.listdir(savedir) if f.endswith(".pkl")]
    assert len(pk) == 1 and np.allclose(pickle.load(open(os.path.join(savedir, pk[0]), "rb")), psnrs, rtol=1e-6)
    # render_factor: half resolution, no PSNR
    r2, d2 = pn.render_path(poses[:1], hwf, sc["K"], 1024, kw, gt_imgs=gts, render_factor=2)
    assert r2.shape == (1, H // 2, W // 2, 3)

    ev = pn.ComprehensiveEvaluator(device="cuda", lpips_fn=None)
    avg, preds, per = ev.evaluate_test_set(pn.render, poses, hwf, sc["K"], 1024, kw, [g for g in gts])
    assert len(preds) == 3 and np.array_equal(preds[1], rgbs[1])
    for i in range(3):
        assert per[i]["psnr"] == pytest.approx(psnrs[i], rel=1e-6)
        assert per[i]["ssim"] == pytest.approx(D.ssim(gts[i].astype(np.float64), rgbs[i].astype(np.float64)), abs=1e-9)
    assert avg["psnr"] == pytest.approx(np.mean(psnrs), rel=1e-6) and avg["std_ssim"] >= 0
    m = ev.compute_metrics(torch.from_numpy(rgbs[0]).cuda(), gts[0])
    assert m["psnr"] == pytest.approx(psnrs[0], rel=1e-6) and np.isnan(m["lpips"]) or ev.lpips_fn is not None
    ev.record_test_metrics(10, avg)
    ev.record_memory_usage(10)
    ev.save_metrics(str(tmp_path / "metrics.pkl"))
    assert pickle.load(open(str(tmp_path / "metrics.pkl"), "rb"))["test"]["iter"] == [10]


# ---- A-CAQ export ------------------------------------------------------------------------------------------------------------
CASES = [("asym", k, False, "table") for k in range(6)] + [("sym", k, True, "w0") for k in range(3)]


@pytest.mark.parametrize("tag,k,sym,src", CASES)
def test_quant_pack_unpack_golden(pn, golden, tag, k, sym, src):
    g = golden("quant_export")
    x = g[src]
    pre = "%s_%d_" % (tag, k)
    bits, scale, zp, qmin, qmax = D.lbq_eval_params(g[pre + "bits"], g[pre + "range_scale"], g[pre + "v_max"], sym)
    row = torch.tensor([scale, np.float32(scale + np.float32(1e-8)), zp, qmin, qmax, 1, 0, 0], dtype=torch.float32, device="cuda")
    words = pn.ops.quant_pack(cu(x), row, bits)
    want = D.pack_bits(D.quant_codes(x, scale, zp, qmin, qmax), bits)
    assert np.array_equal(words.cpu().numpy().view(np.uint32), want)
    y = pn.ops.quant_unpack(words, x.size, row, bits)
    assert np.array_equal(y.cpu().numpy(), g[pre + "out"])                        # the reference quantiser's eval output


def test_pnq_export_renders_identically(pn, tmp_path):
    """Eval-mode fake-quantised model vs the model loaded from its bit-packed export: identical embeddings and frames."""
    from indoor_nerf_b200 import model as pmodel, quant_export
    sc, kw_train, kw = small_model(pn, use_quantization=True)
    emb, nets = kw["embed_fn"], [kw["network_fn"], kw["network_fine"]]
    bits = [3.0, 4.2, 5.0, 6.0, 6.6, 7.0, 8.0, 8.0, 9.0, 10.0, 11.3, 12.0, 14.0, 16.0, 20.0, 26.0]
    x = (torch.rand(4096, 3, device="cuda") - 0.5) * 2.5
    for l, q in enumerate(emb.quantizers):
        q.calibrate(emb.embeddings[l].weight.detach())
        q.soft_bits.data.fill_(bits[l])
    for n in nets:
        n.sigma_weight_quantizer.calibrate(n.sigma_net[0].weight.detach())
        n.sigma_weight_quantizer.soft_bits.data.fill_(7.4)
        with torch.no_grad():
            h = torch.relu(emb.eval()(x)[0] @ n.sigma_net[0].weight.t())
        n.sigma_act_quantizers[0].calibrate(h)
    emb.eval(); [n.eval() for n in nets]
    emb.current_step = 1000
    with torch.no_grad():
        feat_q, keep_q = emb(x)
        frame_q = pn.render(sc["H"], sc["W"], sc["K"], chunk=512, c2w=torch.from_numpy(sc["poses"][0][:3, :4]).cuda(), **kw)[0]
    path = str(tmp_path / "m.pnq")
    header = quant_export.export_quantized(path, emb, {"network_fn": nets[0], "network_fine": nets[1]})
    assert header["bytes"]["payload"] < 0.45 * header["bytes"]["fp32_equivalent"]
    assert os.path.getsize(path) < 0.5 * header["bytes"]["fp32_equivalent"]

    sc2, _, kw2 = small_model(pn, use_quantization=True)
    emb2, nets2 = kw2["embed_fn"], [kw2["network_fn"], kw2["network_fine"]]
    with torch.no_grad():
        emb2.table_storage.zero_()
    quant_export.load_quantized(path, emb2, {"network_fn": nets2[0], "network_fine": nets2[1]})
    emb2.eval(); [n.eval() for n in nets2]
    with torch.no_grad():
        feat_l, keep_l = emb2(x)
        frame_l = pn.render(sc["H"], sc["W"], sc["K"], chunk=512, c2w=torch.from_numpy(sc["poses"][0][:3, :4]).cuda(), **kw2)[0]
    assert torch.equal(feat_l, feat_q) and torch.equal(keep_l, keep_q)
    assert torch.equal(frame_l, frame_q)


def test_quant_pack_full_size_table_properties(pn):
    """BASELINE size (T = 2^19, all 16 levels in one call): unpack(pack(x)) equals the quantiser's eval forward on every
    one of the 16.8 M values and lies on the 6-bit lattice.  (pack(unpack(words)) is NOT the identity: the reference's
    fake-quant is not idempotent, which is why load_quantized switches the table quantisers off.)"""
    torch.manual_seed(0)
    q = pn.LearnedBitwidthQuantizer(init_bits=8.0, min_bits=2.0, max_bits=32.0, symmetric=False).cuda()
    x = (torch.rand(16 * (1 << 19) * 2, device="cuda") - 0.5) * 2e-4
    q.calibrate(x)
    q.soft_bits.data.fill_(5.7)
    q.eval()
    row = q.qrow(training=False).contiguous()
    words = pn.ops.quant_pack(x, row, 6)
    assert words.numel() == x.numel() * 6 // 32
    y = pn.ops.quant_unpack(words, x.numel(), row, 6)
    with torch.no_grad():
        assert torch.equal(y, q(x))
    codes = torch.round(y / row[0] + row[2])                      # exact lattice: (q - zp)*scale / scale + zp
    assert codes.min() >= 0 and codes.max() <= 63 and 62 <= torch.unique(codes).numel() <= 64


# ---- tables resident as integer codes ----------------------------------------------------------------------------------------
def test_packed_gather_kernel_matches_reference_eval_golden(pn, golden):
    """hash_fwd_packed_kernel on u8 codes == the live reference's quantised embedder in eval mode, bit for bit."""
    from oracle.fixtures import synthetic_tables
    g = golden("hash_embed_quant_eval")
    tables = [cu(t) for t in synthetic_tables(16, 15, salt=5)]
    x = cu(g["x"])
    res = [float(r) for r in O.level_resolutions(16, 512)]
    grid = pn.ops.make_grid(g["box_min"].tolist(), g["box_max"].tolist(), res, 15)
    for force_bits in (None, 12.0, 20.0):
        levels, rows = [], []
        for l in range(16):
            st = g["q%d" % l]
            soft = float(st[0]) if force_bits is None else force_bits
            bits, scale, zp, qmin, qmax = D.lbq_eval_params(soft, st[1], st[2], False)
            row = torch.tensor([scale, np.float32(scale + np.float32(1e-8)), zp, qmin, qmax, 1, 0, 0], dtype=torch.float32, device="cuda")
            rows.append(row)
            if bits <= 16:
                t = pn.ops.quant_codes(tables[l], row, 1 if bits <= 8 else 2)
                want = D.quant_codes(tables[l].cpu().numpy().reshape(-1), scale, zp, qmin, qmax)
                got = t.cpu().numpy().reshape(-1)
                assert np.array_equal(got.view(np.uint8 if bits <= 8 else np.uint16).astype(np.int64), want)
                words = pn.ops.quant_pack(tables[l], row, bits)
                assert torch.equal(pn.ops.quant_unpack_codes(words, tables[l].numel(), bits, 1 if bits <= 8 else 2).reshape(t.shape), t)
            else:
                words = pn.ops.quant_pack(tables[l], row, bits)
                t = pn.ops.quant_unpack(words, tables[l].numel(), row, bits).reshape(tables[l].shape)
            levels.append((t, scale, zp, qmin))
        packed = pn.ops.PackedLevels(levels)
        feat, keep = pn.ops.hash_encode_fwd_packed(grid, packed, x)
        if force_bits is None:
            assert np.array_equal(feat.cpu().numpy(), g["feat"])
        feat_fq, keep_fq = pn.ops.hash_encode_fwd(grid, tables, x, torch.stack(rows).contiguous())
        assert torch.equal(feat, feat_fq) and torch.equal(keep, keep_fq)
        assert packed.nbytes() < sum(t.numel() * 4 for t in tables) or force_bits == 20.0


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_packed_inference_renders_identically(pn, tmp_path, mode):
    """Eval-mode fake-quantised model == the same model gathering from integer codes (pack_for_inference) == the model
    loaded from its .pnq export with packed=True: identical embeddings and frames, in both arithmetic modes."""
    from indoor_nerf_b200 import quant_export
    sc, _, kw = small_model(pn, use_quantization=True)
    emb, nets = kw["embed_fn"], [kw["network_fn"], kw["network_fine"]]
    bits = [3.0, 4.2, 5.0, 6.0, 6.6, 7.0, 8.0, 8.0, 9.0, 10.0, 11.3, 12.0, 14.0, 16.0, 20.0, 26.0]
    x = (torch.rand(4096, 3, device="cuda") - 0.5) * 2.5
    for l, q in enumerate(emb.quantizers):
        q.calibrate(emb.embeddings[l].weight.detach())
        q.soft_bits.data.fill_(bits[l])
    emb.eval()
    for n in nets:
        n.sigma_weight_quantizer.calibrate(n.sigma_net[0].weight.detach())
        # a calibration range that straddles zero: with the reference's own calibration (min of a ReLU output = 0) the
        # zero point equals qmax and every hidden activation quantises to 0, which would make
        # the frames independent of the tables and this comparison vacuous
        n.sigma_act_quantizers[0].calibrate(torch.tensor([-0.5, 0.5], device="cuda"))
        n.eval()
    c2w = torch.from_numpy(sc["poses"][0][:3, :4]).cuda()

    def same_frame(a, b, what):
        if mode == "fp32":
            assert torch.equal(a, b), what
        else:
            # bf16 mode: the packed gather interpolates in the code domain and scales once; the fake-quant gather
            # multiplies by a reciprocal — both within fp32 rounding of the exact form before the bf16 rounding
            # (a ray whose last sample has sigma ~ 0 sits on raw2outputs' 1e10-interval discontinuity: allow 3 outliers)
            d = (a - b).abs()
            assert int((d > 2e-3).sum()) <= 3 and d.median().item() < 2e-5, (what, d.max().item(), d.median().item())

    with torch.no_grad():
        frame_exact = pn.render(sc["H"], sc["W"], sc["K"], chunk=512, c2w=c2w, **kw)[0]      # fp32 mode, fake-quant in the gather
    pn.set_mlp_mode(mode)
    try:
        with torch.no_grad():
            feat_q, keep_q = emb(x)
            frame_q = pn.render(sc["H"], sc["W"], sc["K"], chunk=512, c2w=c2w, **kw)[0]
            packed = emb.pack_for_inference()
            assert [t.dtype for t in packed.tensors] == [torch.uint8] * 8 + [torch.int16] * 6 + [torch.float32] * 2
            feat_p, keep_p = emb(x)
            frame_p = pn.render(sc["H"], sc["W"], sc["K"], chunk=512, c2w=c2w, **kw)[0]
        assert torch.equal(feat_p, feat_q) and torch.equal(keep_p, keep_q)
        same_frame(frame_p, frame_q, "codes vs fake-quant")
        if mode == "bf16":                                            # and both sit within the bf16 mode's bar of the exact frame
            for f in (frame_p, frame_q):
                assert (f - frame_exact).abs().median().item() < 2e-3
        emb.train()
        assert emb._packed is None                                   # the snapshot does not survive training
        emb.eval()
        path = str(tmp_path / "m.pnq")
        quant_export.export_quantized(path, emb, {"network_fn": nets[0], "network_fine": nets[1]})
        sc2, _, kw2 = small_model(pn, use_quantization=True)
        emb2, nets2 = kw2["embed_fn"], [kw2["network_fn"], kw2["network_fine"]]
        quant_export.load_quantized(path, emb2, {"network_fn": nets2[0], "network_fine": nets2[1]}, packed=True)
        [n.eval() for n in nets2]
        assert emb2._packed is not None and not emb2.training
        with torch.no_grad():
            feat_l, _ = emb2(x)
            frame_l = pn.render(sc["H"], sc["W"], sc["K"], chunk=512, c2w=c2w, **kw2)[0]
        assert torch.equal(feat_l, feat_q)
        same_frame(frame_l, frame_p, "loaded codes vs packed")
        if mode == "fp32":
            assert torch.equal(frame_l, frame_q)
    finally:
        pn.set_mlp_mode("fp32")
