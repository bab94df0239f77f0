"""CPU-side checks of the host layer: the C-ABI library loads and exports what include/pocketnerf.h
declares, the drop-in modules keep the reference's names / state_dict keys / parameter order, the torch-side
pieces (RAdam, TV loss) match the golden vectors, and nothing silently runs on the CPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import indoor_nerf_b200 as pn
from indoor_nerf_b200 import _lib, loss as ploss, model as pmodel, radam as pradam, utils as putils
from oracle.fixtures import synthetic_tables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = torch.from_numpy


def test_library_exports_header_symbols():
    header = open(os.path.join(ROOT, "include", "pocketnerf.h")).read()
    declared = set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", header))
    declared -= {"pn_stream_t"}
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), "libpocketnerf.so does not export %s" % name
    assert set(_lib.exported_symbols()) <= declared | {"pn_abi_version", "pn_last_error", "pn_launch_count"}
    assert declared == set(_lib.exported_symbols()), "binding and header disagree: %s" % (
        declared ^ set(_lib.exported_symbols()))
    assert _lib.lib().pn_abi_version() == int(re.search(r"#define PN_ABI_VERSION (\d+)", header).group(1))


def test_no_cpu_path():
    emb = pn.HashEmbedder((torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3)), log2_hashmap_size=8)
    with pytest.raises(_lib.PocketNerfError):
        emb(torch.zeros(4, 3))
    net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16)
    with pytest.raises(_lib.PocketNerfError):
        net(torch.zeros(4, 48))
    with pytest.raises(_lib.PocketNerfError):
        pn.raw2outputs(torch.zeros(2, 8, 4), torch.zeros(2, 8), torch.zeros(2, 3))


def test_bad_arguments_are_reported():
    lib = _lib.lib()
    g = _lib.HashGrid()
    g.n_levels, g.log2_hashmap_size = 99, 19
    rc = lib.pn_hash_encode_fwd(ctypes.byref(g), None, None, None, 0, None, None, None)
    assert rc == -3 and b"n_levels" in lib.pn_last_error()
    rc = lib.pn_sort_merge(None, 1, None, 1, 1, None, None)
    assert rc == -1 and b"NULL" in lib.pn_last_error()


def test_hash_embedder_layout_and_keys():
    box = (torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3))
    emb = pn.HashEmbedder(box, log2_hashmap_size=10, use_quantization=True)
    keys = list(emb.state_dict().keys())
    assert keys[:16] == ["embeddings.%d.weight" % l for l in range(16)]
    assert keys[16:21] == ["quantizers.0.soft_bits", "quantizers.0.range_scale", "quantizers.0.v_max",
                           "quantizers.0.running_min", "quantizers.0.running_max"]
    params = list(emb.parameters())
    assert len(params) == 16 + 48 and all(p.shape == (1024, 2) for p in params[:16])
    assert emb._is_flat() and emb.out_dim == 32 and emb.warmup_steps == 500 and emb.current_step == 0
    assert float(emb.embeddings[3].weight.abs().max()) <= 1e-4
    # state_dict round trip keeps the single-buffer layout; double() + float() (an _apply) re-flattens
    sd = {k: v.clone() for k, v in emb.state_dict().items()}
    emb2 = pn.HashEmbedder(box, log2_hashmap_size=10, use_quantization=True)
    emb2.load_state_dict(sd)
    assert emb2._is_flat() and torch.equal(emb2.table_storage, emb.table_storage)
    emb2.double().float()
    assert emb2._is_flat() and torch.equal(emb2.table_storage, emb.table_storage)
    # the TV loss indexes a level with an int64 tensor of any shape
    idx = torch.randint(0, 1024, (3, 4, 5))
    assert emb.embeddings[2](idx).shape == (3, 4, 5, 2)
    # resolutions (float32 knife edges)
    assert [float(r) for r in emb.level_resolutions()] == [16, 20, 25, 32, 40, 50, 64, 80, 101, 128, 161, 203, 256,
                                                           322, 406, 512]


def test_nerf_small_keys_match_reference_names():
    net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32,
                       input_ch_views=16, use_quantization=True, predict_normals=True)
    keys = set(net.state_dict().keys())
    assert {"sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight",
            "color_net.2.weight", "normal_net.0.weight", "normal_net.0.bias", "normal_net.2.weight",
            "normal_net.2.bias", "sigma_act_quantizers.0.soft_bits", "sigma_weight_quantizer.soft_bits"} <= keys
    assert net.sigma_net[0].weight.shape == (64, 32) and net.color_net[0].weight.shape == (64, 31)
    assert sum(p.numel() for n, p in net.named_parameters() if "quantizer" not in n and "normal" not in n) == 9344


@pytest.mark.live_reference
def test_state_dict_keys_equal_live_reference():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not present")
    ref = ref_shim.load()
    box = (torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3))
    for q in (False, True):
        a = ref.hash_encoding.HashEmbedder(box, log2_hashmap_size=8, use_quantization=q)
        b = pn.HashEmbedder(box, log2_hashmap_size=8, use_quantization=q)
        assert list(a.state_dict().keys()) == list(b.state_dict().keys())
        assert [tuple(p.shape) for p in a.parameters()] == [tuple(p.shape) for p in b.parameters()]
        for normals in (False, True):
            ra = ref_shim.make_nerf_small(ref, predict_normals=normals, use_quantization=q)
            rb = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32,
                              input_ch_views=16, use_quantization=q, predict_normals=normals)
            assert list(ra.state_dict().keys()) == list(rb.state_dict().keys())
            assert [tuple(p.shape) for p in ra.parameters()] == [tuple(p.shape) for p in rb.parameters()]


def test_quantizer_module_matches_oracle():
    from oracle import hashnerf_oracle as O
    x = torch.randn(64, 32) * 0.1
    for sym in (True, False):
        q = pn.LearnedBitwidthQuantizer(init_bits=6.3, symmetric=sym).train()
        y = q(x)                                       # calibrates
        rs, vmax, _, _ = O.lbq_calibrate(x, sym)
        for training in (True, False):
            q.train(training)
            want = O.lbq_apply(x, *O.lbq_scalars(q.soft_bits.data, rs, vmax, sym, training), training)
            assert torch.equal(q(x), want)
        row = q.qrow(training=True)
        assert row.shape == (8,) and float(row[5]) == 1.0


def test_radam_golden(golden):
    g = golden("radam")
    p = [torch.nn.Parameter(T(g["init0"].copy())), torch.nn.Parameter(T(g["init1"].copy()))]
    opt = pradam.RAdam([{"params": [p[0]], "weight_decay": 1e-6}, {"params": [p[1]], "eps": 1e-15}], lr=5e-4,
                       betas=(0.9, 0.99))
    for it in range(8):
        for i in range(2):
            p[i].grad = T(g["g%d_%d" % (it, i)].copy())
        opt.step()
        for i in range(2):
            np.testing.assert_allclose(p[i].detach().numpy(), g["p%d_%d" % (it, i)], rtol=2e-6, atol=1e-9)
    assert set(opt.state[p[0]].keys()) == {"step", "exp_avg", "exp_avg_sq"}


def test_tv_loss_golden(golden):
    g = golden("tv_loss")
    table = torch.nn.Embedding(1 << 19, 2, _weight=T(synthetic_tables(1, 19, salt=int(g["salt"]))[0]))
    for level in (0, 5, 15):
        torch.manual_seed(100 + level)
        v = ploss.total_variation_loss(table, torch.tensor(16), torch.tensor(512), level, 19, n_levels=16)
        np.testing.assert_allclose(float(v), float(g["tv_%d" % level]), rtol=1e-6)


def test_utils_hash_cpu_formula(golden):
    g = golden("hash_primitives")
    for k in (12, 19, 22):
        assert (putils.hash(T(g["corners"]), k).numpy() == g["h%d" % k]).all()


def test_create_nerf_structure():
    a = pmodel.default_args(bounding_box=(torch.tensor([-2.0] * 3), torch.tensor([2.0] * 3)), log2_hashmap_size=8)
    kw_train, kw_test, start, grad_vars, opt = pmodel.create_nerf(a, device="cpu")
    assert set(kw_train) == {"network_query_fn", "perturb", "N_importance", "network_fine", "N_samples", "network_fn",
                             "embed_fn", "use_viewdirs", "white_bkgd", "raw_noise_std", "predict_normals", "ndc", "lindisp"}
    assert kw_test["perturb"] is False and kw_test["raw_noise_std"] == 0.
    assert len(grad_vars) == 10 and start == 0
    assert len(opt.param_groups) == 2 and opt.param_groups[0]["weight_decay"] == 1e-6 and opt.param_groups[1]["eps"] == 1e-15
    assert len(opt.param_groups[1]["params"]) == 16


def test_checkpoint_roundtrip(tmp_path):
    a = pmodel.default_args(bounding_box=(torch.tensor([-2.0] * 3), torch.tensor([2.0] * 3)), log2_hashmap_size=8)
    kw, _, _, _, opt = pmodel.create_nerf(a, device="cpu")
    d = tmp_path / "exp"
    d.mkdir()
    pmodel.save_checkpoint(str(d / "000123.tar"), 123, kw, opt)
    a2 = pmodel.default_args(bounding_box=a.bounding_box, log2_hashmap_size=8, basedir=str(tmp_path), expname="exp",
                             no_reload=False)
    kw2, _, start, _, _ = pmodel.create_nerf(a2, device="cpu")
    assert start == 123
    assert torch.equal(kw2["embed_fn"].table_storage, kw["embed_fn"].table_storage)
    assert torch.equal(kw2["network_fine"].color_net[1].weight, kw["network_fine"].color_net[1].weight)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libpocketnerf.so"))
    with pytest.raises(_lib.PocketNerfError, match="there is no fallback"):
        _lib.lib()


def test_table_grad_buffer_semantics():
    """The flat table-gradient buffer (ops.table_grad_buffer): installed as .grad views, re-zeroed after
    zero_grad(set_to_none), keeps what other autograd producers already delivered, accumulates across producers."""
    from indoor_nerf_b200 import ops
    emb = pn.HashEmbedder((torch.zeros(3), torch.ones(3)), log2_hashmap_size=6)
    tables = emb.tables()
    flat = ops.table_grad_buffer(tables)
    assert flat.shape == (16, 64, 2) and float(flat.abs().sum()) == 0.0
    assert all(t.grad.data_ptr() == flat[l].data_ptr() for l, t in enumerate(tables))
    flat += 1.0                                            # a backward kernel accumulating
    assert ops.table_grad_buffer(tables) is flat and float(tables[5].grad.sum()) == 128.0   # second producer: same buffer, kept
    # a framework producer (e.g. the reference's own per-level TV loss through embeddings[i](idx)) adds in place
    emb.embeddings[3](torch.tensor([1, 1, 7])).sum().backward()
    assert tables[3].grad.data_ptr() == flat[3].data_ptr()
    assert tables[3].grad[1].tolist() == [3.0, 3.0] and tables[3].grad[7].tolist() == [2.0, 2.0]
    # optimizer.zero_grad() -> None -> the next backward starts from zeros again
    for t in tables:
        t.grad = None
    assert float(ops.table_grad_buffer(tables).abs().sum()) == 0.0
    # a producer that ran BEFORE ours this step left a dense grad on one level: it is folded in, not lost
    for t in tables:
        t.grad = None
    emb.embeddings[2](torch.tensor([4])).sum().backward()
    flat = ops.table_grad_buffer(tables)
    assert tables[2].grad.data_ptr() == flat[2].data_ptr() and tables[2].grad[4].tolist() == [1.0, 1.0]
    assert float(flat.sum()) == 2.0
    # and the flat view used by the optimiser / all-reduce is zero-copy
    from indoor_nerf_b200 import parallel
    assert parallel.flat_table_grad(emb).data_ptr() == flat.data_ptr()


@pytest.mark.live_reference
def test_patch_installs_renderer_into_live_reference():
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip("reference not present")
    ref = ref_shim.load()
    saved = {n: getattr(ref.run_nerf, n) for n in ("batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays")}
    try:
        pn.patch(ref.run_nerf)
        assert ref.run_nerf.render_rays is pn.render_rays and ref.run_nerf.raw2outputs is pn.raw2outputs
        assert ref.run_nerf.run_network is pn.run_network
        # signatures the driver relies on (run_nerf.py:1007, :173)
        import inspect
        for n, fn in saved.items():
            a, b = inspect.signature(fn).parameters, inspect.signature(getattr(pn, n)).parameters
            assert list(a) == list(b), (n, list(a), list(b))
            assert [p.default for p in a.values()] == [p.default for p in b.values()], n
    finally:
        for n, fn in saved.items():
            setattr(ref.run_nerf, n, fn)


def test_batched_quantiser_rows_equal_the_per_quantiser_rows():
    """hash_encoding._quant_rows builds all levels' kernel rows with vectorised ops; they must equal the rows each
    LearnedBitwidthQuantizer computes for itself (quantization.py:121-175), bit for bit, in both forms."""
    from indoor_nerf_b200.quantization import qrows_batched
    torch.manual_seed(3)
    for symmetric in (False, True):
        qs = []
        for l in range(16):
            q = pn.LearnedBitwidthQuantizer(init_bits=8.0, min_bits=2.0, max_bits=32.0, symmetric=symmetric)
            q.calibrate(torch.randn(1000) * (10.0 ** (-l / 3.0)))
            q.soft_bits.data.fill_(float(torch.empty(()).uniform_(1.0, 33.0)))
            qs.append(q)
        for training in (True, False):
            want = torch.stack([q.qrow(training) for q in qs])
            got = qrows_batched(qs, training)
            assert got.shape == (16, 8) and torch.equal(got, want), (symmetric, training)
    mixed = [pn.LearnedBitwidthQuantizer(symmetric=True), pn.LearnedBitwidthQuantizer(symmetric=False)]
    assert torch.equal(qrows_batched(mixed, False), torch.stack([q.qrow(False) for q in mixed]))


def test_png_writer_roundtrip(tmp_path):
    """render_path saves plain PNGs (8-bit RGB / gray, one IDAT chunk, filter 0): decode by hand and compare."""
    import struct
    import zlib
    from indoor_nerf_b200.render import _write_png
    rs = np.random.RandomState(0)
    for shape in ((5, 7, 3), (9, 4)):
        img = rs.randint(0, 256, shape).astype(np.uint8)
        path = str(tmp_path / ("x%d.png" % len(shape)))
        _write_png(path, img)
        raw = open(path, "rb").read()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        w, h, depth, color = struct.unpack(">IIBB", raw[16:26])
        assert (h, w, depth, color) == (shape[0], shape[1], 8, 2 if len(shape) == 3 else 0)
        pos, chunks = 8, {}
        while pos < len(raw):
            n, tag = struct.unpack(">I", raw[pos:pos + 4])[0], raw[pos + 4:pos + 8]
            data = raw[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == (zlib.crc32(tag + data) & 0xffffffff)
            chunks[tag] = data
            pos += 12 + n
        assert set(chunks) == {b"IHDR", b"IDAT", b"IEND"}
        rows = np.frombuffer(zlib.decompress(chunks[b"IDAT"]), np.uint8).reshape(shape[0], -1)
        assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(shape), img)


def test_trainer_learning_rate_schedule():
    """Trainer.decay_learning_rate follows run_nerf.py:1289-1293: evaluated at the end of the iteration whose
    global_step is g (before `global_step += 1`, :1475), it sets lrate * 0.1 ** (g / (lrate_decay * 1000))."""
    from types import SimpleNamespace
    from indoor_nerf_b200.trainer import Trainer
    p = torch.nn.Parameter(torch.zeros(3))
    opt = pradam.RAdam([{"params": [p], "weight_decay": 1e-6}, {"params": [torch.nn.Parameter(torch.zeros(2))], "eps": 1e-15}],
                       lr=0.01, betas=(0.9, 0.99))
    tr = Trainer.__new__(Trainer)
    tr.args, tr.opt, tr.step_idx = SimpleNamespace(lrate=0.01, lrate_decay=10), opt, 0
    for g in (0, 499, 9999):
        tr.decay_learning_rate(g)
        want = 0.01 * (0.1 ** (g / 10000))
        assert all(grp["lr"] == want for grp in opt.param_groups)
    tr.args = SimpleNamespace(lrate=0.01)
    tr.decay_learning_rate(5)                                     # no lrate_decay: left alone
    assert opt.param_groups[0]["lr"] == 0.01 * (0.1 ** (9999 / 10000))


def test_trainer_resumes_schedule_from_checkpoint_step():
    """A Trainer built with create_nerf's `start` (the checkpoint's global_step, run_nerf.py:299-307) continues the
    schedule instead of restarting it: the learning rate is the reference's lrate * 0.1 ** (start / decay_steps) — what
    the checkpointed optimiser state holds, since the reference saves after the update (:1289-1293 before :1345) — and
    the TV term is cut for good after the first resumed iteration (`if i > 1000`, :1036-1037).  A Trainer without
    `start` would reset lr to lrate * 0.1 ** 0 on its first step and re-enable TV for 1000 iterations."""
    from types import SimpleNamespace
    from indoor_nerf_b200.trainer import Trainer
    params = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(2))]
    opt = pradam.RAdam([{"params": [params[0]], "weight_decay": 1e-6}, {"params": [params[1]], "eps": 1e-15}],
                       lr=5e-4, betas=(0.9, 0.99))
    args = SimpleNamespace(lrate=5e-4, lrate_decay=10, tv_loss_weight=1e-6, sparse_loss_weight=1e-10)
    kw = {"embed_fn": SimpleNamespace(quantizers=None), "network_fn": SimpleNamespace(), "network_fine": None}
    tr = Trainer(args, kw, opt, 4, 4, None, 2.0, 6.0, start=40000)
    want = 5e-4 * (0.1 ** (40000 / 10000))                        # 5e-8, not 5e-4
    assert tr.step_idx == 40000 and all(g["lr"] == want for g in opt.param_groups)
    fresh = Trainer(args, kw, opt, 4, 4, None, 2.0, 6.0)
    assert fresh.step_idx == 0 and all(g["lr"] == want for g in opt.param_groups)     # start=0 leaves the optimiser alone
    # the counters after one (simulated) resumed step: lr(40000) again (the reference repeats global_step = start), TV off
    tr.step_idx += 1
    if tr.step_idx > 1000:
        tr.tv_weight = 0.0
    tr.decay_learning_rate(tr.step_idx - 1)
    assert tr.tv_weight == 0.0 and all(g["lr"] == want for g in opt.param_groups)


def test_no_cpu_path_for_the_io_ops():
    """The §8f entry points refuse host tensors like the rest of the path (no silent CPU fallback)."""
    from indoor_nerf_b200 import ops
    K = np.array([[10., 0, 2], [0, 10., 2], [0, 0, 1]])
    ids = torch.arange(4)
    with pytest.raises(_lib.PocketNerfError):
        ops.ray_bank_batch(ids, 4, 4, K, torch.zeros(1, 3, 4), None, torch.zeros(1, 4, 4, 3))
    with pytest.raises(_lib.PocketNerfError):
        ops.image_sqerr(torch.zeros(8, 8, 3), torch.zeros(8, 8, 3))
    with pytest.raises(_lib.PocketNerfError):
        ops.image_ssim(torch.zeros(8, 8, 3), torch.zeros(8, 8, 3))
    with pytest.raises(_lib.PocketNerfError):
        ops.to8b(torch.zeros(8))
    row = torch.tensor([1e-3, 1e-3, 0, 0, 255, 1, 0, 0])
    with pytest.raises(_lib.PocketNerfError):
        ops.quant_pack(torch.zeros(64), row, 8)
    with pytest.raises(_lib.PocketNerfError):
        ops.quant_codes(torch.zeros(64), row, 1)
    with pytest.raises(_lib.PocketNerfError):
        ops.PackedLevels([(torch.zeros(8, 2, dtype=torch.uint8), 1.0, 0.0, 0.0)])
    emb = pn.HashEmbedder((torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3)), log2_hashmap_size=5, use_quantization=True)
    with pytest.raises(RuntimeError):
        emb.pack_for_inference()                                     # quantisers not calibrated
    # bad arguments are reported through the C ABI's error channel, without a device
    lib = _lib.lib()
    assert lib.pn_quant_pack(None, 64, None, 8, None, None) == -1 and b"NULL" in lib.pn_last_error()
    assert lib.pn_image_ssim(None, None, 8, 8, 3, 1.0, None, None) == -1
    assert lib.pn_quant_unpack_codes(None, 64, 9, 1, None, None) == -1
