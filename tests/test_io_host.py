"""Host-side logic of the §8f components (RayBank, quant_export) on the CPU: ids, RNG consumption, epoch roll-over,
rank sharding and the .pnq file format, against the oracle's restatement of the reference's procedure.  The three
kernels involved are replaced by host builds of the same per-element source (tests/emu_ops.py)."""
import numpy as np
import pytest
import torch

import indoor_nerf_b200 as pn
from indoor_nerf_b200 import ops, quant_export, ray_bank
from oracle import dataio_oracle as D
from oracle import hashnerf_oracle as O
from tests import emu_ops


@pytest.fixture()
def emu_kernels(monkeypatch):
    monkeypatch.setattr(ops, "ray_bank_batch", emu_ops.ray_bank_batch)
    monkeypatch.setattr(ops, "quant_pack", emu_ops.quant_pack)
    monkeypatch.setattr(ops, "quant_unpack", emu_ops.quant_unpack)


def scene(golden):
    g = golden("ray_bank")
    return int(g["H"]), int(g["W"]), g["K"], g["poses"], g["images"], [int(i) for i in g["i_train"]], g


def test_ray_bank_use_batching_matches_reference_procedure(emu_kernels, golden):
    H, W, K, poses, images, i_train, g = scene(golden)
    bank = ray_bank.RayBank(H, W, K, poses, images, i_train, device="cpu")
    assert bank.n_rays == g["unshuffled"].shape[0]
    np.random.seed(0)
    bank.shuffle()
    rays_rgb = torch.from_numpy(g["shuffled"])                   # what the reference holds after run_nerf.py:907
    torch.manual_seed(7)
    ref_state = torch.get_rng_state()
    i_batch, N_rand = 0, 48                                       # 140 rays: batches of 48, 48, 44, then a new epoch
    for it in range(7):
        torch.set_rng_state(ref_state)
        rays, tgt = bank.next_batch(N_rand)
        ours_state = torch.get_rng_state()
        torch.set_rng_state(ref_state)
        batch = torch.transpose(rays_rgb[i_batch:i_batch + N_rand], 0, 1)              # run_nerf.py:962-966
        assert torch.equal(rays, batch[:2]) and torch.equal(tgt, batch[2]), "iteration %d" % it
        i_batch += N_rand
        if i_batch >= rays_rgb.shape[0]:
            rays_rgb = rays_rgb[torch.randperm(rays_rgb.shape[0])]                     # :968-973
            i_batch = 0
        ref_state = torch.get_rng_state()
        assert torch.equal(ours_state, ref_state)                                       # same RNG consumption
    assert bank.bytes_resident()["reference_rays_rgb"] == 140 * 36


def test_ray_bank_uint8_images(emu_kernels, golden):
    H, W, K, poses, images, i_train, g = scene(golden)
    img8 = (images * 255).astype(np.uint8)
    bank = ray_bank.RayBank(H, W, K, poses, img8, i_train, device="cpu")
    np.random.seed(1)
    bank.shuffle()
    order = bank.order.numpy().copy()
    rays, tgt = bank.next_batch(140)
    expect = (img8 / 255.).astype(np.float32)                                           # load_blender.py:62
    flat = np.stack([expect[i] for i in i_train]).reshape(-1, 3)
    assert (tgt.numpy() == flat[order]).all()
    assert (rays.numpy()[1] == g["unshuffled"][order, 1]).all()


@pytest.mark.parametrize("precrop", [None, 0.5])
def test_ray_bank_no_batching_matches_reference_procedure(emu_kernels, golden, precrop):
    H, W, K, poses, images, i_train, g = scene(golden)
    bank = ray_bank.RayBank(H, W, K, poses, images, i_train, device="cpu")
    N_rand = 3 if precrop else 6                                  # the 7x5 golden scene's central crop is 2x2
    np.random.seed(3)
    rays, tgt, img_i = bank.sample_image(N_rand, precrop_frac=precrop)
    ours_pos = np.random.get_state()[1][:8].tolist(), np.random.get_state()[2]
    # the reference's branch (run_nerf.py:976-1004) with the oracle's get_rays
    np.random.seed(3)
    ref_i = np.random.choice(i_train)
    assert ref_i == img_i
    ro, rd = O.get_rays(H, W, K, torch.from_numpy(poses[ref_i, :3, :4]))
    if precrop is not None:
        dH, dW = int(H // 2 * precrop), int(W // 2 * precrop)
        coords = torch.stack(torch.meshgrid(torch.linspace(H // 2 - dH, H // 2 + dH - 1, 2 * dH),
                                            torch.linspace(W // 2 - dW, W // 2 + dW - 1, 2 * dW), indexing="ij"), -1)
    else:
        coords = torch.stack(torch.meshgrid(torch.linspace(0, H - 1, H), torch.linspace(0, W - 1, W), indexing="ij"), -1)
    coords = torch.reshape(coords, [-1, 2])
    sel = np.random.choice(coords.shape[0], size=[N_rand], replace=False)
    sc = coords[sel].long()
    assert torch.equal(rays[0], ro[sc[:, 0], sc[:, 1]]) and torch.equal(rays[1], rd[sc[:, 0], sc[:, 1]])
    assert torch.equal(tgt, torch.from_numpy(images[ref_i])[sc[:, 0], sc[:, 1]])
    assert (np.random.get_state()[1][:8].tolist(), np.random.get_state()[2]) == ours_pos          # same numpy RNG consumption


def test_ray_bank_rank_shards_partition_the_global_batch(emu_kernels, golden):
    H, W, K, poses, images, i_train, g = scene(golden)
    full = ray_bank.RayBank(H, W, K, poses, images, i_train, device="cpu")
    np.random.seed(0)
    full.shuffle()
    want_r, want_t = full.next_batch(50)
    parts = []
    for r in range(3):
        b = ray_bank.RayBank(H, W, K, poses, images, i_train, device="cpu")
        np.random.seed(0)
        b.shuffle()
        b.rank, b.world = r, 3
        b._sync_order = lambda: None                              # no process group in this test
        parts.append(b.next_batch(50))
    assert torch.equal(torch.cat([p[0] for p in parts], 1), want_r)
    assert torch.equal(torch.cat([p[1] for p in parts], 0), want_t)
    assert [p[1].shape[0] for p in parts] == [16, 17, 17]


# ---- .pnq export ------------------------------------------------------------------------------------------------------
def quantised_model(bits_per_level):
    torch.manual_seed(0)
    box = (torch.tensor([-1.5] * 3), torch.tensor([1.5] * 3))
    emb = pn.HashEmbedder(box, log2_hashmap_size=8, use_quantization=True)
    net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16,
                       use_quantization=True)
    with torch.no_grad():
        emb.table_storage.mul_(50.0)
    for l, q in enumerate(emb.quantizers):
        q.calibrate(emb.embeddings[l].weight.detach())
        q.soft_bits.data.fill_(bits_per_level[l])
    net.sigma_weight_quantizer.calibrate(net.sigma_net[0].weight.detach())
    net.sigma_weight_quantizer.soft_bits.data.fill_(6.3)
    return emb.eval(), net.eval()


def test_pnq_roundtrip_equals_eval_fake_quant(emu_kernels, tmp_path):
    bits = [2.0, 3.2, 4.0, 5.5, 6.0, 7.0, 8.0, 9.4, 10.0, 11.0, 12.0, 13.0, 16.0, 20.0, 24.0, 27.0]
    emb, net = quantised_model(bits)
    path = str(tmp_path / "model.pnq")
    header = quant_export.export_quantized(path, emb, {"network_fn": net}, extra={"iter": 7})
    metas = {m["name"]: m for m in header["tensors"]}
    assert metas["embed_fn.embeddings.0.weight"]["bits"] == 2 and metas["embed_fn.embeddings.7.weight"]["bits"] == 9
    assert metas["embed_fn.embeddings.15.weight"]["storage"] == "fp32"                     # 27 bits > 24: stored as fp32
    assert metas["network_fn.sigma_net.0.weight"]["bits"] == 6
    packed_bits = sum(m["bits"] for m in header["tensors"] if m["storage"] == "packed" and m["name"].startswith("embed_fn"))
    assert header["bytes"]["payload"] < header["bytes"]["fp32_equivalent"]
    assert packed_bits == sum(int(round(b)) for b in bits[:15])
    # expected values: the quantisers' own eval forward == the oracle's codes (pinned to the reference by test_oracle_io)
    with torch.no_grad():
        want_tables = [emb.quantizers[l](emb.embeddings[l].weight) for l in range(16)]
        want_w0 = net.sigma_weight_quantizer(net.sigma_net[0].weight)
    q0 = emb.quantizers[3]
    _, scale, zp, qmin, qmax = D.lbq_eval_params(float(q0.soft_bits), float(q0.range_scale), float(q0.v_max), False)
    x3 = emb.embeddings[3].weight.detach().numpy().reshape(-1)
    assert (D.dequant_codes(D.quant_codes(x3, scale, zp, qmin, qmax), scale, zp, qmin) == want_tables[3].numpy().reshape(-1)).all()

    emb2, net2 = quantised_model([8.0] * 16)
    with torch.no_grad():
        emb2.table_storage.zero_()
        net2.sigma_net[0].weight.zero_()
    h2 = quant_export.load_quantized(path, emb2, {"network_fn": net2})
    assert h2["extra"] == {"iter": 7} and emb2._is_flat()
    for l in range(16):
        assert torch.equal(emb2.embeddings[l].weight.detach(), want_tables[l]), "level %d" % l
    assert torch.equal(net2.sigma_net[0].weight.detach(), want_w0)
    assert emb2.use_quantization is False and net2.sigma_weight_quantizer is None and net2.sigma_act_quantizers is not None
    assert torch.equal(net2.color_net[1].weight, net.color_net[1].weight)
    keys, tensors = net2.kernel_weights()
    assert torch.equal(tensors[0], want_w0)                                                  # not quantised a second time


def test_pnq_unquantised_model_is_stored_fp32(emu_kernels, tmp_path):
    box = (torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3))
    emb = pn.HashEmbedder(box, log2_hashmap_size=6)
    path = str(tmp_path / "plain.pnq")
    header = quant_export.export_quantized(path, emb)
    assert not header["table_quantisation"] and all(m["storage"] == "fp32" for m in header["tensors"])
    emb2 = pn.HashEmbedder(box, log2_hashmap_size=6)
    quant_export.load_quantized(path, emb2)
    assert torch.equal(emb2.table_storage, emb.table_storage)
    with pytest.raises(ValueError):
        open(path, "r+b").write(b"XXXX")
        quant_export.read_quantized(path, "cpu")


def test_ray_bank_edge_cases(emu_kernels, golden):
    H, W, K, poses, images, i_train, g = scene(golden)
    bank = ray_bank.RayBank(H, W, K, poses, images, i_train, device="cpu")
    np.random.seed(0)
    bank.shuffle()
    # a batch larger than the bank: the reference's slice is simply short, and the epoch rolls over
    torch.manual_seed(0)
    rays, tgt = bank.next_batch(1000)
    assert rays.shape == (2, 140, 3) and tgt.shape == (140, 3) and bank.i_batch == 0
    assert (rays.numpy()[1] == g["shuffled"][:, 1]).all()
    # more ranks than rays in the (short) batch: some shards are empty, the rest still tile it
    np.random.seed(0)
    parts = []
    for r in range(8):
        b = ray_bank.RayBank(H, W, K, poses, images, i_train, device="cpu")
        np.random.seed(0)
        b.shuffle()
        b.rank, b.world, b._sync_order = r, 8, (lambda: None)
        parts.append(b.next_batch(5)[0])
    assert sorted(p.shape[1] for p in parts) == [0, 0, 0, 1, 1, 1, 1, 1]
    assert (torch.cat(parts, 1).numpy()[1] == g["shuffled"][:5, 1]).all()
    # images must already be RGB
    with pytest.raises(ValueError):
        ray_bank.RayBank(H, W, K, poses, np.zeros((5, H, W, 4), np.float32), i_train, device="cpu")


def test_pnq_tensor_not_multiple_of_32_is_stored_fp32(emu_kernels, tmp_path):
    """A level whose element count is not a multiple of 32 cannot be bit-packed in 32-value groups: it is stored as the
    eval fake-quantised fp32 values instead, and still loads to exactly those values."""
    box = (torch.tensor([-1.0] * 3), torch.tensor([1.0] * 3))
    emb = pn.HashEmbedder(box, log2_hashmap_size=3, use_quantization=True)          # 8 rows x 2 = 16 values per level
    with torch.no_grad():
        emb.table_storage.mul_(100.0)
    for l, q in enumerate(emb.quantizers):
        q.calibrate(emb.embeddings[l].weight.detach())
    emb.eval()
    with torch.no_grad():
        want = [emb.quantizers[l](emb.embeddings[l].weight) for l in range(16)]
    path = str(tmp_path / "tiny.pnq")
    header = quant_export.export_quantized(path, emb)
    assert header["table_quantisation"] and all(m["storage"] == "fp32" for m in header["tensors"])
    emb2 = pn.HashEmbedder(box, log2_hashmap_size=3, use_quantization=True)
    quant_export.load_quantized(path, emb2)
    assert all(torch.equal(emb2.embeddings[l].weight.detach(), want[l]) for l in range(16))
    assert emb2.use_quantization is False
