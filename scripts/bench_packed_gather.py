"""Kernel-only timing (CUDA events, no host logic in the timed region) of the hash gather at the ScanNet-config table
size (log2_hashmap 22) in its three table formats: fp32 entries, fp32 entries fake-quantised in the gather, u8 codes.
Both the fp32-exact kernels (hash_fwd_kernel / hash_fwd_packed_kernel) and the fused tcgen05 field kernel.

    python scripts/bench_packed_gather.py [log2T]
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import ops, synthetic  # noqa: E402
from indoor_nerf_b200._lib import call, dptr, stream  # noqa: E402

log2T = int(sys.argv[1]) if len(sys.argv) > 1 else 22
dev = torch.device("cuda", 0)
scene = synthetic.blender_scene(400, 400, n_views=100)
rays, _ = synthetic.ray_batch(scene, 65536, seed=1, device=dev)
S = 192
z = torch.sort(2.0 + 4.0 * torch.rand(65536, S, device=dev), -1)[0]
pts = ops.make_points(rays[0], rays[1], z).reshape(-1, 3)
vd = rays[1] / rays[1].norm(dim=-1, keepdim=True)
P = pts.shape[0]
emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=log2T, finest_resolution=1024, use_quantization=True).to(dev)
for l, q in enumerate(emb.quantizers):
    q.calibrate(emb.embeddings[l].weight.detach())
emb.eval()
net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16).to(dev).eval()
keys, weights = net.kernel_weights()
w = {k: t.detach().contiguous() for k, t in zip(keys, weights)}
grid = emb.grid()
tables = [t.detach() for t in emb.tables()]
qrows = torch.stack([q.qrow(False) for q in emb.quantizers]).contiguous()
packed = emb.pack_for_inference()
out = torch.empty(P, 4, device=dev)
keep = torch.empty(P, dtype=torch.bool, device=dev)


def timeit(fn, reps=5):
    fn(); fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def fused(q):
    ws = ops._weights_struct(w)
    call("pn_field_fwd_bf16", ctypes.byref(grid), ops._ptr_array(tables), dptr(q, allow_none=True), ctypes.byref(ws),
         dptr(pts), dptr(vd), S, None, P, dptr(out), dptr(keep, torch.bool), None, stream())


res = {"log2T": log2T, "points": P,
       "table_bytes": {"fp32": 16 * (1 << log2T) * 8, "codes": packed.nbytes()},
       "hash_fwd_ms": {"fp32": timeit(lambda: ops.hash_encode_fwd(grid, tables, pts)),
                       "fake_quant": timeit(lambda: ops.hash_encode_fwd(grid, tables, pts, qrows)),
                       "u8_codes": timeit(lambda: ops.hash_encode_fwd_packed(grid, packed, pts))},
       "field_fwd_bf16_ms": {"fp32": timeit(lambda: fused(None)),
                             "fake_quant": timeit(lambda: fused(qrows)),
                             "u8_codes": timeit(lambda: ops.field_fwd_packed(grid, packed, w, pts, vd, S))}}
print(json.dumps(res))
