"""Experiment: the unfused kernels of the bf16 path (hash gather / tcgen05 MLP; tcgen05 MLP backward / hash scatter)
run CONCURRENTLY on two streams over chunks of the point set, against the fused kernels.  Measures how much of the
latency-bound MLP round chain and the latency-bound gather / scatter the hardware scheduler overlaps when both kernels
are resident on every SM."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import indoor_nerf_b200 as pn  # noqa: E402
from indoor_nerf_b200 import ops, synthetic  # noqa: E402

pn.set_mlp_mode("bf16")
dev = torch.device("cuda", 0)
torch.manual_seed(0)
scene = synthetic.blender_scene(400, 400, n_views=100)
emb = pn.HashEmbedder(scene["bounding_box"], log2_hashmap_size=19, finest_resolution=512).to(dev)
with torch.no_grad():
    emb.table_storage.mul_(3000.0)
net = pn.NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, input_ch=32, input_ch_views=16).to(dev)
keys, weights = net.kernel_weights()
w = {k: t.detach().contiguous() for k, t in zip(keys, weights)}
rays, _ = synthetic.ray_batch(scene, 65536, seed=5, device=dev)
vd = (rays[1] / rays[1].norm(dim=-1, keepdim=True)).contiguous()
S = 192
z = torch.sort(2.0 + 4.0 * torch.rand(65536, S, device=dev), -1)[0]
pts = ops.make_points(rays[0], rays[1], z).reshape(-1, 3).contiguous()
P = pts.shape[0]
tables = [t.detach() for t in emb.tables()]
grid = emb.grid()
flat = torch.zeros(16, 1 << 19, 2, device=dev)
dt = list(flat.unbind(0))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
res = {}


def timed(fn, reps=6):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))


feat = torch.empty(P, 32, device=dev)
keep = torch.empty(P, dtype=torch.bool, device=dev)
out = torch.empty(P, 4, device=dev)
dout = torch.ones(P, 4, device=dev)
dfeat = torch.empty(P, 32, device=dev)

import ctypes
from indoor_nerf_b200._lib import call, dptr
from indoor_nerf_b200.ops import _ptr_array, _weights_struct, _mlp_input


def hash_fwd(lo, hi, st):
    call("pn_hash_encode_fwd", ctypes.byref(grid), _ptr_array(tables), None, dptr(pts[lo:hi]), hi - lo, dptr(feat[lo:hi]),
         dptr(keep[lo:hi], torch.bool), ctypes.c_void_p(st.cuda_stream))


def mlp_fwd(lo, hi, st):
    ws = _weights_struct(w)
    inp = _mlp_input(feat[lo:hi], None, vd, S, None, keep[lo:hi])
    inp.dirs = ctypes.c_void_p(vd.data_ptr() + (lo // S) * 12)
    call("pn_mlp_fwd_bf16", ctypes.byref(ws), ctypes.byref(inp), dptr(out[lo:hi]), ctypes.c_void_p(st.cuda_stream))


dw = {k: torch.zeros_like(v) for k, v in w.items()}


def mlp_bwd(lo, hi, st):
    ws, gs = _weights_struct(w), _weights_struct(dw)
    inp = _mlp_input(feat[lo:hi], None, vd, S, None, keep[lo:hi])
    inp.dirs = ctypes.c_void_p(vd.data_ptr() + (lo // S) * 12)
    call("pn_mlp_bwd_bf16", ctypes.byref(ws), ctypes.byref(inp), dptr(dout[lo:hi]), dptr(dfeat[lo:hi]), 32, None, 16,
         ctypes.byref(gs), ctypes.c_void_p(st.cuda_stream))


def hash_bwd(lo, hi, st):
    call("pn_hash_encode_bwd", ctypes.byref(grid), _ptr_array(dt), dptr(pts[lo:hi]), dptr(dfeat[lo:hi]), hi - lo,
         ctypes.c_void_p(st.cuda_stream))


cur = torch.cuda.current_stream()


def serial_fwd():
    hash_fwd(0, P, cur); mlp_fwd(0, P, cur)


def serial_bwd():
    mlp_bwd(0, P, cur); hash_bwd(0, P, cur)


def two_stream(first, second, n_chunks):
    step = (P // n_chunks) // (S * 128) * (S * 128)
    bounds = [(i * step, (i + 1) * step if i < n_chunks - 1 else P) for i in range(n_chunks)]
    evs = [torch.cuda.Event() for _ in bounds]

    def run():
        s1.wait_stream(cur); s2.wait_stream(cur)
        for (lo, hi), ev in zip(bounds, evs):
            first(lo, hi, s1)
            ev.record(s1)
            s2.wait_event(ev)
            second(lo, hi, s2)
        cur.wait_stream(s1); cur.wait_stream(s2)
    return run


with torch.no_grad():
    res["hash_fwd_alone_ms"] = timed(lambda: hash_fwd(0, P, cur))
    res["mlp_fwd_alone_ms"] = timed(lambda: mlp_fwd(0, P, cur))
    res["serial_fwd_ms"] = timed(serial_fwd)
    for n in (4, 8, 16):
        res["two_stream_fwd_%d_chunks_ms" % n] = timed(two_stream(hash_fwd, mlp_fwd, n))
    res["mlp_bwd_alone_ms"] = timed(lambda: mlp_bwd(0, P, cur))
    res["hash_bwd_alone_ms"] = timed(lambda: hash_bwd(0, P, cur))
    res["serial_bwd_ms"] = timed(serial_bwd)
    for n in (4, 8, 16):
        res["two_stream_bwd_%d_chunks_ms" % n] = timed(two_stream(mlp_bwd, hash_bwd, n))
print(json.dumps(res))
