"""TEST INFRASTRUCTURE — host stand-ins for three C-ABI entry points, built from the SAME per-element source the
kernels compile (csrc/io_core.cuh via tests/hostemu).  They let the CPU suite exercise the host-side logic of
RayBank and quant_export (id arithmetic, RNG consumption, sharding, file format) without a GPU by
monkeypatching ``ops.ray_bank_batch`` / ``ops.quant_pack`` / ``ops.quant_unpack`` inside a test.  Never imported by
the product."""
import ctypes
import os
import subprocess

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "hostemu")
_lib = None


def build():
    """Compile tests/hostemu/hostemu.cpp if the .so is missing or older than its sources.  Written to a private
    temporary file and renamed into place, so concurrent callers (two gloo ranks, xdist workers) never load a
    half-written library."""
    import glob
    so = os.path.join(EMU_DIR, "libhostemu.so")
    csrc = os.path.join(os.path.dirname(HERE), "indoor-nerf_b200", "csrc")
    srcs = [os.path.join(EMU_DIR, "hostemu.cpp"), os.path.join(os.path.dirname(HERE), "include", "pocketnerf.h")] + \
        glob.glob(os.path.join(csrc, "*.cuh"))
    if os.path.isfile(so) and os.path.getmtime(so) >= max(os.path.getmtime(f) for f in srcs):
        return so
    tmp = "%s.%d.tmp" % (so, os.getpid())
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-std=c++17",
                           "-I/usr/local/cuda/include", "-Wno-unknown-pragmas", "-o", tmp,
                           os.path.join(EMU_DIR, "hostemu.cpp")])
    os.replace(tmp, so)
    return so


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.emu_ssim_sum.restype = ctypes.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ray_bank_batch(ids, H, W, K, poses, image_index=None, images=None, f64_dirs=True):
    ids_np = np.ascontiguousarray(ids.cpu().numpy(), np.int64)
    B = ids_np.size
    poses_np = np.ascontiguousarray(poses.cpu().numpy(), np.float32)
    Kd = np.ascontiguousarray(np.asarray(K, np.float64))
    idx = np.ascontiguousarray(image_index.cpu().numpy(), np.int32) if image_index is not None else None
    img = images.cpu()
    img_np = np.ascontiguousarray((img.double() / 255.).float().numpy() if img.dtype == torch.uint8 else img.numpy(), np.float32)
    rays, tgt = np.zeros((2, B, 3), np.float32), np.zeros((B, 3), np.float32)
    lib().emu_ray_bank(_p(ids_np), ctypes.c_int64(B), int(H), int(W), _p(Kd), _p(poses_np),
                       ctypes.c_int64(poses_np.shape[1] * poses_np.shape[2]), _p(idx) if idx is not None else None,
                       _p(img_np), int(bool(f64_dirs)), _p(rays), _p(tgt))
    return torch.from_numpy(rays), torch.from_numpy(tgt)


def quant_pack(x, qrow, bits):
    xn = np.ascontiguousarray(x.detach().cpu().numpy(), np.float32).reshape(-1)
    row = np.ascontiguousarray(qrow.detach().cpu().numpy(), np.float32)
    words = np.zeros(xn.size * int(bits) // 32, np.uint32)
    lib().emu_quant_pack(_p(xn), ctypes.c_int64(xn.size), _p(row), int(bits), _p(words))
    return torch.from_numpy(words.view(np.int32))


def quant_unpack(words, n, qrow, bits):
    w = np.ascontiguousarray(words.cpu().numpy()).view(np.uint32)
    row = np.ascontiguousarray(qrow.detach().cpu().numpy(), np.float32)
    x = np.zeros(n, np.float32)
    lib().emu_quant_unpack(_p(w), ctypes.c_int64(n), _p(row), int(bits), _p(x))
    return torch.from_numpy(x)
