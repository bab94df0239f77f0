"""ctypes binding of libpocketnerf.so (include/pocketnerf.h).

There is deliberately no fallback: if the shared library has not been built, or a tensor is not a
contiguous CUDA tensor of the expected dtype, the call raises.  PyTorch is used for device memory and
streams only; every kernel launched through this module is ours.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# POCKETNERF_LIB selects another build of the same sources (tuning experiments: csrc/build.sh with PN_NVCC_EXTRA)
LIB_PATH = os.environ.get("POCKETNERF_LIB") or os.path.join(_HERE, "csrc", "libpocketnerf.so")
ABI_VERSION = 11
MAX_LEVELS = 16
QROW = 8


class PocketNerfError(RuntimeError):
    pass


class HashGrid(ctypes.Structure):
    _fields_ = [("box_min", ctypes.c_float * 3), ("box_max", ctypes.c_float * 3),
                ("resolution", ctypes.c_float * MAX_LEVELS), ("n_levels", ctypes.c_int32),
                ("log2_hashmap_size", ctypes.c_int32)]


class MlpWeights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("s0", "s1", "c0", "c1", "c2", "n0w", "n0b", "n2w", "n2b")]


class MlpInput(ctypes.Structure):
    _fields_ = [("feat", ctypes.c_void_p), ("feat_stride", ctypes.c_int64), ("sh", ctypes.c_void_p),
                ("sh_stride", ctypes.c_int64), ("dirs", ctypes.c_void_p), ("samples_per_ray", ctypes.c_int32),
                ("act_q", ctypes.c_void_p), ("keep", ctypes.c_void_p), ("n_points", ctypes.c_int64)]


class PackedTables(ctypes.Structure):
    _fields_ = [("codes", ctypes.c_void_p * MAX_LEVELS), ("entry_bytes", ctypes.c_int32 * MAX_LEVELS),
                ("scale", ctypes.c_float * MAX_LEVELS), ("zero_point", ctypes.c_float * MAX_LEVELS),
                ("qmin", ctypes.c_float * MAX_LEVELS)]


_P, _I, _L = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
_SIGNATURES = {
    "pn_hash_encode_fwd": [ctypes.POINTER(HashGrid), _P, _P, _P, _L, _P, _P, _P],
    "pn_hash_encode_bwd": [ctypes.POINTER(HashGrid), _P, _P, _P, _L, _P],
    "pn_hash_indices": [ctypes.POINTER(HashGrid), _P, _L, _P, _P],
    "pn_hash_coords": [_P, _L, _I, _I, _P, _P],
    "pn_hash_gather_minmax": [ctypes.POINTER(HashGrid), _P, _P, _L, _P, _P],
    "pn_sh_encode": [_P, _L, _P, _P],
    "pn_mlp_fwd": [ctypes.POINTER(MlpWeights), ctypes.POINTER(MlpInput), _P, _P],
    "pn_mlp_bwd": [ctypes.POINTER(MlpWeights), ctypes.POINTER(MlpInput), _P, _P, _L, _P, _L,
                   ctypes.POINTER(MlpWeights), _P],
    "pn_mlp_fwd_bf16": [ctypes.POINTER(MlpWeights), ctypes.POINTER(MlpInput), _P, _P],
    "pn_mlp_bwd_bf16": [ctypes.POINTER(MlpWeights), ctypes.POINTER(MlpInput), _P, _P, _L, _P, _L,
                        ctypes.POINTER(MlpWeights), _P],
    "pn_field_fwd_bf16": [ctypes.POINTER(HashGrid), _P, _P, ctypes.POINTER(MlpWeights), _P, _P, _I, _P, _L, _P, _P, _P, _P],
    "pn_field_bwd_bf16": [ctypes.POINTER(HashGrid), _P, ctypes.POINTER(MlpWeights), _P, _P, _P, _I, _P, _P, _P, _L,
                          ctypes.POINTER(MlpWeights), _P, _L, _P],
    "pn_tc_selftest": [_P, _P, _P, _P, _P],
    "pn_debug_timeline": [_P, _L],
    "pn_tv_loss_fwd": [_P, _I, _I, _P, _P, _P, _P],
    "pn_tv_loss_bwd": [_P, _P, _I, _I, _P, _P, _P, _P],
    "pn_radam_step": [_P, _P, _P, _P, _L] + [ctypes.c_float] * 5 + [_I, _P],
    "pn_table_fake_quant": [_P, _P, _I, _L, _P, _P],
    "pn_radam_step_dyn": [_P, _P, _P, _P, _L] + [ctypes.c_float] * 3 + [_P, _I, _P],
    "pn_store_floats": [_P, ctypes.POINTER(ctypes.c_float), _I, _P],
    "pn_composite_fwd": [_P, _I, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    "pn_composite_bwd": [_P, _I, _P, _P, _P, _L, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P],
    "pn_sample_pdf": [_P, _P, _L, _P, _L, _L, _I, _I, _P, _P, _P, _P],
    "pn_sample_from_cdf": [_P, _P, _P, _L, _L, _I, _I, _P, _P, _P],
    "pn_sort_merge": [_P, _I, _P, _I, _L, _P, _P],
    "pn_gen_rays": [_I, _I, _P, _P, _P, _P, _P],
    "pn_ndc_rays": [_I, _I, ctypes.c_double, ctypes.c_double, _P, _P, _L, _P, _P, _P],
    "pn_make_points": [_P, _L, _P, _L, _P, _L, _I, _P, _P],
    "pn_coarse_z": [_P, _P, _L, _P, _P, _L, _I, _I, _P, _P],
    "pn_ray_bank_batch": [_P, _L, _I, _I, _P, _P, _L, _P, _P, _I, _I, _P, _P, _P],
    "pn_image_sqerr": [_P, _P, _L, _P, _P],
    "pn_image_ssim": [_P, _P, _I, _I, _I, ctypes.c_double, _P, _P],
    "pn_to8b": [_P, _L, _P, _P],
    "pn_quant_pack": [_P, _L, _P, _I, _P, _P],
    "pn_quant_unpack": [_P, _L, _P, _I, _P, _P],
    "pn_quant_codes": [_P, _L, _P, _I, _P, _P],
    "pn_quant_unpack_codes": [_P, _L, _I, _I, _P, _P],
    "pn_hash_encode_fwd_packed": [ctypes.POINTER(HashGrid), ctypes.POINTER(PackedTables), _P, _L, _P, _P, _P],
    "pn_field_fwd_bf16_packed": [ctypes.POINTER(HashGrid), ctypes.POINTER(PackedTables), ctypes.POINTER(MlpWeights), _P, _P,
                                 _I, _P, _L, _P, _P, _P],
}
# entry points added by later kernels (fused field, tensor-core MLP, optimizer); bound when exported
_OPTIONAL = {}

_lib = None


# entry points that do not return an error code
_PLAIN = {"pn_field_bwd_workspace_bytes": ([], ctypes.c_int64)}


def exported_symbols():
    """Names every build of the library must export (tests check them against include/pocketnerf.h)."""
    return ["pn_abi_version", "pn_last_error", "pn_launch_count"] + sorted(_SIGNATURES) + sorted(_PLAIN)


def register_optional(name, argtypes):
    _OPTIONAL[name] = argtypes
    if _lib is not None and hasattr(_lib, name):
        fn = getattr(_lib, name)
        fn.argtypes, fn.restype = argtypes, ctypes.c_int


def lib():
    """Load (once) and return the CDLL; raises if the library is missing or has the wrong ABI."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise PocketNerfError(
                "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                "indoor-nerf_b200/csrc/build.sh — there is no fallback path." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        l.pn_abi_version.restype = ctypes.c_int
        l.pn_last_error.restype = ctypes.c_char_p
        l.pn_launch_count.restype = ctypes.c_int64
        if l.pn_abi_version() != ABI_VERSION:
            raise PocketNerfError("libpocketnerf.so ABI %d != binding ABI %d: rebuild" % (l.pn_abi_version(), ABI_VERSION))
        for name, argtypes in list(_SIGNATURES.items()) + list(_OPTIONAL.items()):
            if name in _OPTIONAL and not hasattr(l, name):
                continue
            fn = getattr(l, name)
            fn.argtypes, fn.restype = argtypes, ctypes.c_int
        for name, (argtypes, restype) in _PLAIN.items():
            fn = getattr(l, name)
            fn.argtypes, fn.restype = argtypes, restype
        _lib = l
    return _lib


def launch_count():
    return int(lib().pn_launch_count())


def call(name, *args):
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise PocketNerfError("%s failed (%d): %s" % (name, rc, lib().pn_last_error().decode()))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream():
    """The current CUDA stream of the current device as a cudaStream_t.  torch.cuda.current_stream() builds a Stream
    object per call (~18 us, 26 calls per training step); the raw getter is a plain C call."""
    if _raw_stream is not None:
        return ctypes.c_void_p(_raw_stream(torch.cuda.current_device()))
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def dptr(t, dtype=torch.float32, allow_none=False):
    """Device pointer of a contiguous CUDA tensor (None -> NULL when allowed)."""
    if t is None:
        if allow_none:
            return None
        raise PocketNerfError("tensor argument is None")
    if not t.is_cuda:
        raise PocketNerfError("expected a CUDA tensor, got device %s — there is no CPU path" % t.device)
    if dtype is not None and t.dtype != dtype:
        raise PocketNerfError("expected dtype %s, got %s" % (dtype, t.dtype))
    if not t.is_contiguous():
        raise PocketNerfError("expected a contiguous tensor")
    return ctypes.c_void_p(t.data_ptr())


def fcontig(t):
    """float32 + contiguous view/copy of a CUDA tensor."""
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()
