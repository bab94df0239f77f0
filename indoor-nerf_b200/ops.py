"""Operator layer: thin, typed wrappers over the C ABI plus the autograd formulas that tie the
forward and backward kernels together.  Nothing here computes on the CPU or falls back to torch ops for
the hot path; torch allocates the outputs and provides the stream."""
import ctypes
import os

import torch

from . import _lib
from ._lib import HashGrid, MlpInput, MlpWeights, PackedTables, call, dptr, fcontig, stream

_MLP_KEYS = ("s0", "s1", "c0", "c1", "c2", "n0w", "n0b", "n2w", "n2b")

# Arithmetic mode of the NeRFSmall kernels: "fp32" (FFMA, 1e-5 parity) or "bf16" (tcgen05 tensor cores, 2e-3).
_MLP_MODE = os.environ.get("POCKETNERF_MLP", "fp32")


# diagnostic: run the bf16 mode through the UNFUSED kernels (exact fp32 hash kernels + tcgen05 MLP on fp32 feature rows)
_BF16_UNFUSED = bool(os.environ.get("PN_BF16_UNFUSED"))
# diagnostic (unfused path only): PN_EXP=fwd16_bwd32 / fwd32_bwd16 mixes the arithmetic of the two directions
_EXP = os.environ.get("PN_EXP", "")
_EXP_FWD = "bf16" if "fwd16" in _EXP else ("fp32" if "fwd32" in _EXP else None)
_EXP_BWD = "bf16" if "bwd16" in _EXP else ("fp32" if "bwd32" in _EXP else None)


def set_mlp_mode(mode):
    global _MLP_MODE
    if mode not in ("fp32", "bf16"):
        raise ValueError("mlp mode must be 'fp32' or 'bf16'")
    _MLP_MODE = mode


def get_mlp_mode():
    return _MLP_MODE


import contextlib

_SAME_DEVICE = contextlib.nullcontext()


def _guard(t):
    """Context that makes t's device current for the launch.  The common case (already current) costs one comparison
    instead of a torch.cuda.device() enter / exit pair."""
    if not t.is_cuda:
        raise _lib.PocketNerfError("expected a CUDA tensor, got device %s — there is no CPU path" % t.device)
    if t.device.index == torch.cuda.current_device():
        return _SAME_DEVICE
    return torch.cuda.device(t.device)


def _ptr_array(tensors):
    return (ctypes.c_void_p * _lib.MAX_LEVELS)(*[t.data_ptr() for t in tensors])


# ---------------------------------------------------------------------------------------------------
# hash grid
# ---------------------------------------------------------------------------------------------------
def make_grid(box_min, box_max, resolutions, log2_hashmap_size):
    g = HashGrid()
    g.box_min[:] = [float(v) for v in box_min]
    g.box_max[:] = [float(v) for v in box_max]
    res = [float(r) for r in resolutions]
    if not 1 <= len(res) <= _lib.MAX_LEVELS:
        raise _lib.PocketNerfError("n_levels %d outside 1..%d" % (len(res), _lib.MAX_LEVELS))
    g.resolution[:len(res)] = res
    g.n_levels = len(res)
    g.log2_hashmap_size = int(log2_hashmap_size)
    return g


def _check_tables(grid, tables):
    rows = 1 << grid.log2_hashmap_size
    if len(tables) != grid.n_levels:
        raise _lib.PocketNerfError("expected %d tables, got %d" % (grid.n_levels, len(tables)))
    for t in tables:
        if tuple(t.shape) != (rows, 2):
            raise _lib.PocketNerfError("table shape %s != (%d, 2)" % (tuple(t.shape), rows))
        dptr(t)


_QUANT_IN_GATHER = bool(os.environ.get("PN_QUANT_IN_GATHER"))      # A/B switch: fake-quant every gathered corner instead


def quantized_tables(tables, qparams):
    """The tables with each level's fake-quant row applied to every entry (pn_table_fake_quant): what the gather would
    compute corner by corner, once per entry.  Returns a list of views of one fresh [L, T, 2] buffer."""
    L, T = len(tables), tables[0].shape[0]
    out = torch.empty((L, T, 2), dtype=torch.float32, device=tables[0].device)
    outs = list(out.unbind(0))
    with _guard(out):
        call("pn_table_fake_quant", _ptr_array(tables), _ptr_array(outs), L, T, dptr(qparams), stream())
    return outs


def hash_encode_fwd(grid, tables, x, qparams=None):
    """x[P,3] -> feat[P,2L], keep[P] (bool).   Reference: hash_encoding.py:82-107."""
    x = fcontig(x)
    _check_tables(grid, tables)
    if qparams is not None and not _QUANT_IN_GATHER and x.shape[0] > 0:
        tables, qparams = quantized_tables(tables, qparams), None
    P = x.shape[0]
    feat = torch.empty((P, 2 * grid.n_levels), dtype=torch.float32, device=x.device)
    keep = torch.empty((P,), dtype=torch.bool, device=x.device)
    if P == 0:
        _guard(x)
        return feat, keep
    with _guard(x):
        call("pn_hash_encode_fwd", ctypes.byref(grid), _ptr_array(tables), dptr(qparams, allow_none=True),
             dptr(x), P, dptr(feat), dptr(keep, torch.bool), stream())
    return feat, keep


def hash_encode_bwd(grid, dtables, x, dfeat):
    """Accumulate the dense table gradients into dtables (list of [T,2], caller-zeroed)."""
    x, dfeat = fcontig(x), fcontig(dfeat)
    _check_tables(grid, dtables)
    if x.shape[0] == 0:
        return
    with _guard(x):
        call("pn_hash_encode_bwd", ctypes.byref(grid), _ptr_array(dtables), dptr(x), dptr(dfeat), x.shape[0],
             stream())


def hash_indices(grid, x):
    x = fcontig(x)
    idx = torch.empty((x.shape[0], grid.n_levels, 8), dtype=torch.int32, device=x.device)
    if x.shape[0] == 0:
        return idx
    with _guard(x):
        call("pn_hash_indices", ctypes.byref(grid), dptr(x), x.shape[0], dptr(idx, torch.int32), stream())
    return idx


def hash_coords(coords, log2_hashmap_size):
    """utils.hash (utils.py:13-24) for integer coords[..., D<=3]."""
    c = coords.to(torch.int64).contiguous()
    out = torch.empty(c.shape[:-1], dtype=torch.int64, device=c.device)
    n = out.numel()
    if n == 0:
        return out
    with _guard(c):
        call("pn_hash_coords", dptr(c, torch.int64), n, c.shape[-1], int(log2_hashmap_size), dptr(out, torch.int64),
             stream())
    return out


def hash_gather_minmax(grid, tables, x):
    x = fcontig(x)
    _check_tables(grid, tables)
    mm = torch.empty((grid.n_levels, 2), dtype=torch.float32, device=x.device)
    mm[:, 0] = float("inf")
    mm[:, 1] = float("-inf")
    with _guard(x):
        call("pn_hash_gather_minmax", ctypes.byref(grid), _ptr_array(tables), dptr(x), x.shape[0], dptr(mm), stream())
    return mm


# ---------------------------------------------------------------------------------------------------
# table gradients: one persistent flat [L,T,2] buffer installed as the tables' .grad
# ---------------------------------------------------------------------------------------------------
# Each table has up to three gradient producers per step (coarse field, fine field, TV loss).  If every
# producer returned its own dense [T,2] gradients, the autograd engine would sum them out of place (views of
# a shared buffer cannot be accumulated in place in its input buffers): three 64 MiB memsets, 32 add launches
# and — worse — 16 unrelated result tensors, so the optimiser and the data-parallel all-reduce could no longer
# treat the table gradient as ONE buffer.  Instead the backward kernels accumulate (atomics) straight into a
# flat buffer whose per-level views ARE the parameters' .grad (the "main grad" pattern), and the autograd nodes
# return None for the tables.  Gradients that other producers already delivered to .grad are folded in first;
# optimizer.zero_grad() (set_to_none) is respected because the buffer is re-zeroed whenever .grad is found None.
# Consequence: use loss.backward() (as the reference does); torch.autograd.grad() w.r.t. the tables sees None.
_FLAT_GRADS = {}


def table_grad_buffer(tables):
    """The flat gradient buffer of `tables` with every level's view installed as .grad (zeroed / seeded as needed)."""
    arena = arena_of(list(tables))
    if arena is not None and len(tables) == len(arena.tables) and all(a is b for a, b in zip(tables, arena.tables)):
        return arena.ensure().table_flat
    t0 = tables[0]
    L = len(tables)
    key = (t0.data_ptr(), L, tuple(t0.shape), str(t0.device))
    flat = _FLAT_GRADS.get(key)
    if flat is None:
        if len(_FLAT_GRADS) > 8:
            _FLAT_GRADS.clear()
        flat = torch.empty((L,) + tuple(t0.shape), dtype=torch.float32, device=t0.device)
        _FLAT_GRADS[key] = flat
        installed = [False] * L
    else:
        installed = [t.grad is not None and t.grad.data_ptr() == flat[l].data_ptr() for l, t in enumerate(tables)]
    if not all(installed):
        if all(t.grad is None for t in tables):
            flat.zero_()
        else:
            for l, t in enumerate(tables):
                if installed[l]:
                    continue
                if t.grad is None:
                    flat[l].zero_()
                else:
                    flat[l].copy_(t.grad)
        for l, t in enumerate(tables):
            t.grad = flat[l]
    return flat


class GradArena:
    """ONE flat fp32 gradient buffer for a whole model: [ table gradients L*T*2 | MLP gradients of every network ].

    Every registered parameter's ``.grad`` is a view of it; the backward kernels accumulate into those views with
    atomics and the autograd nodes return None for them (the "main grad" pattern of the table buffer above, extended
    to the NeRFSmall weights).  A data-parallel step therefore all-reduces exactly one buffer with no packing or
    copy-back, and ``optimizer.zero_grad()`` (set_to_none) is honoured: a parameter found without ``.grad`` gets its
    segment zeroed before the next accumulation — with one memset when the whole model was released.
    ``create_nerf`` builds the arena; parameters that are not registered (or whose storage moved since, e.g. after
    ``.to()``) keep the ordinary autograd path.  Parameters are identified by their data pointer."""

    def __init__(self, tables, mlp_params):
        self.tables = list(tables)
        self.mlp = list(mlp_params)
        t0 = self.tables[0]
        if not all(t.shape == t0.shape and t.dtype == torch.float32 and t.device == t0.device for t in self.tables):
            raise _lib.PocketNerfError("GradArena: tables must share shape, dtype and device")
        self.device = t0.device
        off, self.views = 0, {}
        segs = []
        for p in self.tables + self.mlp:
            segs.append((p, off))
            off += (p.numel() + 3) // 4 * 4                          # 16-byte aligned segments
        self.numel = off
        self.flat = torch.zeros(off, dtype=torch.float32, device=self.device)
        for p, o in segs:
            self.views[p.data_ptr()] = self.flat[o:o + p.numel()].view(p.shape)
            p._pn_arena = self
        self.table_flat = self.flat[:len(self.tables) * t0.numel()].view((len(self.tables),) + tuple(t0.shape))

    def params(self):
        return self.tables + self.mlp

    def covers(self, p):
        v = self.views.get(p.data_ptr())
        return v is not None and v.shape == p.shape and p.device == self.device and p.dtype == torch.float32

    def view(self, p):
        return self.views[p.data_ptr()]

    def installed(self, p):
        g = p.grad
        return g is not None and g.data_ptr() == self.views[p.data_ptr()].data_ptr()

    def valid(self):
        return all(self.covers(p) for p in self.params())

    def ensure(self):
        """Install every view as its parameter's .grad; zero the segments of released parameters, fold in gradients that
        another producer delivered to a fresh tensor.  Cheap when everything is already installed (one Python loop)."""
        ps = self.params()
        todo = [p for p in ps if not self.installed(p)]
        if not todo:
            return self
        if len(todo) == len(ps) and all(p.grad is None for p in ps):
            self.flat.zero_()
        else:
            for p in todo:
                if p.grad is None:
                    self.view(p).zero_()
                else:
                    self.view(p).copy_(p.grad)
        for p in todo:
            p.grad = self.view(p)
        return self


def arena_of(params):
    """The GradArena that all of `params` are registered with and that still matches their storage, or None."""
    a = getattr(params[0], "_pn_arena", None)
    if a is None:
        return None
    for p in params:
        if getattr(p, "_pn_arena", None) is not a or not a.covers(p):
            return None
    return a


class HashEncodeFn(torch.autograd.Function):
    """feat, keep = HashEncodeFn.apply(x, grid, qparams, *tables).  Gradients flow to the tables only:
    sample positions never require grad on this path (z_samples is detached, run_nerf.py:510)."""

    @staticmethod
    def forward(ctx, x, grid, qparams, *tables):
        feat, keep = hash_encode_fwd(grid, [t.detach() for t in tables], x, qparams)
        ctx.grid = grid
        ctx.save_for_backward(x, *tables)
        ctx.mark_non_differentiable(keep)
        return feat, keep

    @staticmethod
    def backward(ctx, dfeat, _dkeep):
        x, *tables = ctx.saved_tensors
        if dfeat is not None and any(ctx.needs_input_grad[3:]):
            flat = table_grad_buffer(tables)
            hash_encode_bwd(ctx.grid, list(flat.unbind(0)), x, dfeat)
        return (None, None, None) + (None,) * len(tables)


class TVLossFn(torch.autograd.Function):
    """losses[L] = TVLossFn.apply(min_vertex[L,3] int64, cubes (tuple of ints), log2T, *tables):
    total_variation_loss (loss.py:11-43) of every level in one launch, and its gradient in one more."""

    @staticmethod
    def forward(ctx, min_vertex, cubes, log2T, *tables):
        mv = min_vertex.to(torch.int64).contiguous()
        dev = tables[0].device
        L = len(tables)
        loss = torch.zeros(L, dtype=torch.float32, device=dev)
        cube = (ctypes.c_int32 * L)(*[int(c) for c in cubes])
        with _guard(tables[0]):
            call("pn_tv_loss_fwd", _ptr_array([t.detach() for t in tables]), L, int(log2T), cube,
                 dptr(mv, torch.int64), dptr(loss), stream())
        ctx.cubes, ctx.log2T = tuple(int(c) for c in cubes), int(log2T)
        ctx.save_for_backward(mv, *tables)
        return loss

    @staticmethod
    def backward(ctx, dloss):
        mv, *tables = ctx.saved_tensors
        L = len(tables)
        if len({t.data_ptr() for t in tables}) == L:
            flat = table_grad_buffer(tables)
            targets = list(flat.unbind(0))
            ret = (None,) * L
        else:                                   # the same tensor passed for several levels: plain dense gradients
            targets = [torch.zeros_like(t) for t in tables]
            ret = tuple(targets)
        cube = (ctypes.c_int32 * L)(*ctx.cubes)
        with _guard(tables[0]):
            call("pn_tv_loss_bwd", _ptr_array([t.detach() for t in tables]), _ptr_array(targets), L,
                 ctx.log2T, cube, dptr(mv, torch.int64), dptr(fcontig(dloss)), stream())
        return (None, None, None) + ret


def radam_step(p, g, m, v, beta1, beta2, eps, wd_lr, step_lr, mode):
    """One fused RAdam update over contiguous fp32 CUDA buffers of equal numel (radam.py:55-88)."""
    n = p.numel()
    with _guard(p):
        call("pn_radam_step", dptr(p), dptr(g), dptr(m), dptr(v), n, float(beta1), float(beta2), float(eps),
             float(wd_lr), float(step_lr), int(mode), stream())


def radam_step_dyn(p, g, m, v, beta1, beta2, eps, dyn, mode):
    """radam_step with {wd*lr, step_size*lr} read from the device pair `dyn` (a launch that can live in a CUDA graph)."""
    n = p.numel()
    with _guard(p):
        call("pn_radam_step_dyn", dptr(p), dptr(g), dptr(m), dptr(v), n, float(beta1), float(beta2), float(eps),
             dptr(dyn), int(mode), stream())


def store_floats(dst, values):
    """dst[:len(values)] = values (host floats, at most 16) with one stream-ordered launch and no staging buffer."""
    import ctypes
    vals = (ctypes.c_float * len(values))(*values)
    with _guard(dst):
        call("pn_store_floats", dptr(dst), vals, len(values), stream())


# ---------------------------------------------------------------------------------------------------
# view directions
# ---------------------------------------------------------------------------------------------------
def sh_encode(dirs):
    """hash_encoding.py:153-191, degree 4."""
    d = fcontig(dirs.reshape(-1, 3))
    out = torch.empty((d.shape[0], 16), dtype=torch.float32, device=d.device)
    if d.shape[0] == 0:
        return out.reshape(*dirs.shape[:-1], 16)
    with _guard(d):
        call("pn_sh_encode", dptr(d), d.shape[0], dptr(out), stream())
    return out.reshape(*dirs.shape[:-1], 16)


# ---------------------------------------------------------------------------------------------------
# NeRFSmall
# ---------------------------------------------------------------------------------------------------
def _weights_struct(w):
    s = MlpWeights()
    for k in _MLP_KEYS:
        t = w.get(k)
        setattr(s, k, None if t is None else dptr(t))
    return s


def _rowptr(t, width):
    """Pointer + row stride of a 2-D fp32 CUDA tensor whose rows are contiguous (a column slice of a wider
    tensor is fine)."""
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[1] == width
            and (t.stride(1) == 1 or t.shape[1] == 1)):
        raise _lib.PocketNerfError("expected a CUDA fp32 [n,%d] tensor with contiguous rows" % width)
    stride = t.stride(0) if t.shape[0] > 1 else max(t.stride(0), width)
    return ctypes.c_void_p(t.data_ptr()), stride


def _mlp_input(feat, sh, dirs, samples_per_ray, act_q, keep):
    i = MlpInput()
    i.feat, i.feat_stride = _rowptr(feat, 32)
    i.sh, i.sh_stride = _rowptr(sh, 16) if sh is not None else (None, 0)
    i.dirs = dptr(dirs) if dirs is not None else None
    i.samples_per_ray = int(samples_per_ray)
    i.act_q = dptr(act_q, allow_none=True)
    i.keep = dptr(keep, torch.bool, allow_none=True)
    i.n_points = feat.shape[0]
    return i


def mlp_fwd(w, feat, sh=None, dirs=None, samples_per_ray=1, act_q=None, keep=None, mode=None):
    """w: dict of contiguous fp32 CUDA weights (keys s0,s1,c0,c1,c2[,n0w,n0b,n2w,n2b]).  Returns
    out[P, 4|7].   Reference: run_nerf_helpers.py:265-306 (+ run_nerf.py:59-66 when dirs/keep given)."""
    C = 7 if w.get("n0w") is not None else 4
    out = torch.empty((feat.shape[0], C), dtype=torch.float32, device=feat.device)
    if feat.shape[0] == 0:
        _guard(feat)
        return out
    with _guard(feat):
        ws, inp = _weights_struct(w), _mlp_input(feat, sh, dirs, samples_per_ray, act_q, keep)
        call("pn_mlp_fwd_bf16" if (mode or _MLP_MODE) == "bf16" else "pn_mlp_fwd", ctypes.byref(ws), ctypes.byref(inp),
             dptr(out), stream())
    return out


def mlp_bwd(w, feat, dout, sh=None, dirs=None, samples_per_ray=1, act_q=None, keep=None, want_dsh=False, mode=None):
    """Returns dfeat[P,32], dsh[P,16] or None, dict of weight gradients."""
    dout = fcontig(dout)
    P = feat.shape[0]
    dfeat = torch.empty((P, 32), dtype=torch.float32, device=feat.device)
    dsh = torch.empty((P, 16), dtype=torch.float32, device=feat.device) if (want_dsh and sh is not None) else None
    dw = {k: torch.zeros_like(w[k]) for k in _MLP_KEYS if w.get(k) is not None}
    if P == 0:
        return dfeat, dsh, dw
    with _guard(feat):
        ws, inp = _weights_struct(w), _mlp_input(feat, sh, dirs, samples_per_ray, act_q, keep)
        gs = _weights_struct(dw)
        call("pn_mlp_bwd_bf16" if (mode or _MLP_MODE) == "bf16" else "pn_mlp_bwd", ctypes.byref(ws), ctypes.byref(inp),
             dptr(dout), dptr(dfeat), 32,
             dptr(dsh, allow_none=True), 16, ctypes.byref(gs), stream())
    return dfeat, dsh, dw


class MlpFn(torch.autograd.Function):
    """out = MlpFn.apply(x48, act_q, n_weights, *weights) — NeRFSmall.forward on [B,48] inputs."""

    @staticmethod
    def forward(ctx, x, act_q, keys, *weights):
        x = fcontig(x)
        w = {k: fcontig(t.detach()) for k, t in zip(keys, weights)}
        feat, sh = x[:, :32], x[:, 32:48]
        ctx.mode = _MLP_MODE
        out = mlp_fwd(w, feat, sh=sh, act_q=act_q, mode=ctx.mode)
        ctx.keys, ctx.act_q = keys, act_q
        ctx.save_for_backward(x, *weights)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, *weights = ctx.saved_tensors
        w = {k: fcontig(t.detach()) for k, t in zip(ctx.keys, weights)}
        dfeat, dsh, dw = mlp_bwd(w, x[:, :32], dout, sh=x[:, 32:48], act_q=ctx.act_q, want_dsh=True, mode=ctx.mode)
        dx = torch.cat([dfeat, dsh], -1) if ctx.needs_input_grad[0] else None
        return (dx, None, None) + tuple(dw[k] for k in ctx.keys)


_BWD_WORKSPACE = {}
_FROZEN_SINK = {}
# bench.py's live kernel timing: when a list is installed, the two fused field launches are bracketed by CUDA events
# on the launching stream and (name, n_points, start, end) is appended; None (default) costs one `is None` test.
KERNEL_EVENTS = None


def _timed_call(name, n_points, *args):
    if KERNEL_EVENTS is None:
        return call(name, *args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call(name, *args)
    e1.record()
    KERNEL_EVENTS.append((name, int(n_points), e0, e1))


def _frozen_table_sink(tables):
    """Where the fused backward scatters when the tables do not require grad: one cached buffer per (shape, device),
    never read (was: a fresh 64-512 MiB zero tensor per backward)."""
    key = (len(tables), tuple(tables[0].shape), str(tables[0].device))
    buf = _FROZEN_SINK.get(key)
    if buf is None:
        _FROZEN_SINK.clear()
        buf = torch.zeros((len(tables),) + tuple(tables[0].shape), dtype=torch.float32, device=tables[0].device)
        _FROZEN_SINK[key] = buf
    return buf



def _field_bwd_workspace(device):
    """Scratch of the fused backward (its dX ring, see include/pocketnerf.h): one buffer per device, reused by every
    launch — launches on one stream are ordered, and the kernel leaves nothing in it."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    buf = _BWD_WORKSPACE.get(key)
    if buf is None:
        with torch.cuda.device(device):
            n = int(_lib.lib().pn_field_bwd_workspace_bytes())
        buf = torch.empty(n, dtype=torch.uint8, device=device)
        _BWD_WORKSPACE[key] = buf
    return buf


class FieldFn(torch.autograd.Function):
    """raw[P, 4|7] = FieldFn.apply(pts[P,3], viewdirs[N,3], S, grid, qparams, act_q, keys, n_tables,
    *tables, *weights): hash encode -> SH -> NeRFSmall -> keep-mask, i.e. run_network
    (run_nerf.py:53-68) as one autograd node.  The feature tensor is kept for the backward; the table
    gradients are accumulated into the flat buffer installed as the tables' .grad (table_grad_buffer), so
    data-parallel training all-reduces them in one call and RAdam updates them in one launch."""

    @staticmethod
    def forward(ctx, pts, viewdirs, S, grid, qparams, act_q, keys, n_tables, *params):
        tables, weights = params[:n_tables], params[n_tables:]
        pts = fcontig(pts)
        dirs = fcontig(viewdirs)
        w = {k: fcontig(t.detach()) for k, t in zip(keys, weights)}
        ctx.mode = _MLP_MODE
        ctx.fused = ctx.mode == "bf16" and grid.n_levels == 16 and pts.shape[0] > 0 and not _BF16_UNFUSED
        if ctx.fused:
            # one kernel: hash encode -> SH -> NeRFSmall (tcgen05) -> keep mask; features saved as bf16 operand tiles
            P = pts.shape[0]
            C = 7 if w.get("n0w") is not None else 4
            out = torch.empty((P, C), dtype=torch.float32, device=pts.device)
            keep = torch.empty((P,), dtype=torch.bool, device=pts.device)
            need_bwd = any(ctx.needs_input_grad[8:])     # all False under no_grad (render / eval): no feature tiles
            feat = torch.empty(((P + 127) // 128) * 8192, dtype=torch.uint8, device=pts.device) if need_bwd else None
            _check_tables(grid, tables)
            ktabs, kq = [t.detach() for t in tables], qparams
            if kq is not None and not _QUANT_IN_GATHER:
                ktabs, kq = quantized_tables(ktabs, kq), None     # fake-quant per table entry, not per gathered corner
            with _guard(pts):
                ws = _weights_struct(w)
                _timed_call("pn_field_fwd_bf16", P, ctypes.byref(grid), _ptr_array(ktabs),
                     dptr(kq, allow_none=True), ctypes.byref(ws), dptr(pts), dptr(dirs), int(S),
                     dptr(act_q, allow_none=True), P, dptr(out), dptr(keep, torch.bool),
                     dptr(feat, torch.uint8, allow_none=True), stream())
            if feat is None:
                feat = torch.empty(0, dtype=torch.uint8, device=pts.device)
        else:
            feat, keep = hash_encode_fwd(grid, [t.detach() for t in tables], pts, qparams)
            out = mlp_fwd(w, feat, dirs=dirs, samples_per_ray=S, act_q=act_q, keep=keep, mode=_EXP_FWD or ctx.mode)
        ctx.grid, ctx.S, ctx.act_q, ctx.keys, ctx.n_tables = grid, S, act_q, keys, n_tables
        ctx.save_for_backward(pts, dirs, feat, keep, *params)
        return out

    @staticmethod
    def backward(ctx, dout):
        pts, dirs, feat, keep, *params = ctx.saved_tensors
        tables, weights = params[:ctx.n_tables], params[ctx.n_tables:]
        w = {k: fcontig(t.detach()) for k, t in zip(ctx.keys, weights)}
        if ctx.fused:
            # one kernel: NeRFSmall backward (tcgen05) + run-aggregated scatter into the flat table gradient
            dout = fcontig(dout)
            if any(ctx.needs_input_grad[8:8 + ctx.n_tables]):
                flat = table_grad_buffer(tables)
            else:                               # frozen tables: the kernel still needs somewhere to scatter
                flat = _frozen_table_sink(tables)
            # weight gradients: straight into the model's gradient arena where the weight is a registered leaf
            # parameter (the kernel accumulates with atomics); through autograd otherwise (e.g. the fake-quantised W0)
            arena = arena_of(list(tables))
            if arena is not None:
                arena.ensure()
            dw, direct = {}, set()
            for k, t in zip(ctx.keys, weights):
                if (arena is not None and t.is_leaf and t.requires_grad and t.is_contiguous()
                        and getattr(t, "_pn_arena", None) is arena and arena.covers(t)):
                    dw[k] = arena.view(t)
                    direct.add(k)
                else:
                    dw[k] = torch.zeros_like(w[k])
            with _guard(pts):
                ws, gs = _weights_struct(w), _weights_struct(dw)
                wsp = _field_bwd_workspace(pts.device)
                _timed_call("pn_field_bwd_bf16", pts.shape[0], ctypes.byref(ctx.grid), _ptr_array(list(flat.unbind(0))), ctypes.byref(ws),
                     dptr(feat, torch.uint8), dptr(pts), dptr(dirs), int(ctx.S), dptr(ctx.act_q, allow_none=True),
                     dptr(keep, torch.bool), dptr(dout), pts.shape[0], ctypes.byref(gs), dptr(wsp, torch.uint8),
                     wsp.numel(), stream())
            return (None,) * (8 + ctx.n_tables) + tuple(None if k in direct else dw[k] for k in ctx.keys)
        dfeat, _, dw = mlp_bwd(w, feat, dout, dirs=dirs, samples_per_ray=ctx.S, act_q=ctx.act_q, keep=keep,
                               mode=_EXP_BWD or ctx.mode)
        if any(ctx.needs_input_grad[8:8 + ctx.n_tables]):
            flat = table_grad_buffer(tables)
            hash_encode_bwd(ctx.grid, list(flat.unbind(0)), pts, dfeat)
        return (None,) * (8 + ctx.n_tables) + tuple(dw[k] for k in ctx.keys)


# ---------------------------------------------------------------------------------------------------
# compositing
# ---------------------------------------------------------------------------------------------------
class CompositeFn(torch.autograd.Function):
    """(rgb, disp, acc, weights, depth, sparsity, normal|None) = CompositeFn.apply(raw, z, rays_d,
    noise|None, white_bkgd) — raw2outputs (run_nerf.py:347-411)."""

    @staticmethod
    def forward(ctx, raw, z, rays_d, noise, white):
        raw, z, rays_d = fcontig(raw), fcontig(z), fcontig(rays_d)
        noise = fcontig(noise) if noise is not None else None
        N, S, C = raw.shape
        dev = raw.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        rgb, disp, acc, wts, depth, sp = f(N, 3), f(N), f(N), f(N, S), f(N), f(N)
        normal = f(N, 3) if C == 7 else None
        if N > 0:
          with _guard(raw):
            call("pn_composite_fwd", dptr(raw), C, dptr(z), dptr(rays_d), dptr(noise, allow_none=True), N, S,
                 int(bool(white)), dptr(rgb), dptr(disp), dptr(acc), dptr(wts), dptr(depth), dptr(sp),
                 dptr(normal, allow_none=True), stream())
        ctx.white = bool(white)
        ctx.save_for_backward(raw, z, rays_d, noise)
        ctx.set_materialize_grads(False)
        if normal is None:
            return rgb, disp, acc, wts, depth, sp
        return rgb, disp, acc, wts, depth, sp, normal

    @staticmethod
    def backward(ctx, d_rgb, d_disp, d_acc, d_wts, d_depth, d_sp, d_normal=None):
        raw, z, rays_d, noise = ctx.saved_tensors
        N, S, C = raw.shape
        draw = torch.empty_like(raw)
        c = lambda t: fcontig(t) if t is not None else None
        d_rgb, d_disp, d_acc, d_wts, d_depth, d_sp, d_normal = map(c, (d_rgb, d_disp, d_acc, d_wts, d_depth, d_sp, d_normal))
        if N > 0:
          with _guard(raw):
            call("pn_composite_bwd", dptr(raw), C, dptr(z), dptr(rays_d), dptr(noise, allow_none=True), N, S,
                 int(ctx.white), *[dptr(t, allow_none=True) for t in (d_rgb, d_disp, d_acc, d_wts, d_depth, d_sp, d_normal)],
                 dptr(draw), stream())
        return draw, None, None, None, None


# ---------------------------------------------------------------------------------------------------
# sampling, sorting, rays
# ---------------------------------------------------------------------------------------------------
def _u_arg(u, N, M):
    u = fcontig(u)
    if u.dim() == 1:
        return u, 0
    if tuple(u.shape) != (N, M):
        raise _lib.PocketNerfError("u shape %s != (%d, %d)" % (tuple(u.shape), N, M))
    return u, M


def sample_pdf(bins, weights, u, return_inds=False, return_cdf=False):
    """bins[N,nb], weights[N,nb-1] (any row stride), u[N,M] or [M] (broadcast).  Not differentiable
    (the reference detaches the result, run_nerf.py:510)."""
    bins = fcontig(bins.detach())
    weights = weights.detach()
    if weights.dtype != torch.float32 or weights.stride(-1) != 1:
        weights = fcontig(weights)
    N, nb = bins.shape
    M = u.shape[-1]
    u, us = _u_arg(u.detach(), N, M)
    samples = torch.empty((N, M), dtype=torch.float32, device=bins.device)
    inds = torch.empty((N, M), dtype=torch.int32, device=bins.device) if return_inds else None
    cdf = torch.empty((N, nb), dtype=torch.float32, device=bins.device) if return_cdf else None
    if N > 0:
      with _guard(bins):
        call("pn_sample_pdf", dptr(bins), ctypes.c_void_p(weights.data_ptr()), weights.stride(0), dptr(u), us, N, nb, M,
             dptr(samples), dptr(inds, torch.int32, allow_none=True), dptr(cdf, allow_none=True), stream())
    out = (samples,)
    if return_inds:
        out += (inds,)
    if return_cdf:
        out += (cdf,)
    return out if len(out) > 1 else samples


def sample_from_cdf(cdf, bins, u):
    cdf, bins = fcontig(cdf), fcontig(bins)
    N, nb = bins.shape
    M = u.shape[-1]
    u, us = _u_arg(u, N, M)
    samples = torch.empty((N, M), dtype=torch.float32, device=bins.device)
    inds = torch.empty((N, M), dtype=torch.int32, device=bins.device)
    if N > 0:
      with _guard(bins):
        call("pn_sample_from_cdf", dptr(cdf), dptr(bins), dptr(u), us, N, nb, M, dptr(samples),
             dptr(inds, torch.int32), stream())
    return samples, inds


def sort_merge(a, b):
    """torch.sort(torch.cat([a, b], -1), -1)[0] for 2-D inputs (run_nerf.py:512)."""
    a, b = fcontig(a.detach()), fcontig(b.detach())
    N = a.shape[0]
    out = torch.empty((N, a.shape[1] + b.shape[1]), dtype=torch.float32, device=a.device)
    if N > 0:
      with _guard(a):
        call("pn_sort_merge", dptr(a), a.shape[1], dptr(b), b.shape[1], N, dptr(out), stream())
    return out


def gen_rays(H, W, K, c2w, device):
    """get_rays (run_nerf_helpers.py:311-320).  K: 3x3 array-like (host), c2w: [3,4] (host or device)."""
    Kf = (ctypes.c_float * 9)(*[float(K[i][j]) for i in range(3) for j in range(3)])
    c = c2w.detach().cpu().tolist() if torch.is_tensor(c2w) else [list(r) for r in c2w]
    Cf = (ctypes.c_float * 12)(*[float(c[i][j]) for i in range(3) for j in range(4)])
    rays_o = torch.empty((H, W, 3), dtype=torch.float32, device=device)
    rays_d = torch.empty((H, W, 3), dtype=torch.float32, device=device)
    with _guard(rays_o):
        call("pn_gen_rays", int(H), int(W), Kf, Cf, dptr(rays_o), dptr(rays_d), stream())
    return rays_o, rays_d


def ndc_rays(H, W, focal, near, rays_o, rays_d):
    """run_nerf_helpers.py:333-350."""
    shape = rays_o.shape
    o, d = fcontig(rays_o.reshape(-1, 3)), fcontig(rays_d.reshape(-1, 3))
    oo, od = torch.empty_like(o), torch.empty_like(d)
    if o.shape[0] > 0:
      with _guard(o):
        call("pn_ndc_rays", int(H), int(W), float(focal), float(near), dptr(o), dptr(d), o.shape[0], dptr(oo), dptr(od),
             stream())
    return oo.reshape(shape), od.reshape(shape)


def make_points(rays_o, rays_d, z):
    """pts[N,S,3] = o + d*z (run_nerf.py:490,513); rays_o/rays_d may be strided row views."""
    z = fcontig(z)
    N, S = z.shape

    def rows(t):
        if t.dtype != torch.float32 or t.stride(-1) != 1:
            t = fcontig(t)
        return t

    o, d = rows(rays_o), rows(rays_d)
    pts = torch.empty((N, S, 3), dtype=torch.float32, device=z.device)
    if N > 0:
      with _guard(z):
        call("pn_make_points", ctypes.c_void_p(o.data_ptr()), o.stride(0), ctypes.c_void_p(d.data_ptr()), d.stride(0),
             dptr(z), N, S, dptr(pts), stream())
    return pts


def coarse_z(near, far, t_vals, t_rand=None, lindisp=False):
    """run_nerf.py:466-488.  near/far: [N] or [N,1] views (row stride allowed); t_vals[S]."""
    N = near.shape[0]
    S = t_vals.shape[0]
    if near.stride(0) != far.stride(0) or near.dtype != torch.float32 or far.dtype != torch.float32:
        near, far = fcontig(near.reshape(N)), fcontig(far.reshape(N))
    t_vals = fcontig(t_vals)
    t_rand = fcontig(t_rand) if t_rand is not None else None
    z = torch.empty((N, S), dtype=torch.float32, device=t_vals.device)
    if N > 0:
      with _guard(z):
        call("pn_coarse_z", ctypes.c_void_p(near.data_ptr()), ctypes.c_void_p(far.data_ptr()), near.stride(0),
             dptr(t_vals), dptr(t_rand, allow_none=True), N, S, int(bool(lindisp)), dptr(z), stream())
    return z


# ---------------------------------------------------------------------------------------------------
# data formats either side of the path
# ---------------------------------------------------------------------------------------------------
def ray_bank_batch(ids, H, W, K, poses, image_index=None, images=None, f64_dirs=True):
    """batch_rays [2,B,3] (and target_s [B,3] when images is given) for ray ids (slot*H + j)*W + i.
    poses: CUDA fp32 [N,3,4] or [N,4,4]; images: CUDA fp32 or uint8 [N,H,W,3].  f64_dirs=True reproduces the
    precomputed bank of run_nerf.py:899-905 (get_rays_np with a float64 K), False the get_rays of :981."""
    if ids.dtype != torch.int64:
        ids = ids.to(torch.int64)
    ids = ids.contiguous()
    B = ids.shape[0]
    poses = fcontig(poses)
    if poses.dim() != 3 or tuple(poses.shape[1:]) not in ((3, 4), (4, 4)):
        raise _lib.PocketNerfError("poses must be [N,3,4] or [N,4,4], got %s" % (tuple(poses.shape),))
    Kd = (ctypes.c_double * 9)(*[float(K[i][j]) for i in range(3) for j in range(3)])
    rays = torch.empty((2, B, 3), dtype=torch.float32, device=ids.device)
    target, img_ptr, img_dt = None, None, 0
    if images is not None:
        if images.dtype not in (torch.float32, torch.uint8) or tuple(images.shape[1:]) != (H, W, 3):
            raise _lib.PocketNerfError("images must be fp32 or uint8 [N,%d,%d,3], got %s %s" % (H, W, images.dtype, tuple(images.shape)))
        img_dt = 1 if images.dtype == torch.uint8 else 0
        img_ptr = dptr(images, images.dtype)
        target = torch.empty((B, 3), dtype=torch.float32, device=ids.device)
    if B > 0:
        with _guard(ids):
            call("pn_ray_bank_batch", dptr(ids, torch.int64), B, int(H), int(W), Kd, dptr(poses),
                 poses.shape[1] * poses.shape[2], dptr(image_index, torch.int32, allow_none=True), img_ptr, img_dt,
                 int(bool(f64_dirs)), dptr(rays), dptr(target, allow_none=True), stream())
    else:
        _guard(ids)
    return (rays, target) if images is not None else rays


def image_sqerr(a, b, out=None):
    """Device double scalar: sum((a-b)^2) — accumulated into `out` when given (caller-zeroed)."""
    a, b = fcontig(a), fcontig(b)
    if a.shape != b.shape:
        raise _lib.PocketNerfError("shape mismatch %s vs %s" % (tuple(a.shape), tuple(b.shape)))
    s = out if out is not None else torch.zeros((), dtype=torch.float64, device=a.device)
    if a.numel() > 0:
        with _guard(a):
            call("pn_image_sqerr", dptr(a), dptr(b), a.numel(), dptr(s, torch.float64), stream())
    else:
        _guard(a)
    return s


def image_ssim(a, b, data_range=1.0):
    """Mean SSIM of two [H,W,C] images as a device double scalar (scikit-image defaults, evaluation_utils.py:33)."""
    a, b = fcontig(a), fcontig(b)
    if a.shape != b.shape or a.dim() != 3:
        raise _lib.PocketNerfError("expected two [H,W,C] images, got %s and %s" % (tuple(a.shape), tuple(b.shape)))
    H, W, C = a.shape
    s = torch.zeros((), dtype=torch.float64, device=a.device)
    with _guard(a):
        call("pn_image_ssim", dptr(a), dptr(b), H, W, C, float(data_range), dptr(s, torch.float64), stream())
    return s / float((H - 6) * (W - 6) * C)


def to8b(x):
    """uint8 image (255*clip(x,0,1)).astype(uint8) on the device (run_nerf_helpers.py:13)."""
    x = fcontig(x)
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    if x.numel() > 0:
        with _guard(x):
            call("pn_to8b", dptr(x), x.numel(), dptr(out, torch.uint8), stream())
    else:
        _guard(x)
    return out


def quant_pack(x, qrow, bits):
    """int32 words [n*bits/32]: the eval-form integer codes of a LearnedBitwidthQuantizer, bit-packed."""
    x = fcontig(x).reshape(-1)
    n = x.numel()
    words = torch.empty((n * int(bits) // 32,), dtype=torch.int32, device=x.device)
    with _guard(x):
        call("pn_quant_pack", dptr(x), n, dptr(fcontig(qrow)), int(bits), dptr(words, torch.int32), stream())
    return words


def quant_unpack(words, n, qrow, bits):
    """fp32 [n]: (code + qmin - zp) * scale — what the quantiser's eval forward returns for the packed tensor."""
    x = torch.empty((n,), dtype=torch.float32, device=words.device)
    with _guard(words):
        call("pn_quant_unpack", dptr(words.contiguous(), torch.int32), n, dptr(fcontig(qrow)), int(bits), dptr(x), stream())
    return x


# ---------------------------------------------------------------------------------------------------
# hash tables resident as integer codes (inference with a trained A-CAQ model)
# ---------------------------------------------------------------------------------------------------
def quant_codes(x, qrow, code_bytes):
    """u8 / u16 container codes of fp32 values under a quantiser's eval-form row."""
    x = fcontig(x)
    out = torch.empty(x.shape, dtype=torch.uint8 if code_bytes == 1 else torch.int16, device=x.device)
    with _guard(x):
        call("pn_quant_codes", dptr(x), x.numel(), dptr(fcontig(qrow)), int(code_bytes), dptr(out, out.dtype), stream())
    return out


def quant_unpack_codes(words, n, bits, code_bytes):
    """u8 / u16 container codes from a pn_quant_pack bit stream."""
    out = torch.empty((n,), dtype=torch.uint8 if code_bytes == 1 else torch.int16, device=words.device)
    with _guard(words):
        call("pn_quant_unpack_codes", dptr(words.contiguous(), torch.int32), n, int(bits), int(code_bytes),
             dptr(out, out.dtype), stream())
    return out


class PackedLevels:
    """Per-level code tensors + decode scalars in the form the kernels take (struct pn_packed_tables).
    levels: list of (tensor, scale, zp, qmin); tensor is uint8 [T,2], int16 [T,2] (u16 codes) or float32 [T,2]."""

    def __init__(self, levels):
        if not 1 <= len(levels) <= _lib.MAX_LEVELS:
            raise _lib.PocketNerfError("1..16 levels expected")
        self.tensors = []
        self.struct = PackedTables()
        rows = levels[0][0].shape[0]
        for l, (t, scale, zp, qmin) in enumerate(levels):
            if not t.is_cuda or not t.is_contiguous() or tuple(t.shape) != (rows, 2):
                raise _lib.PocketNerfError("level %d: expected a contiguous CUDA [%d,2] tensor" % (l, rows))
            eb = {torch.uint8: 2, torch.int16: 4, torch.float32: 8}.get(t.dtype)
            if eb is None:
                raise _lib.PocketNerfError("level %d: dtype %s (uint8, int16 or float32)" % (l, t.dtype))
            self.tensors.append(t)
            self.struct.codes[l] = t.data_ptr()
            self.struct.entry_bytes[l] = eb
            self.struct.scale[l], self.struct.zero_point[l], self.struct.qmin[l] = float(scale), float(zp), float(qmin)
        self.rows = rows

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors)


def hash_encode_fwd_packed(grid, packed, x):
    """hash_encode_fwd on code tables -> feat[P,2L], keep[P]; equals the quantised embedder's eval output bit for bit."""
    x = fcontig(x)
    if len(packed.tensors) != grid.n_levels or packed.rows != (1 << grid.log2_hashmap_size):
        raise _lib.PocketNerfError("packed tables do not match the grid")
    P = x.shape[0]
    feat = torch.empty((P, 2 * grid.n_levels), dtype=torch.float32, device=x.device)
    keep = torch.empty((P,), dtype=torch.bool, device=x.device)
    with _guard(x):
        if P > 0:
            call("pn_hash_encode_fwd_packed", ctypes.byref(grid), ctypes.byref(packed.struct), dptr(x), P, dptr(feat),
                 dptr(keep, torch.bool), stream())
    return feat, keep


def field_fwd_packed(grid, packed, w, pts, viewdirs, S, act_q=None):
    """run_network on code tables in the bf16 tensor-core mode (one kernel, inference only) -> raw[P, 4|7]."""
    pts, dirs = fcontig(pts), fcontig(viewdirs)
    if grid.n_levels != 16 or len(packed.tensors) != 16 or packed.rows != (1 << grid.log2_hashmap_size):
        raise _lib.PocketNerfError("the fused field kernel needs 16 packed levels matching the grid")
    P = pts.shape[0]
    C = 7 if w.get("n0w") is not None else 4
    out = torch.empty((P, C), dtype=torch.float32, device=pts.device)
    with _guard(pts):
        if P > 0:
            ws = _weights_struct(w)
            call("pn_field_fwd_bf16_packed", ctypes.byref(grid), ctypes.byref(packed.struct), ctypes.byref(ws), dptr(pts),
                 dptr(dirs), int(S), dptr(act_q, allow_none=True), P, dptr(out), None, stream())
    return out
