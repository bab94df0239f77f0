"""HashEmbedder / SHEncoder with the reference's constructor, attributes, state_dict keys and return
values (hash_encoding.py:11-191) on top of the sm_100a kernels.

What is kept from the reference on purpose
  * 16 separate ``embeddings[l]`` modules, each callable on an int64 index tensor (the TV loss does
    that, loss.py:30) with a ``.weight [2^T, 2]`` Parameter -> same state_dict keys, same
    ``parameters()`` order for the optimiser (run_nerf.py:228,281-285);
  * ``base_resolution`` / ``finest_resolution`` as 0-d int64 tensors, ``b`` computed with the reference's
    float32 torch expression, ``warmup_steps``, ``current_step`` counting training forward calls;
  * ``quantizers`` (ModuleList of LearnedBitwidthQuantizer or None).
What is different
  * all 16 tables live in ONE contiguous [L, T, 2] buffer (``self.table_storage``); every
    ``embeddings[l].weight`` is a view of it, so a data-parallel step can all-reduce the whole table
    gradient in one NCCL call and an optimiser could walk it in one pass;
  * ``forward`` is one kernel launch (gather + fake-quant + interpolation for all levels) instead of
    ~60 ATen ops and 1-2 host syncs per level.
"""
import torch
import torch.nn as nn

from . import ops
from .quantization import LearnedBitwidthQuantizer, qrows_batched


class TableLevel(nn.Embedding):
    """One level of the hash table: an nn.Embedding whose weight is a view into the shared storage."""

    def __init__(self, storage_row):
        super().__init__(storage_row.shape[0], storage_row.shape[1], _weight=storage_row)


class HashEmbedder(nn.Module):
    def __init__(self, bounding_box, n_levels=16, n_features_per_level=2, log2_hashmap_size=19,
                 base_resolution=16, finest_resolution=512, use_quantization=False, quantization_bits=8):
        super().__init__()
        if n_features_per_level != 2:
            raise ValueError("the kernels are built for 2 features per level (the reference default)")
        if not 1 <= n_levels <= 16:
            raise ValueError("n_levels must be in 1..16")
        self.bounding_box = bounding_box
        self.n_levels = n_levels
        self.n_features_per_level = n_features_per_level
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = torch.tensor(base_resolution)
        self.finest_resolution = torch.tensor(finest_resolution)
        self.out_dim = self.n_levels * self.n_features_per_level
        self.use_quantization = use_quantization
        self.warmup_steps = 500
        self.current_step = 0
        # hash_encoding.py:28 — float32 on purpose (the floor() below sits on knife edges)
        self.b = torch.exp((torch.log(self.finest_resolution) - torch.log(self.base_resolution)) / (n_levels - 1))

        storage = torch.empty(n_levels, 2 ** log2_hashmap_size, n_features_per_level)
        nn.init.uniform_(storage, a=-0.0001, b=0.0001)          # hash_encoding.py:33-34
        self.table_storage = storage                             # plain attribute: not a second set of parameters
        self.embeddings = nn.ModuleList([TableLevel(storage[l]) for l in range(n_levels)])
        if use_quantization:
            self.quantizers = nn.ModuleList([
                LearnedBitwidthQuantizer(init_bits=float(quantization_bits), min_bits=2.0, max_bits=32.0,
                                         symmetric=False) for _ in range(n_levels)])
        else:
            self.quantizers = None
        self._grid = None
        self._grid_key = None
        self._packed = None                                       # ops.PackedLevels: tables as integer codes (inference)

    # -- storage management ---------------------------------------------------------------------------
    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._reflatten()
        self._packed = None                                       # a code snapshot does not follow .to() / .float()
        return out

    def _is_flat(self):
        s = self.table_storage
        return all(e.weight.data_ptr() == s[l].data_ptr() and e.weight.device == s.device
                   for l, e in enumerate(self.embeddings))

    def _reflatten(self):
        """Re-establish 'every weight is a view of one buffer' after .to()/.cuda()/load_state_dict(assign=True)."""
        if self._is_flat():
            return
        w0 = self.embeddings[0].weight
        storage = torch.empty((self.n_levels,) + tuple(w0.shape), dtype=w0.dtype, device=w0.device)
        with torch.no_grad():
            for l, e in enumerate(self.embeddings):
                storage[l].copy_(e.weight.data)
                e.weight.data = storage[l]
        self.table_storage = storage
        self._grid = None

    def tables(self):
        return [e.weight for e in self.embeddings]

    # -- geometry ---------------------------------------------------------------------------------------
    def level_resolutions(self):
        """hash_encoding.py:89 evaluated with the reference's expression on the device the tables live on."""
        dev = self.embeddings[0].weight.device
        base, b = self.base_resolution.to(dev), self.b.to(dev)
        return [torch.floor(base * b ** i) for i in range(self.n_levels)]

    def grid(self):
        box_min, box_max = self.bounding_box
        dev = self.embeddings[0].weight.device
        key = (str(dev), self.log2_hashmap_size, id(box_min), id(box_max))
        if self._grid is None or self._grid_key != key:
            res = torch.stack(self.level_resolutions()).float().cpu().tolist()          # one sync, cached
            bmin = torch.as_tensor(box_min).float().cpu().tolist()
            bmax = torch.as_tensor(box_max).float().cpu().tolist()
            self._grid = ops.make_grid(bmin, bmax, res, self.log2_hashmap_size)
            self._grid_key = key
        return self._grid

    # -- quantisation -------------------------------------------------------------------------------------
    def _quant_rows(self, x):
        """Device [L, 8] rows for the kernel, or None (hash_encoding.py:97-101)."""
        if not (self.use_quantization and self.quantizers is not None):
            return None
        if self.training and self.current_step < self.warmup_steps:
            return None
        if self.training and not all(q.calibrated for q in self.quantizers):
            mm = ops.hash_gather_minmax(self.grid(), [t.detach() for t in self.tables()], x)
            for l, q in enumerate(self.quantizers):
                if not q.calibrated:
                    q.calibrate_minmax(mm[l, 0], mm[l, 1])
        return qrows_batched(list(self.quantizers), self.training).contiguous()

    # -- tables as integer codes (inference) ----------------------------------------------------------------------
    def pack_for_inference(self):
        """Hold every level as the integer codes of its quantiser's eval form (u8 pairs up to 8 learned bits, u16
        pairs up to 16, dequantised fp32 above) and gather from those in eval mode: 2-4 bytes per corner instead of
        8, same output bit for bit as the fake-quantised fp32 tables (hash_encoding.py:97-101 in eval mode).  The
        copy is a snapshot: it is dropped when the module goes back to training."""
        if not (self.use_quantization and self.quantizers is not None and all(q.calibrated for q in self.quantizers)):
            raise RuntimeError("pack_for_inference needs calibrated table quantisers (use_quantization=True, trained)")
        if not self._is_flat():
            self._reflatten()
        levels = []
        with torch.no_grad():
            for l, q in enumerate(self.quantizers):
                row = q.qrow(training=False).contiguous()
                bits = q.integer_bit_width
                w = self.embeddings[l].weight.detach()
                r = row.cpu().tolist()
                if bits <= 16:
                    t = ops.quant_codes(w, row, 1 if bits <= 8 else 2)
                else:
                    was = q.training
                    q.eval()
                    t = q(w).float().contiguous()
                    q.train(was)
                levels.append((t, r[0], r[2], r[3]))
        self._packed = ops.PackedLevels(levels)
        return self._packed

    def set_packed(self, packed):
        self._packed = packed

    def train(self, mode=True):
        if mode:
            self._packed = None
        return super().train(mode)

    def forward(self, x):
        if self.training:
            self.current_step += 1
        if not self._is_flat():
            self._reflatten()
        x = x.reshape(-1, 3)
        if self._packed is not None and not self.training:
            return ops.hash_encode_fwd_packed(self.grid(), self._packed, x.detach())
        qrows = self._quant_rows(x.detach())
        feat, keep = ops.HashEncodeFn.apply(x, self.grid(), qrows, *self.tables())
        return feat, keep


class SHEncoder(nn.Module):
    """hash_encoding.py:110-191; only degree 4 (what get_embedder builds) has a kernel."""

    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        assert input_dim == 3
        if degree != 4:
            raise ValueError("SHEncoder kernel is built for degree 4 (the value get_embedder uses)")
        self.input_dim, self.degree = input_dim, degree
        self.out_dim = degree ** 2

    def forward(self, input, **kwargs):
        return ops.sh_encode(input)
