"""A-CAQ fake-quantisers with the reference's interface (quantization.py).

``LearnedBitwidthQuantizer`` keeps the reference's parameters / buffers / attributes
(``soft_bits, range_scale, v_max, running_min, running_max, min_bits, max_bits, symmetric,
calibrated, bit_width, integer_bit_width``) because the training loop reads and nudges them
(run_nerf.py:1165-1252).  Besides the module ``forward`` (used on small tensors such as the first
sigma-layer weight), it can export its scalars as one device row (``qrow``) that the hash-gather and
MLP kernels consume, so the quantisation of the gathered embeddings / hidden activations is fused
into those kernels and costs no extra pass.  No quantiser parameter receives a gradient (pure
straight-through estimator), exactly as in the reference.
"""
import torch
import torch.nn as nn

from ._lib import QROW


class LearnedBitwidthQuantizer(nn.Module):
    """quantization.py:68-193."""

    def __init__(self, init_bits=8.0, min_bits=2.0, max_bits=32.0, symmetric=True):
        super().__init__()
        self.soft_bits = nn.Parameter(torch.tensor(float(init_bits)))
        self.min_bits, self.max_bits, self.symmetric = min_bits, max_bits, symmetric
        self.range_scale = nn.Parameter(torch.tensor(0.0002))
        if symmetric:
            self.register_buffer("v_max", None)
        else:
            self.v_max = nn.Parameter(torch.tensor(0.0001))
        self.calibrated = False
        self.register_buffer("running_min", torch.tensor(float("inf")))
        self.register_buffer("running_max", torch.tensor(float("-inf")))

    # -- statistics ---------------------------------------------------------------------------------
    def calibrate_minmax(self, batch_min, batch_max):
        """quantization.py:97-119 given the batch min / max (0-d tensors)."""
        with torch.no_grad():
            self.running_min = torch.min(self.running_min, batch_min.to(self.running_min.dtype))
            self.running_max = torch.max(self.running_max, batch_max.to(self.running_max.dtype))
            if self.symmetric:
                self.range_scale.data = 2 * torch.max(torch.abs(self.running_min), torch.abs(self.running_max))
            else:
                self.range_scale.data = self.running_max - self.running_min
                self.v_max.data = self.running_max.clone()
            self.calibrated = True

    def calibrate(self, x):
        self.calibrate_minmax(x.min(), x.max())

    @property
    def bit_width(self):
        return torch.clamp(self.soft_bits, self.min_bits, self.max_bits)

    @property
    def integer_bit_width(self):
        return int(torch.round(self.bit_width).item())

    def get_quantization_params(self):
        B = self.integer_bit_width
        if self.symmetric:
            return -(2 ** (B - 1)), 2 ** (B - 1) - 1
        return 0, 2 ** B - 1

    # -- device-side scalars (no host sync) ------------------------------------------------------------
    def scalars(self, training=None):
        """(scale, zero_point, qmin, qmax) as 0-d device tensors, computed with the reference's
        expressions (quantization.py:121-175) but without ``.item()``."""
        training = self.training if training is None else training
        bw = self.bit_width
        b_int = torch.round(bw)
        if self.symmetric:
            qmin, qmax = -(2 ** (b_int - 1)), 2 ** (b_int - 1) - 1
        else:
            qmin, qmax = torch.zeros_like(b_int), 2 ** b_int - 1
        B = bw if training else b_int
        if self.symmetric:
            scale = self.range_scale / (2 ** (B - 1))
            zp = torch.zeros_like(scale)
        else:
            scale = torch.clamp(self.range_scale, min=1e-8) / (2 ** B - 1)
            zp = torch.round(torch.min(torch.max(self.v_max / scale, qmin), qmax))
        return scale.detach(), zp.detach(), qmin.detach(), qmax.detach()

    def qrow(self, training=None, enabled=True):
        """One PN_QROW row for the kernels: scale, scale+1e-8, zp, qmin, qmax, enabled, train-form, 0."""
        training = self.training if training is None else training
        scale, zp, qmin, qmax = self.scalars(training)
        one = torch.ones_like(scale)
        row = torch.stack([scale, scale + 1e-8, zp, qmin, qmax, one * float(enabled), one * float(training),
                           torch.zeros_like(scale)])
        assert row.numel() == QROW
        return row.float()

    def forward(self, x):
        if self.training and not self.calibrated:
            self.calibrate(x)
        scale, zp, qmin, qmax = self.scalars()
        q = torch.min(torch.max(torch.round(x / (scale + 1e-8) + zp), qmin), qmax)
        dq = (q - zp) * scale
        if self.training:
            return x + (dq - x).detach()
        return dq

    def extra_repr(self):
        return "soft_bits=%.2f, range=[%s, %s], symmetric=%s, range_scale=%.6f" % (
            float(self.soft_bits), self.min_bits, self.max_bits, self.symmetric, float(self.range_scale))


def qrows_batched(quantizers, training):
    """torch.stack([q.qrow(training) for q in quantizers]) computed with ONE set of vectorised ops over the levels
    (about 20 launches instead of 20 per quantiser: the per-call host cost of a quantised forward drops ~15x).
    Element for element the same fp32 operations as ``LearnedBitwidthQuantizer.scalars``, hence identical rows.
    All quantisers must share symmetric / min_bits / max_bits (they do: hash_encoding.py:37-44)."""
    q0 = quantizers[0]
    if any((q.symmetric, q.min_bits, q.max_bits) != (q0.symmetric, q0.min_bits, q0.max_bits) for q in quantizers):
        return torch.stack([q.qrow(training) for q in quantizers])
    soft = torch.stack([q.soft_bits.detach() for q in quantizers])
    rng = torch.stack([q.range_scale.detach() for q in quantizers])
    bw = torch.clamp(soft, q0.min_bits, q0.max_bits)
    b_int = torch.round(bw)
    if q0.symmetric:
        qmin, qmax = -(2 ** (b_int - 1)), 2 ** (b_int - 1) - 1
    else:
        qmin, qmax = torch.zeros_like(b_int), 2 ** b_int - 1
    B = bw if training else b_int
    if q0.symmetric:
        scale = rng / (2 ** (B - 1))
        zp = torch.zeros_like(scale)
    else:
        vmax = torch.stack([q.v_max.detach() for q in quantizers])
        scale = torch.clamp(rng, min=1e-8) / (2 ** B - 1)
        zp = torch.round(torch.min(torch.max(vmax / scale, qmin), qmax))
    one = torch.ones_like(scale)
    rows = torch.stack([scale, scale + 1e-8, zp, qmin, qmax, one, one * float(training), torch.zeros_like(scale)], 1)
    return rows.float()


class FakeQuantizer(nn.Module):
    """quantization.py:6-65 — fixed-bit variant (not instantiated by create_nerf; kept for API parity)."""

    def __init__(self, num_bits=8, symmetric=True, initialize_scale=True):
        super().__init__()
        self.num_bits, self.symmetric = num_bits, symmetric
        self.qmin, self.qmax = (-(2 ** (num_bits - 1)), 2 ** (num_bits - 1) - 1) if symmetric else (0, 2 ** num_bits - 1)
        self.scale = nn.Parameter(torch.tensor(1.0))
        if symmetric:
            self.register_buffer("zero_point", torch.tensor(0.0))
        else:
            self.zero_point = nn.Parameter(torch.tensor(0.0))

    def forward(self, x):
        xs = x / self.scale
        if not self.symmetric:
            xs = xs + self.zero_point
        q = torch.clamp(torch.round(xs), self.qmin, self.qmax)
        if self.training:
            dq = (q - self.zero_point) * self.scale
            return x + (dq - x).detach()
        return (q - self.zero_point) * self.scale

    def extra_repr(self):
        return "num_bits=%d, symmetric=%s" % (self.num_bits, self.symmetric)


class PassthroughQuantizer(nn.Module):
    """quantization.py:197-208."""

    def __init__(self, **kwargs):
        super().__init__()
        self.bit_width, self.integer_bit_width = 32.0, 32

    def forward(self, x):
        return x


def calculate_fqr(quantizers):
    """quantization.py:211-225 — mean bit-width."""
    if not quantizers:
        return 32.0
    total = 0
    for q in quantizers:
        total += q.bit_width if hasattr(q, "bit_width") else getattr(q, "num_bits", 32)
    return total / len(quantizers)
