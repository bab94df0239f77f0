"""Export / import of A-CAQ-quantised models at their learned integer bit-widths.

The reference only ever *fake*-quantises: checkpoints are fp32 state_dicts (run_nerf.py:1345-1362) and the
"compressed" model exists as a number in a log (quantization.py:211-225, run_nerf.py:1412-1425).  This module
makes the compression real: every hash table is stored as the integer codes of its level's
LearnedBitwidthQuantizer in eval form (quantization.py:157-187), bit-packed at round(clamp(soft_bits)) bits per
value by ``pn_quant_pack``; the first sigma layer's weight likewise with its symmetric quantiser
(run_nerf_helpers.py:272-276); everything else stays fp32.  ``pn_quant_unpack`` reproduces, bit for bit, the
values the quantisers' eval forward yields, so a model loaded from the file renders exactly what the
fake-quantised model renders in eval mode — with the table quantisers switched off, because the reference's
fake-quant is not idempotent (the 1e-8 in ``x / (scale + 1e-8)`` is ~1 % of a table-sized scale).

File layout (little endian):  b"PNQ1" | u32 header_len | header (JSON, utf-8) | zero padding to 16 B | payload.
The header lists, per tensor: name, shape, storage ("packed" | "fp32"), and for packed tensors bits, scale, zp,
qmin, qmax (exact float32 values) plus offset / nbytes of its words in the payload.
"""
import json
import struct

import numpy as np
import torch

from . import ops
from ._lib import QROW

MAGIC = b"PNQ1"
MAX_PACKED_BITS = 24      # qmax = 2^B - 1 must be exact in fp32; wider levels are stored as fp32


def _eval_row(q):
    """The eval-form kernel row of a LearnedBitwidthQuantizer and its integer width."""
    row = q.qrow(training=False).contiguous()
    assert row.numel() == QROW
    return row, q.integer_bit_width


def _entry_packed(name, t, q):
    row, bits = _eval_row(q)
    n = t.numel()
    if bits > MAX_PACKED_BITS or n % 32 != 0:
        with torch.no_grad():
            was = q.training
            q.eval()
            y = q(t.detach())
            q.train(was)
        return {"name": name, "shape": list(t.shape), "storage": "fp32", "note": "eval fake-quant applied (bits=%d)" % bits}, \
            y.float().contiguous().reshape(-1)
    words = ops.quant_pack(t.detach(), row, bits)
    r = row.cpu().tolist()
    meta = {"name": name, "shape": list(t.shape), "storage": "packed", "bits": bits, "scale": r[0], "zp": r[2],
            "qmin": r[3], "qmax": r[4]}
    return meta, words


def export_quantized(path, embed_fn, networks=None, extra=None):
    """Write a .pnq file.  embed_fn: HashEmbedder (quantised levels are packed when its quantisers are calibrated,
    otherwise stored fp32); networks: dict name -> NeRFSmall.  Returns the header (with 'bytes' totals)."""
    networks = networks or {}
    entries, blobs = [], []

    def add(meta, data):
        entries.append(meta)
        blobs.append(data)

    use_q = bool(embed_fn.use_quantization and embed_fn.quantizers is not None and
                 all(q.calibrated for q in embed_fn.quantizers))
    for l, e in enumerate(embed_fn.embeddings):
        name = "embed_fn.embeddings.%d.weight" % l
        if use_q:
            add(*_entry_packed(name, e.weight, embed_fn.quantizers[l]))
        else:
            add({"name": name, "shape": list(e.weight.shape), "storage": "fp32"}, e.weight.detach().float().reshape(-1))
    for prefix, net in networks.items():
        wq = getattr(net, "sigma_weight_quantizer", None)
        for k, v in net.state_dict().items():
            name = "%s.%s" % (prefix, k)
            if k == "sigma_net.0.weight" and net.use_quantization and wq is not None and wq.calibrated:
                add(*_entry_packed(name, v, wq))
            elif k.startswith("sigma_weight_quantizer."):
                continue                                   # folded into the packed weight
            else:
                add({"name": name, "shape": list(v.shape), "storage": "fp32"}, v.detach().float().reshape(-1))
    off = 0
    payload = []
    for meta, data in zip(entries, blobs):
        raw = data.contiguous().cpu().numpy().tobytes()
        meta["offset"], meta["nbytes"] = off, len(raw)
        pad = (-len(raw)) % 16
        payload.append(raw + b"\0" * pad)
        off += len(raw) + pad
    fp32_bytes = sum(4 * int(np.prod(m["shape"])) for m in entries)
    header = {"version": 1, "tensors": entries, "table_quantisation": use_q,
              "bytes": {"payload": off, "fp32_equivalent": fp32_bytes},
              "embedder": {"n_levels": embed_fn.n_levels, "log2_hashmap_size": embed_fn.log2_hashmap_size,
                           "base_resolution": int(embed_fn.base_resolution), "finest_resolution": int(embed_fn.finest_resolution),
                           "bounding_box": [torch.as_tensor(b).float().cpu().tolist() for b in embed_fn.bounding_box]},
              "extra": extra or {}}
    hj = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(MAGIC + struct.pack("<I", len(hj)) + hj)
        f.write(b"\0" * ((-(8 + len(hj))) % 16))
        for p in payload:
            f.write(p)
    return header


def read_quantized(path, device):
    """-> (header, dict name -> fp32 CUDA tensor with the dequantised values)."""
    with open(path, "rb") as f:
        blob = f.read()
    if blob[:4] != MAGIC:
        raise ValueError("%s is not a PNQ1 file" % path)
    (hl,) = struct.unpack("<I", blob[4:8])
    header = json.loads(blob[8:8 + hl].decode())
    base = 8 + hl + ((-(8 + hl)) % 16)
    out = {}
    for m in header["tensors"]:
        raw = np.frombuffer(blob, dtype=np.uint8, count=m["nbytes"], offset=base + m["offset"])
        n = int(np.prod(m["shape"]))
        if m["storage"] == "packed":
            words = torch.from_numpy(raw.view("<i4").copy()).to(device)
            scale = np.float32(m["scale"])
            row = torch.tensor([scale, np.float32(scale + np.float32(1e-8)), m["zp"], m["qmin"], m["qmax"], 1, 0, 0],
                               dtype=torch.float32, device=device)
            out[m["name"]] = ops.quant_unpack(words, n, row, m["bits"]).reshape(m["shape"])
        else:
            out[m["name"]] = torch.from_numpy(raw.view("<f4").copy()).to(device).reshape(m["shape"])
    return header, out


def packed_levels_from_file(path, device, header=None):
    """ops.PackedLevels built straight from the file's bit streams (u8 codes up to 8 bits, u16 up to 16, fp32 above):
    what HashEmbedder gathers from in eval mode after load_quantized(..., packed=True)."""
    with open(path, "rb") as f:
        blob = f.read()
    (hl,) = struct.unpack("<I", blob[4:8])
    header = header or json.loads(blob[8:8 + hl].decode())
    base = 8 + hl + ((-(8 + hl)) % 16)
    levels = []
    for m in header["tensors"]:
        if not m["name"].startswith("embed_fn.embeddings."):
            continue
        raw = np.frombuffer(blob, dtype=np.uint8, count=m["nbytes"], offset=base + m["offset"])
        n = int(np.prod(m["shape"]))
        if m["storage"] == "packed":
            words = torch.from_numpy(raw.view("<i4").copy()).to(device)
            if m["bits"] <= 16:
                t = ops.quant_unpack_codes(words, n, m["bits"], 1 if m["bits"] <= 8 else 2).reshape(m["shape"])
            else:
                scale = np.float32(m["scale"])
                row = torch.tensor([scale, np.float32(scale + np.float32(1e-8)), m["zp"], m["qmin"], m["qmax"], 1, 0, 0],
                                   dtype=torch.float32, device=device)
                t = ops.quant_unpack(words, n, row, m["bits"]).reshape(m["shape"])
            levels.append((t, m["scale"], m["zp"], m["qmin"]))
        else:
            levels.append((torch.from_numpy(raw.view("<f4").copy()).to(device).reshape(m["shape"]), 1.0, 0.0, 0.0))
    return ops.PackedLevels(levels)


def load_quantized(path, embed_fn, networks=None, packed=False):
    """Fill embed_fn / networks from a .pnq file.  Tables (and the first sigma weight) receive the dequantised
    values and their quantisers are switched off, so eval-mode output equals the exporting model's.  With
    packed=True the embedder additionally keeps the levels as integer codes and gathers from those in eval mode
    (same output, 2-4 bytes per corner instead of 8)."""
    networks = networks or {}
    dev = embed_fn.embeddings[0].weight.device
    header, tensors = read_quantized(path, dev)
    with torch.no_grad():
        for l, e in enumerate(embed_fn.embeddings):
            e.weight.copy_(tensors["embed_fn.embeddings.%d.weight" % l])
    if header["table_quantisation"]:
        embed_fn.use_quantization = False                   # hash_encoding.py:97 — values are already dequantised
        if packed:
            embed_fn.eval()
            embed_fn.set_packed(packed_levels_from_file(path, dev, header))
    for prefix, net in networks.items():
        sd = {k[len(prefix) + 1:]: v for k, v in tensors.items() if k.startswith(prefix + ".")}
        packed_w0 = any(m["name"] == prefix + ".sigma_net.0.weight" and m["storage"] == "packed" for m in header["tensors"])
        missing = net.load_state_dict(sd, strict=False)
        bad = [k for k in missing.missing_keys if not k.startswith("sigma_weight_quantizer.")]
        if bad or missing.unexpected_keys:
            raise KeyError("state mismatch for %s: missing %s unexpected %s" % (prefix, bad, missing.unexpected_keys))
        if packed_w0:
            net.sigma_weight_quantizer = None               # run_nerf_helpers.py:272-276 — weight already dequantised
        for q in (net.sigma_act_quantizers or []):
            q.calibrated = True
    return header
