"""Data-parallel plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch on the
GPUs, gloo in the CPU tests).

Training shards the ray batch across ranks; parameters are replicated; after backward there is exactly
one exchange: a SUM all-reduce of the hash-table gradient (one flat [L,T,2] buffer, 64 MiB at T=2^19) and
one of the concatenated MLP gradients (~75 KB).  Rendering shards pixels (row blocks) and needs no
collective except the optional gather of the finished image.  The reference has no distributed code at
all; this module is what `train()` / `render_path()` call once per step / frame."""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def rank(group=None):
    return dist.get_rank(group) if (dist.is_available() and dist.is_initialized()) else 0


def shard_rays(batch_rays, target_s, rank_, world):
    """Contiguous shard of a [2,N,3] ray batch and its [N,3] targets (N divisible by world not required)."""
    N = batch_rays.shape[1]
    lo, hi = (N * rank_) // world, (N * (rank_ + 1)) // world
    return batch_rays[:, lo:hi], target_s[lo:hi]


def flat_table_grad(embed_fn):
    """The table gradients as ONE tensor.  The backward kernels write all levels into one [L,T,2] buffer and
    hand autograd per-level views of it, so normally this is a zero-copy re-assembly; if something
    re-materialised the grads (e.g. a second accumulation into fresh tensors) they are packed and re-pointed."""
    ws = [e.weight for e in embed_fn.embeddings]
    if any(w.grad is None for w in ws):
        return None
    g0 = ws[0].grad
    n = g0.numel()
    esz = g0.element_size()
    same = all(w.grad.is_contiguous() and w.grad.dtype == g0.dtype and w.grad.device == g0.device and
               w.grad.data_ptr() == g0.data_ptr() + l * n * esz for l, w in enumerate(ws))
    if same and g0.untyped_storage().nbytes() - g0.storage_offset() * esz >= len(ws) * n * esz:
        return torch.as_strided(g0, (len(ws),) + tuple(g0.shape), (n,) + tuple(g0.stride()))
    flat = torch.stack([w.grad for w in ws])
    for l, w in enumerate(ws):
        w.grad = flat[l]
    return flat


def allreduce_gradients(embed_fn, nets, group=None, async_op=False):
    """SUM all-reduce of every gradient the step produced.  With the model's gradient arena (ops.GradArena, built by
    create_nerf) that is ONE collective over one flat buffer — table gradients and the weights of both networks — with
    no packing and no copy-back; gradients outside the arena (none in the stock configuration) go in a second, small
    call.  Without an arena: one call for the flat table gradient, one for everything else.
    Returns the list of work handles when async_op (the caller waits before the optimiser step)."""
    works = []
    arena = getattr(embed_fn, "grad_arena", None)
    covered = set()
    if arena is not None and arena.valid() and any(p.grad is not None for p in arena.params()):
        arena.ensure()
        works.append(dist.all_reduce(arena.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op))
        covered = {id(p) for p in arena.params()}
    else:
        flat = flat_table_grad(embed_fn)
        if flat is not None:
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op))
        covered = {id(e.weight) for e in embed_fn.embeddings}
    small = [p for m in nets for p in m.parameters() if p.grad is not None and id(p) not in covered]
    small += [p for p in embed_fn.parameters() if p.grad is not None and id(p) not in covered]
    if small:
        buf = torch.cat([p.grad.reshape(-1) for p in small])
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for p in small:
            k = p.grad.numel()
            p.grad.copy_(buf[off:off + k].view_as(p.grad))
            off += k
    return [w for w in works if w is not None] if async_op else None


def broadcast_parameters(modules, src=0, group=None):
    """Make every rank start from rank `src`'s parameters and buffers (tables in one call when flat)."""
    for m in modules:
        storage = getattr(m, "table_storage", None)
        done = set()
        if storage is not None and m._is_flat():
            dist.broadcast(storage, src=src, group=group)
            done = {id(e.weight) for e in m.embeddings}
        for t in list(m.parameters()) + list(m.buffers()):
            if id(t) not in done and t is not None:
                dist.broadcast(t.data, src=src, group=group)


def sync_quantizer_calibration(quantizers, group=None):
    """Calibration statistics come from the local batch (quantization.py:97-119); make them the global
    min / max so that every rank fake-quantises identically.  One MIN and one MAX collective for all quantisers."""
    qs = [q for q in quantizers if q is not None]
    if not qs:
        return
    lo = torch.stack([q.running_min.reshape(()) for q in qs])
    hi = torch.stack([q.running_max.reshape(()) for q in qs])
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    for i, q in enumerate(qs):
        q.running_min = lo[i].clone()
        q.running_max = hi[i].clone()
        q.calibrate_minmax(q.running_min, q.running_max)


def model_quantizers(embed_fn, nets):
    """Every quantiser whose calibration depends on the batch: table levels, activation and weight quantisers."""
    qs = list(embed_fn.quantizers) if getattr(embed_fn, "quantizers", None) is not None else []
    for n in nets:
        if getattr(n, "sigma_act_quantizers", None) is not None:
            qs += list(n.sigma_act_quantizers)
        if getattr(n, "sigma_weight_quantizer", None) is not None:
            qs.append(n.sigma_weight_quantizer)
    return qs


def pixel_rows(H, rank_, world):
    """Row block [lo, hi) of an H-row image owned by a rank."""
    return (H * rank_) // world, (H * (rank_ + 1)) // world


def render_sharded(render_fn, H, W, rays_o, rays_d, group=None, gather=True, **kwargs):
    """Pixel-sharded test-view rendering: each rank renders its row block of the [H,W,3] ray image with
    `render_fn(rays=(o, d), ...)` (no collective); with gather=True the blocks are all-gathered so every
    rank returns the full frame [rgb, depth, acc]."""
    world, r = world_size(group), rank(group)
    lo, hi = pixel_rows(H, r, world)
    out = render_fn(H, W, rays=(rays_o[lo:hi], rays_d[lo:hi]), **kwargs)
    rgb, depth, acc = out[0], out[1], out[2]
    if world == 1 or not gather:
        return rgb, depth, acc
    rows = [pixel_rows(H, k, world) for k in range(world)]
    equal = len({b - a for a, b in rows}) == 1
    res = []
    for t in (rgb, depth, acc):
        t = t.contiguous()
        parts = [torch.empty((b - a,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for a, b in rows]
        if equal:
            dist.all_gather(parts, t, group=group)
        else:                                   # ragged row blocks: one broadcast per owner
            parts[r].copy_(t)
            for k in range(world):
                dist.broadcast(parts[k], src=k, group=group)
        res.append(torch.cat(parts, 0))
    return tuple(res)
