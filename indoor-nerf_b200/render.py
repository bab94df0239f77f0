"""Drop-in for the renderer functions that live inside the reference's run_nerf.py: batchify,
run_network, batchify_rays, render, render_path, raw2outputs, render_rays (run_nerf.py:43-215, 347-549).

Same signatures, same return structure, same use of torch's global RNG; the arithmetic runs in the
sm_100a kernels (ops.py).  ``patch(run_nerf_module)`` installs them into an imported reference driver.
"""
import numpy as np
import torch

from . import ops
from .hash_encoding import HashEmbedder, SHEncoder
from .run_nerf_helpers import NeRFSmall, get_rays, ndc_rays, sample_pdf


def batchify(fn, chunk):
    """run_nerf.py:43-50."""
    if chunk is None:
        return fn

    def ret(inputs):
        return torch.cat([fn(inputs[i:i + chunk]) for i in range(0, inputs.shape[0], chunk)], 0)
    return ret


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """run_nerf.py:53-68.  With our HashEmbedder + SHEncoder + NeRFSmall the whole function is one
    autograd node (hash encode -> SH -> MLP -> keep mask); any other combination of callables goes
    through the generic composition below, still on the GPU."""
    if (isinstance(embed_fn, HashEmbedder) and isinstance(fn, NeRFSmall) and viewdirs is not None
            and isinstance(embeddirs_fn, SHEncoder) and inputs.dim() == 3):
        N, S = inputs.shape[0], inputs.shape[1]
        if embed_fn.training:
            embed_fn.current_step += 1                                   # hash_encoding.py:83-84
        if not embed_fn._is_flat():
            embed_fn._reflatten()
        pts = inputs.reshape(-1, 3)
        keys, weights = fn.kernel_weights()
        if embed_fn._packed is not None and not embed_fn.training and not fn.training:
            # inference from tables held as integer codes (HashEmbedder.pack_for_inference): no autograd
            w = {k: ops.fcontig(t.detach()) for k, t in zip(keys, weights)}
            act_q = fn.act_qrow(None, weights[0])
            if ops.get_mlp_mode() == "bf16" and embed_fn.n_levels == 16:
                out = ops.field_fwd_packed(embed_fn.grid(), embed_fn._packed, w, pts.detach(), viewdirs, S, act_q)
            else:
                feat, keep = ops.hash_encode_fwd_packed(embed_fn.grid(), embed_fn._packed, pts.detach())
                out = ops.mlp_fwd(w, feat, dirs=ops.fcontig(viewdirs), samples_per_ray=S, act_q=act_q, keep=keep)
            return out.reshape(N, S, out.shape[-1])
        qrows = embed_fn._quant_rows(pts.detach())
        act_q = None
        if fn.use_quantization and fn.sigma_act_quantizers is not None:
            feat0 = None
            if fn.training and not fn.sigma_act_quantizers[0].calibrated:
                # the reference calibrates on the first netchunk of the first training call (run_nerf.py:65)
                feat0, _ = ops.hash_encode_fwd(embed_fn.grid(), [t.detach() for t in embed_fn.tables()],
                                               pts.detach()[:netchunk], qrows)
            act_q = fn.act_qrow(feat0, weights[0])
        tables = embed_fn.tables()
        out = ops.FieldFn.apply(pts, viewdirs, S, embed_fn.grid(), qrows, act_q, keys, len(tables), *tables, *weights)
        return out.reshape(N, S, out.shape[-1])

    inputs_flat = torch.reshape(inputs, [-1, inputs.shape[-1]])
    embedded, keep_mask = embed_fn(inputs_flat)
    if viewdirs is not None:
        input_dirs = viewdirs[:, None].expand(inputs.shape)
        embedded_dirs = embeddirs_fn(torch.reshape(input_dirs, [-1, input_dirs.shape[-1]]))
        embedded = torch.cat([embedded, embedded_dirs], -1)
    outputs_flat = batchify(fn, netchunk)(embedded)
    outputs_flat = outputs_flat.clone()
    outputs_flat[~keep_mask, -1] = 0
    return torch.reshape(outputs_flat, list(inputs.shape[:-1]) + [outputs_flat.shape[-1]])


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, predict_normals=False):
    """run_nerf.py:347-411 -> (rgb_map, disp_map, acc_map, weights, depth_map, sparsity_loss[, normal_map])."""
    noise = None
    if raw_noise_std > 0.:
        noise = torch.randn(raw[..., 3].shape, device=raw.device) * raw_noise_std
        if pytest:
            np.random.seed(0)
            noise = torch.tensor(np.random.rand(*list(raw[..., 3].shape)) * raw_noise_std, dtype=torch.float32,
                                 device=raw.device)
    out = ops.CompositeFn.apply(raw, z_vals, rays_d, noise, bool(white_bkgd))
    if predict_normals:
        if raw.shape[-1] != 7:
            raise ValueError("predict_normals needs a 7-channel raw tensor (network with a normal head)")
        return out
    return out[:6]


def render_rays(ray_batch, network_fn, network_query_fn, N_samples, embed_fn=None, retraw=False, lindisp=False,
                perturb=0., N_importance=0, network_fine=None, white_bkgd=False, raw_noise_std=0., verbose=False,
                pytest=False, predict_normals=False):
    """run_nerf.py:414-549."""
    N_rays = ray_batch.shape[0]
    dev = ray_batch.device
    rays_o, rays_d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:] if ray_batch.shape[-1] > 8 else None
    near, far = ray_batch[:, 6], ray_batch[:, 7]

    t_vals = torch.linspace(0., 1., steps=N_samples, device=dev)
    t_rand = None
    if perturb > 0.:
        t_rand = torch.rand([N_rays, N_samples], device=dev)
        if pytest:
            np.random.seed(0)
            t_rand = torch.tensor(np.random.rand(N_rays, N_samples), dtype=torch.float32, device=dev)
    z_vals = ops.coarse_z(near, far, t_vals, t_rand, lindisp)
    pts = ops.make_points(rays_o, rays_d, z_vals)

    raw = network_query_fn(pts, viewdirs, network_fn)
    outs = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest=pytest, predict_normals=predict_normals)
    rgb_map, disp_map, acc_map, weights, depth_map, sparsity_loss = outs[:6]
    normal_map = outs[6] if predict_normals else None

    if N_importance > 0:
        rgb_map_0, depth_map_0, acc_map_0, sparsity_loss_0, normal_map_0 = rgb_map, depth_map, acc_map, sparsity_loss, normal_map
        z_vals_mid = .5 * (z_vals[..., 1:] + z_vals[..., :-1])
        z_samples = sample_pdf(z_vals_mid, weights[..., 1:-1], N_importance, det=(perturb == 0.), pytest=pytest)
        z_samples = z_samples.detach()
        z_vals = ops.sort_merge(z_vals, z_samples)
        pts = ops.make_points(rays_o, rays_d, z_vals)
        run_fn = network_fn if network_fine is None else network_fine
        raw = network_query_fn(pts, viewdirs, run_fn)
        outs = raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest=pytest, predict_normals=predict_normals)
        rgb_map, disp_map, acc_map, weights, depth_map, sparsity_loss = outs[:6]
        normal_map = outs[6] if predict_normals else None

    ret = {"rgb_map": rgb_map, "depth_map": depth_map, "acc_map": acc_map, "sparsity_loss": sparsity_loss}
    ret["pts"] = pts
    ret["rays_d"] = rays_d
    if predict_normals:
        ret["normal_map"] = normal_map
    if retraw:
        ret["raw"] = raw
    if N_importance > 0:
        ret["rgb0"] = rgb_map_0
        ret["depth0"] = depth_map_0
        ret["acc0"] = acc_map_0
        ret["sparsity_loss0"] = sparsity_loss_0
        ret["z_std"] = torch.std(z_samples, dim=-1, unbiased=False)
        if predict_normals:
            ret["normal0"] = normal_map_0
    return ret


def batchify_rays(rays_flat, chunk=1024 * 32, **kwargs):
    """run_nerf.py:71-83."""
    all_ret = {}
    for i in range(0, rays_flat.shape[0], chunk):
        ret = render_rays(rays_flat[i:i + chunk], **kwargs)
        for k in ret:
            all_ret.setdefault(k, []).append(ret[k])
    return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in all_ret.items()}


def render(H, W, K, chunk=1024 * 32, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False,
           c2w_staticcam=None, **kwargs):
    """run_nerf.py:86-151 -> [rgb_map, depth_map, acc_map, extras]."""
    if c2w is not None:
        rays_o, rays_d = get_rays(H, W, K, c2w)
    else:
        rays_o, rays_d = rays
    viewdirs = None
    if use_viewdirs:
        viewdirs = rays_d
        if c2w_staticcam is not None:
            rays_o, rays_d = get_rays(H, W, K, c2w_staticcam)
        viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)
        viewdirs = torch.reshape(viewdirs, [-1, 3]).float()
    sh = rays_d.shape
    if ndc:
        rays_o, rays_d = ndc_rays(H, W, K[0][0], 1., rays_o, rays_d)
    rays_o = torch.reshape(rays_o, [-1, 3]).float()
    rays_d = torch.reshape(rays_d, [-1, 3]).float()
    near, far = near * torch.ones_like(rays_d[..., :1]), far * torch.ones_like(rays_d[..., :1])
    parts = [rays_o, rays_d, near, far]
    if use_viewdirs:
        parts.append(viewdirs)
    rays = torch.cat(parts, -1)

    all_ret = batchify_rays(rays, chunk, **kwargs)
    for k in all_ret:
        all_ret[k] = torch.reshape(all_ret[k], list(sh[:-1]) + list(all_ret[k].shape[1:]))
    k_extract = ["rgb_map", "depth_map", "acc_map"]
    return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]


def _write_png(path, img8):
    """Minimal PNG writer (8-bit gray [H,W] or RGB [H,W,3]); the reference saves matplotlib figures
    (run_nerf.py:190-205), which is plotting and out of scope — the pixels are what matters here."""
    import struct
    import zlib
    img8 = np.ascontiguousarray(img8)
    h, w = img8.shape[:2]
    color = 2 if img8.ndim == 3 else 0
    raw = b"".join(b"\x00" + img8[r].tobytes() for r in range(h))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, color, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


_STAGING = {}


def _frame_staging(H, W, slots=2):
    """Pinned host slots for finished frames, kept per frame size: cudaHostAlloc of a test set's worth of frames cost more
    than rendering one of them (27 of 73 ms per 800x800 frame in round 2's render_path leg)."""
    key = (H, W, slots)
    if key not in _STAGING:
        _STAGING.clear()
        _STAGING[key] = [(torch.empty((H, W, 3), dtype=torch.float32, pin_memory=True),
                          torch.empty((H, W), dtype=torch.float32, pin_memory=True)) for _ in range(slots)]
    return _STAGING[key]


def render_path(render_poses, hwf, K, chunk, render_kwargs, gt_imgs=None, savedir=None, render_factor=0, group=None):
    """run_nerf.py:154-215 -> (rgbs [n,H,W,3], depths [n,H,W]) as numpy arrays, depths normalised by near/far.

    Differences in HOW, not WHAT: the squared error of every frame is reduced on the device (pn_image_sqerr) and
    all PSNRs are read back with one transfer at the end instead of a host round trip per frame; frames go to
    two cached pinned slots asynchronously and from there into the result arrays while the next frame renders; with a process group every frame is rendered pixel-sharded
    (row blocks, no collective on the data path, one all-gather of the finished frame).  Saved images are plain
    PNGs of to8b(rgb) and of the normalised depth."""
    import os
    import pickle
    from . import parallel

    H, W, focal = hwf
    near, far = render_kwargs["near"], render_kwargs["far"]
    if render_factor != 0:
        H, W, focal = H // render_factor, W // render_factor, focal / render_factor
    H, W = int(H), int(W)
    dev = torch.device("cuda", torch.cuda.current_device())
    n = len(render_poses)
    # group=None means "this process renders whole frames", even inside an initialised process group
    world = parallel.world_size(group) if group is not None else 1
    rank = parallel.rank(group) if group is not None else 0
    rgbs = np.empty((n, H, W, 3), dtype=np.float32)
    depths = np.empty((n, H, W), dtype=np.float32)
    stage = _frame_staging(H, W)
    pending = []                                     # (frame, slot, event) whose device->host copy is in flight

    def drain(keep):
        # finished frames: pinned slot -> the result arrays, on the host while the GPU renders the next frame
        while len(pending) > keep:
            j, slot, ev = pending.pop(0)
            ev.synchronize()
            np.copyto(rgbs[j], stage[slot][0].numpy())
            np.copyto(depths[j], stage[slot][1].numpy())

    want_psnr = gt_imgs is not None and render_factor == 0
    sq = torch.zeros(n, dtype=torch.float64, device=dev) if want_psnr else None
    rgb8s = []
    with torch.no_grad():
        for i, c2w in enumerate(render_poses):
            drain(len(stage) - 1)                    # the slot about to be reused has been copied out
            c2w = torch.as_tensor(c2w)[:3, :4]
            if world > 1:
                ro, rd = get_rays(H, W, K, c2w)
                kw = {k: v for k, v in render_kwargs.items()}
                rgb, depth, _ = parallel.render_sharded(lambda h, w, **a: render(h, w, K, chunk=chunk, **a), H, W, ro, rd,
                                                        group=group, gather=True, **kw)
            else:
                rgb, depth, _, _ = render(H, W, K, chunk=chunk, c2w=c2w, **render_kwargs)
            slot = i % len(stage)
            stage[slot][0].copy_(rgb, non_blocking=True)
            depth = (depth - near) / (far - near)
            stage[slot][1].copy_(depth, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            pending.append((i, slot, ev))
            if want_psnr:
                gt = torch.as_tensor(gt_imgs[i]).to(dev, non_blocking=True).float()
                ops.image_sqerr(rgb, gt[..., :3], out=sq[i])
            if savedir is not None:
                rgb8s.append((ops.to8b(rgb).cpu(), ops.to8b(depth).cpu()))
    drain(0)
    torch.cuda.synchronize(dev)
    if savedir is not None and rank == 0:
        os.makedirs(savedir, exist_ok=True)
        for i, (c8, d8) in enumerate(rgb8s):
            _write_png(os.path.join(savedir, "{:03d}.png".format(i)), c8.numpy())
            _write_png(os.path.join(savedir, "depth_{:03d}.png".format(i)), d8.numpy())
    if want_psnr:
        psnrs = (-10. * torch.log10(sq / float(H * W * 3))).cpu().tolist()
        avg_psnr = sum(psnrs) / len(psnrs)
        print("Avg PSNR over Test set: ", avg_psnr)
        if savedir is not None and rank == 0:
            with open(os.path.join(savedir, "test_psnrs_avg{:0.2f}.pkl".format(avg_psnr)), "wb") as fp:
                pickle.dump(psnrs, fp)
        render_path.last_psnrs = psnrs
    return rgbs, depths


def patch(run_nerf_module):
    """Install the renderer into an imported reference driver (`import run_nerf; patch(run_nerf)`):
    the driver's train()/render_path() then run on the kernels.  See INTEGRATION.md."""
    for name in ("batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays", "render_path"):
        setattr(run_nerf_module, name, globals()[name])
    return run_nerf_module
