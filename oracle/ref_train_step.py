"""TEST INFRASTRUCTURE — one training iteration of the UNMODIFIED reference (run_nerf.py:1007-1162, chair config),
driven through oracle/ref_shim.py.  This is what `bench.py --impl reference` and bench.py's `cpu_baseline` /
`reference_eager_b200` legs time; the product package never imports it.

Every arithmetic line that runs is the reference's own: run_nerf.render -> batchify_rays -> render_rays ->
run_network -> HashEmbedder / SHEncoder / NeRFSmall -> raw2outputs -> sample_pdf, img2mse, the sparsity term,
loss.total_variation_loss on the 16 levels, loss.backward(), radam.RAdam.step().  What this file adds is only what
`train()` does around that body and cannot be called in isolation (it is one 800-line function): building the models
with the create_nerf shapes (create_nerf itself raises at run_nerf.py:260-268) and the two RAdam
parameter groups (run_nerf.py:281-285).

    python -m oracle.ref_train_step --device cpu  --rays 1024  --steps 5 --warmup 1     # one JSON line
    python -m oracle.ref_train_step --device cuda --rays 65536 --steps 2 --warmup 1     # torch.set_default_tensor_type
                                                                                        # ('torch.cuda.FloatTensor'), run_nerf.py:1486
"""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def build(ref, bounding_box, device, log2T=19, finest=512, lrate=0.01, predict_normals=False, use_quantization=False,
          quantization_bits=8):
    """The objects create_nerf builds (run_nerf.py:218-344) for i_embed=1, use_viewdirs, N_importance>0."""
    import torch
    emb = ref.hash_encoding.HashEmbedder(bounding_box=bounding_box, log2_hashmap_size=log2T, finest_resolution=finest,
                                         use_quantization=use_quantization, quantization_bits=quantization_bits).to(device)
    sh = ref.hash_encoding.SHEncoder()
    from oracle.ref_shim import make_nerf_small
    kw = dict(use_quantization=use_quantization, quantization_bits=quantization_bits)
    net = make_nerf_small(ref, predict_normals, **kw).to(device)
    net_fine = make_nerf_small(ref, predict_normals, **kw).to(device)
    grad_vars = list(net.parameters()) + list(net_fine.parameters())
    RN = ref.run_nerf
    query = lambda inputs, viewdirs, network_fn: RN.run_network(inputs, viewdirs, network_fn, embed_fn=emb,
                                                                embeddirs_fn=sh, netchunk=1024 * 64)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt = ref.radam.RAdam([{"params": grad_vars, "weight_decay": 1e-6},
                               {"params": list(emb.parameters()), "eps": 1e-15}], lr=lrate, betas=(0.9, 0.99))
    kwargs = {"network_query_fn": query, "perturb": 1.0, "N_importance": 128, "network_fine": net_fine, "N_samples": 64,
              "network_fn": net, "embed_fn": emb, "use_viewdirs": True, "white_bkgd": True, "raw_noise_std": 0.0,
              "predict_normals": predict_normals, "ndc": False, "lindisp": False}
    return kwargs, opt


def train_step(ref, kwargs, opt, H, W, K, batch_rays, target_s, near=2.0, far=6.0, chunk=1024 * 32,
               sparse_loss_weight=1e-10, tv_loss_weight=1e-6):
    """run_nerf.py:1007-1037 + 1159-1160, statement for statement."""
    RN = ref.run_nerf
    rgb, depth, acc, extras = RN.render(H, W, K, chunk=chunk, rays=batch_rays, verbose=False, retraw=True,
                                        near=near, far=far, **kwargs)
    opt.zero_grad()
    img_loss = RN.img2mse(rgb, target_s)
    loss = img_loss
    if "rgb0" in extras:
        loss = loss + RN.img2mse(extras["rgb0"], target_s)
    loss = loss + sparse_loss_weight * (extras["sparsity_loss"].sum() + extras["sparsity_loss0"].sum())
    emb = kwargs["embed_fn"]
    tv = sum(ref.loss.total_variation_loss(emb.embeddings[i], emb.base_resolution, emb.finest_resolution, i,
                                           emb.log2_hashmap_size, n_levels=emb.n_levels) for i in range(emb.n_levels))
    loss = loss + tv_loss_weight * tv
    loss.backward()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        opt.step()
    return loss


def run(device, rays, steps, warmup, threads=None):
    """Times `steps` iterations after `warmup`; returns a dict with the median and mean step time."""
    import numpy as np
    import torch
    from oracle import ref_shim
    from indoor_nerf_b200 import synthetic          # scene / ray-batch generator only (numpy + torch, no kernels)
    if device == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    ref = ref_shim.load()
    if device == "cuda":
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            torch.set_default_tensor_type("torch.cuda.FloatTensor")           # run_nerf.py:1486
    dev = torch.device(device)
    torch.manual_seed(0)
    np.random.seed(0)
    scene = synthetic.blender_scene(400, 400, n_views=100)
    box = tuple(t.to(dev) for t in scene["bounding_box"])
    kwargs, opt = build(ref, box, dev)
    batches = [synthetic.ray_batch(scene, rays, seed=100 + i, device=dev) for i in range(2)]
    sync = torch.cuda.synchronize if device == "cuda" else (lambda: None)
    times, loss = [], None
    for i in range(warmup + steps):
        r, t = batches[i % 2]
        sync()
        t0 = time.perf_counter()
        loss = train_step(ref, kwargs, opt, scene["H"], scene["W"], scene["K"], r, t, near=scene["near"], far=scene["far"])
        sync()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    return {"device": device, "rays_per_step": rays, "steps": steps, "warmup": warmup, "s_per_step_median": med,
            "s_per_step_mean": float(np.mean(times)), "rays_per_s": rays / med, "final_loss": float(loss.detach()),
            "threads": torch.get_num_threads() if device == "cpu" else None, "reference_root": ref_shim.REF_ROOT,
            "cpu_model": _cpu_model()}


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--rays", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()
    if a.device == "cpu":
        os.environ["CUDA_VISIBLE_DEVICES"] = ""       # the reference picks cuda whenever it is visible (run_nerf.py:38)
    res = run(a.device, a.rays, a.steps, a.warmup, a.threads or None)
    sys.stdout.flush()
    os.write(1, (json.dumps(res) + "\n").encode())


if __name__ == "__main__":
    main()
