"""create_nerf (run_nerf.py:218-344) for the HashNeRF configuration (i_embed=1, i_embed_views=2): builds
the embedders, the coarse and fine NeRFSmall, the RAdam parameter groups and the render kwargs with the
reference's keys, so that a driver written against the reference finds the same objects."""
import os
from types import SimpleNamespace

import torch

from .radam import RAdam
from .render import run_network
from .run_nerf_helpers import NeRFSmall, get_embedder


def default_args(**over):
    """The hot-path subset of config_parser's defaults (run_nerf.py:552-715) + configs/chair.txt."""
    a = dict(multires=10, multires_views=4, i_embed=1, i_embed_views=2, use_viewdirs=True, N_samples=64,
             N_importance=128, perturb=1.0, raw_noise_std=0.0, white_bkgd=True, lindisp=False, no_ndc=False,
             dataset_type="blender", netchunk=1024 * 64, chunk=1024 * 32, lrate=0.01, lrate_decay=10,
             log2_hashmap_size=19, finest_res=512, use_quantization=False, quantization_bits=8,
             predict_normals=False, sparse_loss_weight=1e-10, tv_loss_weight=1e-6, bounding_box=None,
             basedir=None, expname=None, ft_path=None, no_reload=True)
    a.update(over)
    return SimpleNamespace(**a)


def create_nerf(args, device="cuda"):
    """run_nerf.py:218-344 -> (render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer)."""
    if args.i_embed != 1:
        raise NotImplementedError("only the hash-grid configuration (i_embed=1) is on this path")
    use_q, q_bits = getattr(args, "use_quantization", False), getattr(args, "quantization_bits", 8)
    embed_fn, input_ch = get_embedder(args.multires, args, i=args.i_embed)
    embed_fn = embed_fn.to(device)
    embedding_params = list(embed_fn.parameters())
    input_ch_views, embeddirs_fn = 0, None
    if args.use_viewdirs:
        embeddirs_fn, input_ch_views = get_embedder(args.multires_views, args, i=args.i_embed_views)

    def small(normals):
        return NeRFSmall(num_layers=2, hidden_dim=64, geo_feat_dim=15, num_layers_color=3, hidden_dim_color=64,
                         input_ch=input_ch, input_ch_views=input_ch_views, use_quantization=use_q,
                         quantization_bits=q_bits, predict_normals=normals).to(device)

    normals = bool(getattr(args, "predict_normals", False))
    # the reference passes predict_normals to the fine net only (run_nerf.py:260-268); raw2outputs is then
    # called with predict_normals for the coarse pass too, which needs 7 channels -> give both nets the head
    model = small(normals)
    grad_vars = list(model.parameters())
    model_fine = None
    if args.N_importance > 0:
        model_fine = small(normals)
        grad_vars += list(model_fine.parameters())

    def network_query_fn(inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn,
                           netchunk=args.netchunk)

    optimizer = RAdam([{"params": grad_vars, "weight_decay": 1e-6}, {"params": embedding_params, "eps": 1e-15}],
                      lr=args.lrate, betas=(0.9, 0.99))                                   # run_nerf.py:281-285
    start = 0
    ckpts = []
    if getattr(args, "ft_path", None) not in (None, "None"):
        ckpts = [args.ft_path]
    elif getattr(args, "basedir", None) and getattr(args, "expname", None):
        d = os.path.join(args.basedir, args.expname)
        if os.path.isdir(d):
            ckpts = [os.path.join(d, f) for f in sorted(os.listdir(d)) if "tar" in f]
    if ckpts and not getattr(args, "no_reload", False):                                    # run_nerf.py:296-315
        ckpt = torch.load(ckpts[-1], map_location=device)
        start = ckpt["global_step"]
        optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        model.load_state_dict(ckpt["network_fn_state_dict"])
        if model_fine is not None:
            model_fine.load_state_dict(ckpt["network_fine_state_dict"])
        embed_fn.load_state_dict(ckpt["embed_fn_state_dict"])

    if torch.device(device).type == "cuda":
        # one flat gradient buffer for the tables and both networks (ops.GradArena): the fused backward accumulates
        # into it, a data-parallel step all-reduces it in one call
        from . import ops
        mlp_leaves = [p for net in (model, model_fine) if net is not None for k, p in net.named_parameters()
                      if "quantizer" not in k]
        embed_fn.grad_arena = ops.GradArena(embed_fn.tables(), mlp_leaves)

    render_kwargs_train = {
        "network_query_fn": network_query_fn, "perturb": args.perturb, "N_importance": args.N_importance,
        "network_fine": model_fine, "N_samples": args.N_samples, "network_fn": model, "embed_fn": embed_fn,
        "use_viewdirs": args.use_viewdirs, "white_bkgd": args.white_bkgd, "raw_noise_std": args.raw_noise_std,
        "predict_normals": normals,
    }
    if args.dataset_type != "llff" or args.no_ndc:
        render_kwargs_train["ndc"] = False
        render_kwargs_train["lindisp"] = args.lindisp
    render_kwargs_test = dict(render_kwargs_train)
    render_kwargs_test["perturb"] = False
    render_kwargs_test["raw_noise_std"] = 0.
    return render_kwargs_train, render_kwargs_test, start, grad_vars, optimizer


def save_checkpoint(path, global_step, render_kwargs_train, optimizer):
    """The reference's checkpoint dictionary (run_nerf.py:1345-1362)."""
    d = {"global_step": global_step,
         "network_fn_state_dict": render_kwargs_train["network_fn"].state_dict(),
         "embed_fn_state_dict": render_kwargs_train["embed_fn"].state_dict(),
         "optimizer_state_dict": optimizer.state_dict()}
    if render_kwargs_train["network_fine"] is not None:
        d["network_fine_state_dict"] = render_kwargs_train["network_fine"].state_dict()
    torch.save(d, path)
