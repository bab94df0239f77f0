"""Training equivalence of the arithmetic modes (-m gpu).

north_star allows the bf16 tensor-core NeRFSmall at 2e-3; bench.py's headline is measured in that mode.  Per-op and
per-render bars live in test_gpu_tc.py; what they cannot show is that TRAINING in that mode is equivalent.  Here the
same seeded scene is trained for 300 iterations through the public Trainer in the fp32 mode (1e-5 parity with the
reference, op by op) and in the bf16 mode, with identical initial parameters, ray batches and random draws, and — for
a short prefix, where two fp32 implementations have not yet diverged chaotically — by the oracle's training step
(oracle/train_step.py: the reference's op sequence in eager torch) on the same device."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_RAYS, STEPS, LOG2T = 4096, 300, 17


def _scene():
    from indoor_nerf_b200 import synthetic
    return synthetic.blender_scene(200, 200, n_views=20)


def _train(mode, steps, scene):
    import indoor_nerf_b200 as pn
    from indoor_nerf_b200 import model as pmodel, synthetic
    from indoor_nerf_b200.trainer import Trainer
    dev = torch.device("cuda", 0)
    a = pmodel.default_args(bounding_box=scene["bounding_box"], lrate=0.01, log2_hashmap_size=LOG2T)
    torch.manual_seed(0)
    kw, _, _, _, opt = pmodel.create_nerf(a, device=dev)
    tr = Trainer(a, kw, opt, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"])
    init = ([e.weight.detach().clone() for e in kw["embed_fn"].embeddings],
            [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in (kw["network_fn"], kw["network_fine"])])
    pn.set_mlp_mode(mode)
    losses, psnrs = [], []
    try:
        torch.manual_seed(1)
        for i in range(steps):
            r, t = synthetic.ray_batch(scene, N_RAYS, seed=i, device=dev)
            loss, psnr = tr.step(r, t)
            losses.append(loss)
            psnrs.append(psnr)
    finally:
        pn.set_mlp_mode("fp32")
    return torch.stack(losses).cpu().numpy(), torch.stack(psnrs).cpu().numpy(), init


def _train_oracle(steps, scene, init):
    from oracle.train_step import OracleModel, train_step
    from indoor_nerf_b200 import synthetic
    dev = torch.device("cuda", 0)
    om = OracleModel(*scene["bounding_box"], log2T=LOG2T, finest=512, device=dev, lr=0.01)
    tables, nets = init
    with torch.no_grad():
        for l in range(16):
            om.tables[l].copy_(tables[l])
        for n, sd in zip(om.nets, nets):
            n["s0"].copy_(sd["sigma_net.0.weight"]); n["s1"].copy_(sd["sigma_net.1.weight"])
            n["c0"].copy_(sd["color_net.0.weight"]); n["c1"].copy_(sd["color_net.1.weight"]); n["c2"].copy_(sd["color_net.2.weight"])
    losses = []
    torch.manual_seed(1)
    for i in range(steps):
        r, t = synthetic.ray_batch(scene, N_RAYS, seed=i, device=dev)
        losses.append(train_step(om, r, t, near=scene["near"], far=scene["far"], chunk=N_RAYS))
    return np.array(losses)


def test_training_trajectories_fp32_bf16_oracle():
    scene = _scene()
    l32, p32, init = _train("fp32", STEPS, scene)
    l16, p16, init16 = _train("bf16", STEPS, scene)
    assert all(torch.equal(a, b) for a, b in zip(init[0], init16[0]))          # same initial tables in both runs
    lo = _train_oracle(12, scene, init)
    assert np.isfinite(l32).all() and np.isfinite(l16).all() and np.isfinite(lo).all()

    # (1) fp32 mode vs the oracle: same parameters, same draws (t_rand, u, TV cube origins in the same RNG order) ->
    #     the loss of every early iteration agrees to fp32-accumulation level before the trajectories decorrelate
    rel = np.abs(l32[:12] - lo) / lo
    print("fp32 mode vs oracle, first 12 losses: rel diff", np.round(rel, 7))
    assert rel[0] < 2e-5 and rel[:4].max() < 2e-4 and rel.max() < 5e-3, rel

    # (2) bf16 mode vs fp32 mode: first step within the op-level bar, then the same learning curve
    assert abs(l16[0] - l32[0]) / l32[0] < 2e-3
    sm = lambda x, a, b: float(np.mean(x[a:b]))
    rows = []
    for a, b in ((20, 40), (40, 80), (80, 150), (150, 220), (220, 300)):
        rows.append((a, b, sm(p32, a, b), sm(p16, a, b), sm(l32, a, b), sm(l16, a, b)))
    print("window  psnr fp32  psnr bf16 | loss fp32  loss bf16")
    for r in rows:
        print("%3d-%3d  %8.3f  %8.3f | %.5f  %.5f" % r)
    for a, b, q32, q16, m32, m16 in rows:
        assert abs(q16 - q32) < 0.35, ("PSNR windows differ", a, b, q32, q16)          # dB, mean over the window
        assert abs(m16 - m32) / m32 < 0.08, ("loss windows differ", a, b, m32, m16)
    # both actually learn: > 6 dB over the first-iteration PSNR by the end
    assert sm(p32, 250, 300) - p32[0] > 6.0 and sm(p16, 250, 300) - p16[0] > 6.0, (p32[0], sm(p32, 250, 300), sm(p16, 250, 300))
    # and the bf16 run is not systematically worse: final-window PSNR within 0.25 dB
    assert sm(p16, 250, 300) > sm(p32, 250, 300) - 0.25


def test_resume_continues_learning_rate_and_tv_cutoff():
    """create_nerf's `start` -> Trainer(start=...): after loading a checkpoint written at global_step 1500 the learning
    rate is the reference's lrate * 0.1 ** (1500 / decay_steps) (run_nerf.py:1289-1293), not lrate, and the TV term is
    off for good after the first resumed iteration (`i > 1000`, :1036-1037)."""
    import os
    import tempfile
    from indoor_nerf_b200 import model as pmodel, synthetic
    from indoor_nerf_b200.trainer import Trainer
    dev = torch.device("cuda", 0)
    scene = _scene()
    with tempfile.TemporaryDirectory() as d:
        a = pmodel.default_args(bounding_box=scene["bounding_box"], lrate=5e-4, lrate_decay=10, log2_hashmap_size=14,
                                basedir=d, expname="run", no_reload=False)
        os.makedirs(os.path.join(d, "run"))
        kw, _, start, _, opt = pmodel.create_nerf(a, device=dev)
        assert start == 0
        tr = Trainer(a, kw, opt, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"], start=start)
        r, t = synthetic.ray_batch(scene, 512, seed=0, device=dev)
        tr.step(r, t)
        pmodel.save_checkpoint(os.path.join(d, "run", "001500.tar"), 1500, kw, opt)
        kw2, _, start2, _, opt2 = pmodel.create_nerf(a, device=dev)
        assert start2 == 1500
        assert "buffer" in opt2.state_dict()["param_groups"][0]                       # the reference's RAdam reads it
        tr2 = Trainer(a, kw2, opt2, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"], start=start2)
        want = 5e-4 * 0.1 ** (1500 / 10000)
        assert all(abs(g["lr"] - want) < 1e-12 for g in opt2.param_groups)
        tr2.step(r, t)
        assert tr2.step_idx == 1501 and tr2.tv_weight == 0.0
        assert all(abs(g["lr"] - want) < 1e-12 for g in opt2.param_groups)            # lr(1500) again: global_step repeats on resume
        assert torch.equal(kw2["embed_fn"].embeddings[3].weight.detach() != kw["embed_fn"].embeddings[3].weight.detach(),
                           kw2["embed_fn"].embeddings[3].weight.detach() != kw["embed_fn"].embeddings[3].weight.detach())
