#!/usr/bin/env python
"""bench.py — headline benchmark of the HashNeRF hot path on B200 (contract: see DESIGN.md §Measurement).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours       one "step" = one training iteration of BASELINE config 3 on every rank: 65536 rays/rank,
           64+128 samples, 16 levels, T=2^19, finest 512 -> render fwd (coarse+fine), image+sparsity+TV
           losses, backward, gradient all-reduce (N>1), RAdam.  `value` = rays/s with the ray batch already
           in HBM; `e2e` = the same through the public API with pinned-host batches (H2D inside the timed
           region, loss read back every step).  Also reported: the hash-encode kernel's achieved GB/s vs the
           measured HBM peak (roofline), full-frame render Mpix/s (config 2), the oracle on the host cores
           (cpu_baseline) and the reference's eager torch path on this GPU.
reference  the reference's algorithm (oracle port: the Python reference cannot travel to the GPU box) on the
           host cores, same metric/unit, bounded sample of the workload per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MLP_MODE = os.environ.get("POCKETNERF_MLP", "bf16")     # "bf16": tcgen05 tensor-core NeRFSmall (north_star's 2e-3 mode); "fp32": FFMA
RAYS_PER_RANK = 65536
N_SAMPLES, N_IMPORTANCE = 64, 128
HASH_BYTES_PER_POINT = 12 + 16 * 8 * 2 * 4 + 16 * 2 * 4      # 1164 B/point


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        if self.index is None:                  # ranks other than 0 do not sample
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def __exit__(self, *a):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [r for t, r in self.rows if self.t0 is not None and self.t1 is not None and self.t0 <= t <= self.t1 + 0.15]
        window = "timed region"
        if not inside:      # region shorter than the sampling period: fall back to the samples taken under load (warm-up)
            inside, window = [r for t, r in self.rows], "warm-up + timed region"
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm),
                "window": window}


def dist_setup(n):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        import datetime
        # a short collective timeout: a rank-divergent code path must fail in minutes, not hold N GPUs for the
        # default 10 minutes
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))
        return dist.get_rank(), world, local
    return 0, 1, 0


def max_over_ranks(ms, world):
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def time_steps(fn, steps, world):
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    barrier(world)
    return max_over_ranks(e0.elapsed_time(e1), world)


def cpu_reference_steps(steps, warmup, rays_per_step, threads):
    """The oracle's train step on the host cores.  Returns (rays_per_s, ms_per_step)."""
    from indoor_nerf_b200 import synthetic
    from oracle.train_step import OracleModel, train_step
    torch.set_num_threads(threads)
    scene = synthetic.blender_scene(400, 400, n_views=100)
    model = OracleModel(*scene["bounding_box"], log2T=19, finest=512, device="cpu", lr=0.01)
    batches = [synthetic.ray_batch(scene, rays_per_step, seed=100 + i) for i in range(2)]
    for i in range(warmup):
        train_step(model, *batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        train_step(model, *batches[i % 2])
    dt = time.perf_counter() - t0
    return rays_per_step * steps / dt, dt / steps * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    total = args.steps + args.warmup
    rays = 1024 if total <= 16 else (256 if total <= 64 else 64)
    v, ms = cpu_reference_steps(args.steps, args.warmup, rays, threads)
    line = {
        "impl": "reference", "metric": "train_rays_per_s", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "cpu_baseline": {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                         "sample": "%d rays/step of the 65536-ray workload (64+128 samples, T=2^19), full train step "
                                   "(render fwd, img+sparsity+TV losses, backward, RAdam) by the oracle on the host" % rays},
        "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


def workload_config(n):
    return {"workload": "BASELINE configs[2]: chair-shaped synthetic Blender scene 400x400, HashNeRF training step, "
                        "N_rand=65536 rays per GPU, 64 coarse + 128 importance samples, 16 levels, log2_hashmap 19, "
                        "finest_res 512, NeRFSmall coarse+fine, white_bkgd, perturb=1",
            "rays_per_gpu": RAYS_PER_RANK, "global_rays_per_step": RAYS_PER_RANK * n,
            "points_per_ray": 2 * N_SAMPLES + N_IMPORTANCE, "parallelism": "dp%d" % n,
            "l2_policy": "inputs larger than L2: 16.8 M points/step (201 MB positions, 2.1 GB features) stream "
                         "through; the 64 MiB table set is L2-resident by design"}


def fp32_mode_leg(pn, step_resident):
    try:
        pn.set_mlp_mode("fp32")
        for i in range(2):
            step_resident(i)
        ms32 = time_steps(step_resident, 3, 1)
        return {"value": RAYS_PER_RANK * 3 / (ms32 / 1e3), "unit": "rays/s", "ms_per_step": ms32 / 3,
                "what": "same step with the fp32 FFMA NeRFSmall kernels and unfused hash kernels"}
    except Exception as ex:
        return {"error": repr(ex)}
    finally:
        pn.set_mlp_mode(MLP_MODE)


def hash_t22_leg(pn, ops, pts, dev, peak):
    """BASELINE configs[3] table size (log2_hashmap 22: 512 MiB of tables, 4x the L2): the hash-encode kernels where the
    table set cannot be L2-resident, same ray-ordered points."""
    try:
        box = (pts.min(0)[0].cpu() - 0.1, pts.max(0)[0].cpu() + 0.1)
        emb = pn.HashEmbedder(box, log2_hashmap_size=22, finest_resolution=512).to(dev)
        tables = [t.detach() for t in emb.tables()]
        P = pts.shape[0]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            for _ in range(2):
                ops.hash_encode_fwd(emb.grid(), tables, pts)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                ops.hash_encode_fwd(emb.grid(), tables, pts)
            e1.record()
            torch.cuda.synchronize()
            tf = e0.elapsed_time(e1) / 5
            dfeat = torch.randn(P, 32, device=dev)
            flat = torch.zeros(16, 1 << 22, 2, device=dev)
            ops.hash_encode_bwd(emb.grid(), list(flat.unbind(0)), pts, dfeat)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                ops.hash_encode_bwd(emb.grid(), list(flat.unbind(0)), pts, dfeat)
            e1.record()
            torch.cuda.synchronize()
            tb = e0.elapsed_time(e1) / 5
        ach = HASH_BYTES_PER_POINT * P / (tf * 1e-3) / 1e9
        return {"kernel": "hash_fwd_kernel", "log2_hashmap_size": 22, "table_bytes": 16 * (1 << 22) * 8,
                "points_per_launch": P, "ms_per_launch": tf, "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "bound": "hbm",
                "backward": {"kernel": "hash_bwd_kernel", "ms_per_launch": tb,
                             "achieved": HASH_BYTES_PER_POINT * P / (tb * 1e-3) / 1e9},
                "note": "algorithmic bytes (8 B per gathered corner); the table set is 4x the L2, so the fine levels "
                        "miss it and every 8-byte gather moves a 32-byte DRAM sector"}
    except Exception as ex:
        return {"error": repr(ex)}


def io_legs(pn, scene, kw2, scene2, dev):
    """on-GPU ray batching (rays/s of batch generation, bank bytes vs the reference's precomputed
    tensor) and test-set evaluation through render_path (frames to pinned host memory + on-device PSNR)."""
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        n_img = 100
        images = torch.rand(n_img, scene["H"], scene["W"], 3, device=dev)
        bank = pn.RayBank(scene["H"], scene["W"], scene["K"], scene["poses"], images, list(range(n_img)), device=dev)
        bank.shuffle()
        for _ in range(3):
            bank.next_batch(RAYS_PER_RANK)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            bank.next_batch(RAYS_PER_RANK)
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 20
        out["ray_bank"] = {"rays_per_s": RAYS_PER_RANK / (t * 1e-3), "ms_per_batch": t, "batch": RAYS_PER_RANK,
                           "bytes": bank.bytes_resident(),
                           "what": "use_batching batches generated from (image, pixel) ids, 100 views of 400x400"}
        del bank, images
    except Exception as ex:
        out["ray_bank"] = {"error": repr(ex)}
    try:
        poses = torch.from_numpy(scene2["poses"][:3])
        gts = torch.rand(3, 800, 800, 3).pin_memory()          # ground truth as a loader would hold it: pinned host memory
        kw = dict(kw2, near=2., far=6.)
        hwf = [800, 800, scene2["focal"]]
        pn.render_path(poses[:1], hwf, scene2["K"], 1 << 17, kw, gt_imgs=gts[:1])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        pn.render_path(poses, hwf, scene2["K"], 1 << 17, kw, gt_imgs=gts)
        dt = (time.perf_counter() - t0) / 3
        out["render_path"] = {"mpix_per_s": 0.64 / dt, "ms_per_frame": dt * 1e3, "d2h_bytes_per_frame": 800 * 800 * 16,
                              "what": "render_path over 3 test views of configs[1] incl. host copies of rgb/depth and "
                                      "on-device PSNR against host ground truth (wall clock, synchronised)"}
    except Exception as ex:
        out["render_path"] = {"error": repr(ex)}
    return out


def render_t22_leg(pn, pmodel, synthetic, dev):
    """Test-view rendering at the ScanNet-config table size (log2_hashmap 22, 512 MiB of fp32 tables = 4x the L2) with a
    quantised (A-CAQ, 8-bit) model: gathering fp32 entries, fake-quantising them in the gather (the reference's eval
    semantics), and gathering u8 codes (HashEmbedder.pack_for_inference: 128 MiB; same embeddings bit for bit in the fp32 mode,
    within fp32 rounding before the bf16 cast in this mode)."""
    try:
        sc = synthetic.blender_scene(800, 800, n_views=2)
        a = pmodel.default_args(bounding_box=sc["bounding_box"], finest_res=1024, log2_hashmap_size=22,
                                use_quantization=True, quantization_bits=8)
        _, kw, _, _, _ = pmodel.create_nerf(a, device=dev)
        emb, nets = kw["embed_fn"], [kw["network_fn"], kw["network_fine"]]
        for l, q in enumerate(emb.quantizers):
            q.calibrate(emb.embeddings[l].weight.detach())
        for n in nets:
            n.sigma_weight_quantizer.calibrate(n.sigma_net[0].weight.detach())
            n.sigma_act_quantizers[0].calibrate(torch.tensor([-0.5, 0.5], device=dev))   # non-degenerate zero point
            n.eval()
        emb.eval()
        c2w = torch.from_numpy(sc["poses"][0][:3, :4])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def mpix():
            def frame():
                with torch.no_grad():
                    return pn.render(800, 800, sc["K"], chunk=1 << 17, c2w=c2w, near=2., far=6., **kw)[0]
            frame(); frame()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                img = frame()
            e1.record()
            torch.cuda.synchronize()
            return 0.64 / (e0.elapsed_time(e1) / 3 * 1e-3), img

        out = {"workload": "800x800 test view, log2_hashmap 22, finest_res 1024, 64+128 samples, 8-bit A-CAQ model"}
        emb.use_quantization = False
        out["fp32_tables_mpix_per_s"], _ = mpix()
        emb.use_quantization = True
        out["fake_quant_in_gather_mpix_per_s"], img_fq = mpix()
        packed = emb.pack_for_inference()
        out["u8_codes_mpix_per_s"], img_pk = mpix()
        out["max_abs_frame_diff_codes_vs_fake_quant"] = float((img_fq - img_pk).abs().max())
        out["table_bytes"] = {"fp32": 16 * (1 << 22) * 8, "codes": packed.nbytes()}
        return out
    except Exception as ex:
        return {"error": repr(ex)}


def run_ours(args):
    import indoor_nerf_b200 as pn
    from indoor_nerf_b200 import _lib, model as pmodel, ops, synthetic
    from indoor_nerf_b200.trainer import Trainer

    pn.set_mlp_mode(MLP_MODE)
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(1234 + rank)
    scene = synthetic.blender_scene(400, 400, n_views=100)
    a = pmodel.default_args(bounding_box=scene["bounding_box"], lrate=0.01)
    torch.manual_seed(0)                                   # identical initial parameters on every rank
    kw_train, kw_test, _, _, opt = pmodel.create_nerf(a, device=dev)
    torch.manual_seed(1234 + rank)
    tr = Trainer(a, kw_train, opt, scene["H"], scene["W"], scene["K"], scene["near"], scene["far"],
                 group=None if world == 1 else torch.distributed.group.WORLD)
    pool = [synthetic.ray_batch(scene, RAYS_PER_RANK, seed=1000 * rank + i, pin=True) for i in range(4)]
    pool_dev = [(r.to(dev), t.to(dev)) for r, t in pool]
    h2d = pool[0][0].numel() * 4 + pool[0][1].numel() * 4

    def step_resident(i):
        r, t = pool_dev[i % len(pool_dev)]
        tr.step(r, t)

    losses = []

    def step_e2e(i):
        r, t = pool[i % len(pool)]
        loss, _ = tr.step(r.to(dev, non_blocking=True), t.to(dev, non_blocking=True))
        losses.append(loss.item())

    with ClockSampler(local if rank == 0 else None) as clk:
        for i in range(args.warmup):
            step_e2e(i)
        torch.cuda.synchronize()
        clk.mark_start()
        l0 = _lib.launch_count()
        ms = time_steps(step_resident, args.steps, world)
        launches = _lib.launch_count() - l0
        ms_e2e = time_steps(step_e2e, args.steps, world)
        clk.mark_end()
    rays_total = RAYS_PER_RANK * world * args.steps
    value = rays_total / (ms / 1e3)
    e2e = rays_total / (ms_e2e / 1e3)

    line = {
        "metric": "train_rays_per_s", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "bf16 NeRFSmall on tcgen05 (fp32 accumulate) + f32 hash grid / compositing / sampling" if MLP_MODE == "bf16"
        else "f32", "data": "synthetic", "config": workload_config(world),
        "clocks": clk.summary(),
        "e2e": {"value": e2e, "unit": "rays/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "final_loss": losses[-1] if losses else None,
    }

    if rank == 0:
        # ---- roofline of the dominant kernel: hash-encode forward on the fine pass's point set ---------
        embed = kw_train["embed_fn"]
        P = RAYS_PER_RANK * (N_SAMPLES + N_IMPORTANCE)
        r, t = pool_dev[0]
        with torch.no_grad():
            z = torch.sort(2.0 + 4.0 * torch.rand(RAYS_PER_RANK, N_SAMPLES + N_IMPORTANCE, device=dev), -1)[0]
            pts = ops.make_points(r[0], r[1], z).reshape(-1, 3)
            tables = [tt.detach() for tt in embed.tables()]
            for _ in range(3):
                ops.hash_encode_fwd(embed.grid(), tables, pts)
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ops.hash_encode_fwd(embed.grid(), tables, pts)
            e1.record()
            torch.cuda.synchronize()
            t_fwd = e0.elapsed_time(e1) / reps
            dfeat = torch.randn(P, 32, device=dev)
            flat = torch.zeros(16, 1 << 19, 2, device=dev)
            for _ in range(2):
                ops.hash_encode_bwd(embed.grid(), list(flat.unbind(0)), pts, dfeat)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                ops.hash_encode_bwd(embed.grid(), list(flat.unbind(0)), pts, dfeat)
            e1.record()
            torch.cuda.synchronize()
            t_bwd = e0.elapsed_time(e1) / reps
            del dfeat, flat, z
        peak, how = peaks()
        ach = HASH_BYTES_PER_POINT * P / (t_fwd * 1e-3) / 1e9
        line["roofline"] = {"kernel": "hash_fwd_kernel", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s",
                            "frac": ach / peak, "traffic": None, "peak_source": how,
                            "algorithmic_bytes_per_point": HASH_BYTES_PER_POINT, "points_per_launch": P,
                            "ms_per_launch": t_fwd,
                            "backward": {"kernel": "hash_bwd_kernel", "ms_per_launch": t_bwd,
                                         "achieved": HASH_BYTES_PER_POINT * P / (t_bwd * 1e-3) / 1e9}}
        line["roofline"]["traffic"] = 1.890e9 if P == 12582912 else None   # ncu dram bytes read+write per launch (profiles/r01_ncu_full_hash_fwd_mlp_fwd_fp32.csv)
        line["roofline_t22"] = hash_t22_leg(pn, ops, pts, dev, peak)
        del pts
        torch.cuda.empty_cache()
        # ---- the fused field kernels of the bf16 mode (what the training step actually launches) ---------
        if MLP_MODE == "bf16":
            try:
                r, t = pool_dev[0]
                z = torch.sort(2.0 + 4.0 * torch.rand(RAYS_PER_RANK, N_SAMPLES + N_IMPORTANCE, device=dev), -1)[0]
                pts3 = ops.make_points(r[0], r[1], z)
                vd = r[1] / r[1].norm(dim=-1, keepdim=True)
                net = kw_train["network_fine"]
                sh = pn.SHEncoder()
                fw, bw = [], []
                for it in range(6):
                    torch.cuda.synchronize()
                    e0.record()
                    out = pn.run_network(pts3, vd, net, embed, sh)
                    e1.record()
                    torch.cuda.synchronize()
                    fw.append(e0.elapsed_time(e1))
                    dout = torch.randn_like(out)
                    torch.cuda.synchronize()
                    e0.record()
                    out.backward(dout)
                    e1.record()
                    torch.cuda.synchronize()
                    bw.append(e0.elapsed_time(e1))
                tf, tb = float(np.median(fw[2:])), float(np.median(bw[2:]))
                tf_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1400.0) \
                    if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 1400.0
                line["fused_kernels"] = {
                    "points_per_launch": P,
                    "field_fwd": {"kernel": "mlp_tc_fwd_kernel<3,SRC_HASH>", "ms": tf,
                                  "hbm_algorithmic_GBs": (12 + 1024 + 16 + 64 + 1) * P / (tf * 1e-3) / 1e9,
                                  "tensor_TFLOPs": 18688 * P / (tf * 1e-3) / 1e12, "tensor_frac": 18688 * P / (tf * 1e-3) / 1e12 / tf_peak},
                    "field_bwd": {"kernel": "mlp_tc_bwd_kernel<SRC_TILE> (+ torch memset/adds of the autograd node)", "ms": tb,
                                  "hbm_algorithmic_GBs": (12 + 64 + 16 + 1 + 1024) * P / (tb * 1e-3) / 1e9,
                                  "tensor_TFLOPs": 56064 * P / (tb * 1e-3) / 1e12, "tensor_frac": 56064 * P / (tb * 1e-3) / 1e12 / tf_peak},
                    "note": "latency/L2-atomic bound, not HBM or tensor bound: see DESIGN.md §4 and profiles/"}
                del pts3, z, out, dout
                for prm in list(embed.parameters()) + list(net.parameters()):
                    prm.grad = None
            except Exception as ex:
                line["fused_kernels"] = {"error": repr(ex)}
            # the same training step in the fp32 (FFMA, 1e-5 parity) mode.  Single-GPU runs only: Trainer.step
            # all-reduces, and this block runs on rank 0 alone.
            if world == 1:
                line["fp32_mode"] = fp32_mode_leg(pn, step_resident)
        # ---- config 2: full 800x800 test-view render, finest_res 1024 -----------------------------------
        try:
            scene2 = synthetic.blender_scene(800, 800, n_views=8)
            a2 = pmodel.default_args(bounding_box=scene2["bounding_box"], finest_res=1024)
            _, kw2, _, _, _ = pmodel.create_nerf(a2, device=dev)
            c2w = torch.from_numpy(scene2["poses"][0][:3, :4])
            def frame():
                with torch.no_grad():
                    return pn.render(800, 800, scene2["K"], chunk=1 << 17, c2w=c2w, near=2., far=6., **kw2)
            frame(); frame()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                frame()
            e1.record()
            torch.cuda.synchronize()
            t_frame = e0.elapsed_time(e1) / 3
            line["render"] = {"workload": "BASELINE configs[1]: 800x800 test view, finest_res 1024, 64+128 samples",
                              "mpix_per_s": 0.64 / (t_frame * 1e-3), "ms_per_frame": t_frame}
            line.update(io_legs(pn, scene, kw2, scene2, dev))
            del kw2
            torch.cuda.empty_cache()
            line["render_t22_quantised"] = render_t22_leg(pn, pmodel, synthetic, dev)
        except Exception as ex:                                     # keep the headline even if this leg fails
            line["render"] = {"error": repr(ex)}
        torch.cuda.empty_cache()
        if world == 1 and not args.no_baselines:
            # ---- reference eager path on this GPU (the oracle's ATen ops), for the >=50x target ------------
            try:
                from oracle.train_step import OracleModel, train_step
                om = OracleModel(*scene["bounding_box"], log2T=19, finest=512, device=dev, lr=0.01)
                r, t = pool_dev[0]
                train_step(om, r, t)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(2):
                    train_step(om, r, t)
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / 2
                line["reference_eager_b200"] = {"value": RAYS_PER_RANK / dt, "unit": "rays/s", "ms_per_step": dt * 1e3,
                                                "what": "oracle (the reference's eager ATen op sequence) on the same "
                                                        "B200, same 65536-ray step, chunk 32768"}
                del om
                torch.cuda.empty_cache()
            except Exception as ex:
                line["reference_eager_b200"] = {"error": repr(ex)}
            # ---- oracle on the host cores -----------------------------------------------------------------------
            threads = os.cpu_count() or 1
            v, msc = cpu_reference_steps(3, 1, 1024, threads)
            line["cpu_baseline"] = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port", "ms_per_step": msc,
                                    "sample": "1024 rays/step of the same workload (64+128 samples, T=2^19), full "
                                              "train step by the oracle, 1 warm-up + 3 timed steps"}
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return line if rank == 0 else None


class StdoutToStderr:
    """Everything written to fd 1 while active goes to stderr (NCCL prints its version banner on stdout), so
    that the ONE JSON line is the only thing this program ever writes to stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *a):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def emit(line):
    sys.stdout.flush()
    os.write(1, (json.dumps(line) + "\n").encode())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-baselines", action="store_true", help="skip the CPU / eager-GPU baseline legs")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    with StdoutToStderr() as guard:
        line = run_reference(args) if args.impl == "reference" else run_ours(args)
    if line is not None:
        emit(line)


if __name__ == "__main__":
    main()
