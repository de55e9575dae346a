#!/usr/bin/env python
"""bench.py -- SRGAN train images/sec (full G+E+D step) on N B200s, with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                      (the reference algorithm on the host CPU cores)

A "step" is one `sg.train(x, label)` of the notebook-03 recipe (SRGAN, proposed losses, unrolled k=5:
5 discriminator updates + generator/encoder phase 1 + generator phase 2, optimizer steps included) on a
synthetic CelebA-shaped batch (U(-1,1) 3x128x128 images, 4 domains), random-init weights.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "style-restricted_gan_b200")
for p in (os.path.join(PKG, "pyfiles"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

GF_PER_IMG = {"srgan_nb03": 433.09, "srgan_nb05": 433.09, "nb02_solo": 423.22}      # conv+linear GFLOP / image / step
WORKLOADS = {
    "srgan_nb03": dict(kind="srgan", nch=64, dis_nch=64, enc_nch=64, res_num=6, k=5, feature="mu",
                       desc="SRGAN nb03 recipe: G(3,64,2,2,6)+Encoder+solo-multi D, proposed losses "
                            "(class1 cycle5 idt5 reg.5 idt_reg.5 bKL10 corr100 hist100), unrolled k=5"),
    "srgan_nb05": dict(kind="srgan", nch=64, dis_nch=64, enc_nch=64, res_num=6, k=5, feature="mu", frozen=True,
                       desc="SRGAN nb05 recipe: nb03 with the encoder trunk frozen while optE is built "
                            "(Adam lr 1e-3 over fcmean / fcvar only); the trunk comes out of the notebook-04 "
                            "classifier job (Classifier_training, --pretrain-iters)"),
    "nb02_solo": dict(kind="single_solo", nch=64, dis_nch=64, enc_nch=64, res_num=6, k=5, feature="mu",
                      desc="SingleGAN nb02 recipe: Encoder_original + solo-multi D, proposed losses, k=5"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm=d["hbm_gbs"], bf16=d["bf16_tflops"], bf16_sus=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                self.rows.append([c.strip() for c in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def build_case(name, batch):
    import cases
    w = WORKLOADS[name]
    lbd = dict(cases.PROPOSED, **{"class": 1})
    return dict(kind=w["kind"], nch=w["nch"], dis_nch=w["dis_nch"], enc_nch=w["enc_nch"], res_num=w["res_num"],
                batch=batch, k=w["k"], lbd=lbd, feature=w["feature"], seed=0, frozen=w.get("frozen", False))


def synthetic(batch, seed, get_target):
    import cases
    return cases.synthetic_batch(batch, get_target, seed=seed)


# ------------------------------------------------------------------------------------------- CPU arm
def time_oracle(case, steps, warmup, threads, budget_s=None):
    """Reference algorithm (CPU oracle port) on the host cores: images/s over up to `steps` train() calls; stops
    early once `budget_s` seconds of timed work have passed.  Returns (images/s, s/step, steps timed)."""
    import cases
    import srgan_oracle as so
    torch.set_num_threads(threads)
    model, util, _ = cases.use_product_modules()       # module classes are only used to draw the default init
    torch.manual_seed(0)
    np.random.seed(0)
    nets = cases.build_nets(model, case)
    tr = cases.build_oracle(case, cases.state_dicts(nets), so)
    x, label = synthetic(case["batch"], 123, util.get_target)
    for _ in range(warmup):
        tr.train(x, label)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        tr.train(x, label)
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return case["batch"] * done / dt, dt / done, done


def run_reference(args):
    """The reference algorithm on the host CPU (all host threads), same recipe and the same PER-GPU batch as the GPU
    arm (64): at N = 1 this is the GPU arm's exact configuration; at N > 1 one step is a bounded sample (one rank's
    share, 64 images) of the global batch.  K timed steps unless the time budget (--ref-budget seconds, default 1200)
    runs out first; the number of steps actually timed is reported."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    batch = args.cpu_batch if args.cpu_batch else args.batch
    case = build_case(args.workload, batch)
    warm = min(args.warmup, 1)
    ips, spt, steps = time_oracle(case, max(1, args.steps), warm, threads, budget_s=args.ref_budget)
    sample = "%d step(s) of %s at batch %d on %d host threads (oracle port of the reference; the reference itself is " \
             "pure PyTorch, imported from /root/reference to pin the port, and cannot travel to the GPU box)" % (
                 steps, args.workload, batch, threads)
    line = {"impl": "reference", "metric": "srgan_train_images_per_sec", "value": ips, "unit": "images/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": spt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "per_gpu_batch": batch,
                       "global_batch": batch * args.gpus, "image": "3x128x128", "domains": 4,
                       "parallelism": "dp%d" % args.gpus, "steps_requested": args.steps,
                       "time_budget_s": args.ref_budget},
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
def time_dominant_kernel(batch, dev):
    """CUDA-event time of the dominant kernel alone: the residual-block convolution forward
    (3x3, 256->256 channels, 32x32, batch `batch`) through the C ABI, L2 flushed between launches."""
    import srgan_ops as ops
    bf16 = ops.get_conv_engine() == "bf16"
    x = torch.randn(batch, 256, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    if bf16:
        x = x.to(torch.bfloat16)          # the trunk's storage type (channels-last is preserved)
    w = (torch.randn(256, 256, 3, 3, device=dev) * 0.02).contiguous(memory_format=torch.channels_last)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.conv2d(x, w, None, 1, 1)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d(x, w, None, 1, 1)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.mean(ts))
    flops = 2.0 * batch * 1024 * 256 * 2304
    d = ops._desc(batch, 32, 32, 256, 256, 3, 3, 1, 1)
    engine = ops._lib().srgan_conv2d_engine(d, 0)
    return ms, flops, ("tcgen05_bf16" if bf16 else ("tcgen05_tf32" if engine == 2 else "ffma_fp32"))


def time_norm_kernel(batch, dev):
    """CUDA-event time of the fused instance-norm forward on [batch, 256, 32, 32] as the training step runs it, L2
    flushed between launches.  Returns (ms as shipped, ms stand-alone, description).
    bf16 trunk: the convolution that produces the tensor leaves per-tile sums (conv_umma STATS), so the norm is ONE
    pass (fold of the tile rows + apply kernel: read x, write y); stand-alone (and on fp32 storage) it is the
    statistics kernel + the apply kernel."""
    import srgan_ops as ops
    bf16 = ops.get_conv_engine() == "bf16"
    x = torch.randn(batch, 256, 32, 32, device=dev).contiguous(memory_format=torch.channels_last)
    g, b = torch.ones(256, device=dev), torch.zeros(256, device=dev)
    cb = torch.randn(batch, 256, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tiles = None
    if bf16:
        w = (torch.randn(256, 256, 3, 3, device=dev) * 0.02).contiguous(memory_format=torch.channels_last)
        x = ops.conv2d(x.to(torch.bfloat16), w, None, 1, 1)          # a real trunk tensor (+ tile statistics if enabled)
        tiles, ops._tile_stats = ops._tile_stats, None

    def run(fused):
        if fused:
            ops._tile_stats = tiles
        ops.instance_norm_act(x, g, b, cb, None, 1e-5, ops.ACT_RELU, 0.0)

    def timed(fused):
        for _ in range(3):
            run(fused)
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run(fused)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.mean(ts))
    alone = timed(False)
    if tiles is None:
        return alone, alone, "statistics kernel + apply kernel"
    return timed(True), alone, "tile statistics from the producing convolution's epilogue: fold + apply kernel"


def measure_tf32_peak(dev):
    a = torch.randn(8192, 8192, device=dev)
    b = torch.randn(8192, 8192, device=dev)
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for _ in range(2):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    return 2 * 8192 ** 3 / (best * 1e-3) / 1e12


def time_latent_losses(batch, dev):
    """The latent-loss pair (batch-KL + correlation + soft histogram, forward and backward: ONE kernel each) on a
    [batch, 8] latent batch: CUDA-event microseconds per launch (latency bound: batch x 8 floats)."""
    import srgan_ops as ops
    import util
    hi = util.histogram_imitation(dev)
    g = hi.gausshist
    mu = torch.randn(batch, 8, device=dev, requires_grad=True)
    kw = dict(n_cfg=batch, target=hi.target, bins=g.bins, hmin=g.min, hmax=g.max, sigma=g.sigma,
              flags=ops.LAT_BKL | ops.LAT_CORR | ops.LAT_HIST)
    w4 = torch.tensor([10.0, 100.0, 100.0, 0.0], device=dev)

    def ev():
        return torch.cuda.Event(enable_timing=True)
    fwd, bwd = [], []
    for it in range(13):
        e0, e1, e2 = ev(), ev(), ev()
        c0 = ops.abi_calls
        e0.record()
        losses, _ = ops.latent_losses(mu, None, **kw)
        e1.record()
        c1 = ops.abi_calls
        gmu, = torch.autograd.grad(losses, mu, w4)
        e2.record()
        c2 = ops.abi_calls
        torch.cuda.synchronize()
        if it >= 3:
            fwd.append(e0.elapsed_time(e1) * 1e3)
            bwd.append(e1.elapsed_time(e2) * 1e3)
    return {"fwd_us": float(np.median(fwd)), "bwd_us": float(np.median(bwd)), "fwd_launches": c1 - c0,
            "bwd_launches": c2 - c1, "note": "event-to-event times of eager calls (include the Python dispatch of "
            "the autograd node); inside the replayed graph the two kernels run back to back"}


def _fresh_trainer(workload, batch_cfg, dev, ops, cases):
    """Nets + trainer of `workload` with the fixed parity seeds (same recipe as __graft_entry__.smoke())."""
    case = build_case(workload, batch_cfg)
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    np.random.seed(0)
    nets = cases.build_nets(model, case, dev)
    sds = cases.state_dicts(nets)
    torch.manual_seed(1)
    sg = cases.build_trainer(nb, case, tuple(n.to(dev) for n in nets), dev, adam=ops.FusedAdam)
    return case, sds, sg, util


def parity_check(workload, dev, tol):
    """Before anything is timed: ONE eager training step of the benchmarked recipe (full-width nets, k = 5) at batch 8
    on the GPU against the CPU oracle on the same seeded weights, batch and host-drawn noise.  The three reported
    losses must agree within `tol` (relative to max(1, |ref|)); raises otherwise."""
    import cases
    import srgan_ops as ops
    import srgan_oracle as so
    case, sds, sg, util = _fresh_trainer(workload, 8, dev, ops, cases)
    torch.manual_seed(1)
    oracle = cases.build_oracle(case, sds, so)
    x, label = cases.synthetic_batch(8, util.get_target)
    torch.manual_seed(2)
    t0 = time.perf_counter()
    ref = [float(e) for e in oracle.train(x, label)]
    cpu_s = time.perf_counter() - t0
    torch.manual_seed(2)
    got = [float(e) for e in sg.train(x.to(dev), {"source": label["source"].to(dev), "target": label["target"]})]
    torch.cuda.synchronize()
    err = [abs(g - r) / max(1.0, abs(r)) for g, r in zip(got, ref)]
    out = {"what": "one eager step of the same recipe at batch 8 vs the CPU oracle (errG, errD, errE)",
           "gpu": got, "oracle": ref, "rel_err": err, "tol": tol, "ok": bool(max(err) <= tol),
           "oracle_step_s": cpu_s}
    if not out["ok"]:
        raise SystemExit("bench.py parity check FAILED: %s" % json.dumps(out))
    return out


def dp_invariance_check(workload, dev, rank, world, tol=None):
    """N > 1: one eager data-parallel step at 8 images per GPU.  (1) the reported losses and the latent statistics
    blob (mean / variance / correlation / histogram bins of the GLOBAL batch) are bit-identical on every rank;
    (2) rank 0 repeats the step as a single-GPU job on the same global batch of 8 N images (same seeds, same host
    noise) and the losses agree to `tol` relative (summation order of the mean over ranks vs over the batch)."""
    import torch.distributed as dist
    import cases
    import srgan_ops as ops
    B = 8
    if tol is None:
        # the N-rank and the single-GPU run differ in the summation order of the weight gradients (pixel splits of
        # wgrad follow the per-rank batch) and of the mean over ranks; TF32 / fp32 storage keeps the losses of the
        # step within 5e-6 (measured 3e-7), bf16 storage turns the same differences into 2^-9 steps of stored
        # activations (measured 1.2e-5)
        tol = 2e-4 if ops.get_conv_engine() == "bf16" else 5e-6
    case, _, sg, util = _fresh_trainer(workload, B * world, dev, ops, cases)
    xg, lab = cases.synthetic_batch(B * world, util.get_target)
    sl = slice(rank * B, (rank + 1) * B)
    torch.manual_seed(2)
    errs = sg.train(xg[sl].to(dev), {"source": lab["source"][sl].to(dev), "target": lab["target"][sl]})
    mine = torch.cat([torch.stack([e.detach().float().reshape(()) for e in errs]), sg.latent_stats.detach().float()])
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    torch.cuda.synchronize()
    same = all(torch.equal(allv[0].view(torch.int32), v.view(torch.int32)) for v in allv)
    out = {"what": "one eager step at 8 images/GPU: ranks bit-identical; rank 0 re-runs the global batch single-GPU",
           "ranks_bit_identical": bool(same), "blob_floats": int(mine.numel() - 3), "tol": tol}
    del sg
    if rank == 0:
        with ops.single_process():
            _, _, sg1, _ = _fresh_trainer(workload, B * world, dev, ops, cases)
            torch.manual_seed(2)
            e1 = sg1.train(xg.to(dev), {"source": lab["source"].to(dev), "target": lab["target"]})
            one = [float(e) for e in e1]
            blob1 = sg1.latent_stats.detach().float()
        dp = [float(v) for v in allv[0][:3]]
        rel = [abs(a - b) / max(1.0, abs(b)) for a, b in zip(dp, one)]
        blob_rel = float((allv[0][3:] - blob1).norm() / blob1.norm().clamp_min(1e-30))
        out.update({"dp_losses": dp, "single_gpu_losses": one, "rel_err": rel, "latent_stats_rel_l2": blob_rel,
                    "ok": bool(same and max(rel) <= tol and blob_rel <= tol)})
        del sg1
    flag = torch.tensor([1 if (rank != 0 or out.get("ok")) else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag) != 1 or not same:
        raise SystemExit("bench.py data-parallel invariance check FAILED: %s" % json.dumps(out))
    return out


def run_ours(args):
    import torch.distributed as dist
    import cases
    import srgan_ops as ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = "cuda:%d" % local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    if args.engine:
        ops.set_conv_engine(args.engine)

    checks = {}
    if not args.skip_checks:
        tol = 5e-3 if ops.get_conv_engine() != "bf16" else 2e-2
        if world == 1:
            checks["parity_check"] = parity_check(args.workload, dev, tol)
        else:
            checks["dp_check"] = dp_invariance_check(args.workload, dev, rank, world)
        import gc
        gc.collect()
        torch.cuda.empty_cache()

    batch = args.batch                                   # per-GPU batch: weak scaling
    case = build_case(args.workload, batch * world)      # n_batch (batch-KL) = configured GLOBAL batch
    model, util, nb = cases.use_product_modules()
    torch.manual_seed(0)
    np.random.seed(0)
    nets = cases.build_nets(model, case, dev)
    G, D, E = nets
    G, D, E = G.to(dev), D.to(dev), E.to(dev)
    # every rank gets its own slice of the synthetic global batch
    xg, lab = synthetic(batch * world, 123, util.get_target)
    sl = slice(rank * batch, (rank + 1) * batch)
    x_host = xg[sl].contiguous().pin_memory()
    src_host = lab["source"][sl].contiguous().pin_memory()
    tgt = lab["target"][sl].contiguous()
    x_dev = x_host.to(dev)
    label_dev = {"source": src_host.to(dev), "target": tgt}
    pretrained = None
    if case.get("frozen") and args.pretrain_iters > 0:
        # notebook 04 -> 05: the encoder trunk handed to freeze_melt comes out of the classifier job (synthetic labels
        # here), data parallel like the step (gradients all-reduced, identical weights on every rank)
        cls = model.Encoder_classifier(3, 8, case["enc_nch"], 4, "instance", 4).to(dev)
        cls.load_state_dict({k: v for k, v in E.state_dict().items() if not k.startswith(("fcmean", "fcvar"))})
        job = nb.Classifier_training(cls, lr=1e-4)
        for _ in range(args.pretrain_iters):
            ploss, pacc = job.train_step(x_dev, label_dev["source"])
        E.load_state_dict(cls.state_dict(), strict=False)
        pretrained = "Encoder_classifier via Classifier_training (notebook 04 job), %d iterations on the synthetic " \
                     "batch: loss %.4f, accuracy %.3f" % (args.pretrain_iters, float(ploss), float(pacc))
        del cls, job
    sg = cases.build_trainer(nb, case, (G, D, E), dev, adam=ops.FusedAdam)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, warmup, sample_clocks):
        for _ in range(warmup):
            step_fn()
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        calls0 = ops.abi_calls
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step_fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if sampler else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms, ops.abi_calls - calls0, clocks

    # one eager step first: counts the kernels of a step (a replayed CUDA graph launches them without passing
    # through the Python-side counter) and sizes every lazily grown buffer
    calls0 = ops.abi_calls
    sg.train(x_dev, label_dev)
    launches_per_step = ops.abi_calls - calls0
    use_graph = args.graph in ("on", "auto")
    if use_graph:
        sg.enable_cuda_graph(warmup=1)

    def step_resident():
        sg.train(x_dev, label_dev)

    sink = torch.empty(3, dtype=torch.float32).pin_memory()

    def step_e2e():
        xd = x_host.to(dev, non_blocking=True)
        ld = {"source": src_host.to(dev, non_blocking=True), "target": tgt}
        errs = sg.train(xd, ld)
        sink.copy_(torch.stack([e.detach().float() for e in errs]), non_blocking=False)   # device -> host read

    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, True)
    launches = launches_per_step * args.steps
    value = batch * world * args.steps / (ms * 1e-3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, 1, False)
    e2e = batch * world * args.steps / (ms_e2e * 1e-3)

    line = None
    if rank == 0:
        pk = peaks()
        kms, kflops, kname = time_dominant_kernel(batch, dev)
        cublas_tf32 = measure_tf32_peak(dev)
        achieved = kflops / (kms * 1e-3) / 1e12
        # Denominator: MEASURED_PEAKS.json (driver-written).  It holds the dense bf16 rate; tcgen05 kind::tf32 issues
        # at half the bf16 rate, so a TF32 kernel is measured against half of it.  (The kernel is timed alone, after a
        # cold L2 flush: the burst figure applies.)
        is_tf32 = kname == "tcgen05_tf32"
        peak = 0.5 * pk["bf16"] if is_tf32 else pk["bf16"]
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            # dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed `ncu --set full` capture
            # of the SHIPPED kernel (regenerated every round, see profiles/README.md), per launch at batch 64;
            # the kernel's traffic is linear in the batch
            t = json.load(open(tpath))
            key = "res_conv_fprop_dram_bytes_b64" + ("" if is_tf32 else "_bf16")
            if key in t:
                traffic = t[key] * batch / 64.0
        roof = {"bound": "tensor", "kernel": "conv2d fprop 3x3 256->256 @32x32 (residual block), " + kname,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": "%s x MEASURED_PEAKS.json bf16_tflops (%.1f TF/s, %s)" % (
                    "0.5" if is_tf32 else "1.0", pk["bf16"], pk["src"]),
                "algorithmic_flop": kflops,
                "algorithmic_bytes": (2 * batch * 1024 * 256 + 256 * 2304) * (4 if is_tf32 else 2),
                "note_cublas_tf32_tflops_this_run": cublas_tf32,
                "kernel_ms": kms, "step_gflop_per_image": GF_PER_IMG[args.workload],
                "step_gflop_per_image_executed": GF_PER_IMG[args.workload] - 16.6,   # skipped encoder work
                "step_tflops": value * GF_PER_IMG[args.workload] / 1e3}
        # secondary roofline: the fused instance-norm (+ conditional bias + affine + ReLU) forward of the residual
        # blocks, HBM bound; algorithmic bytes = read x + write y
        nms, nms_alone, nhow = time_norm_kernel(batch, dev)
        nbytes = 2.0 * (4 if is_tf32 or kname == "ffma_fp32" else 2) * batch * 256 * 1024
        glue = {"bound": "hbm", "kernel": "instance norm + cond. bias + affine + ReLU forward, 256 ch @32x32 (" + nhow + ")",
                "achieved": nbytes / (nms * 1e-3) / 1e9, "peak": pk["hbm"], "unit": "GB/s",
                "frac": nbytes / (nms * 1e-3) / 1e9 / pk["hbm"], "kernel_ms": nms,
                "algorithmic_bytes": nbytes, "standalone_ms": nms_alone,
                "standalone_frac": nbytes / (nms_alone * 1e-3) / 1e9 / pk["hbm"],
                "note": "algorithmic bytes = read x + write y (SURVEY 8d); peak = MEASURED_PEAKS.json hbm_gbs, a copy "
                        "of 2 GB - a plain device copy of a plane of this size reaches ~4.7 TB/s "
                        "(tools/stream_probe.cu)"}
        cpu = None
        if world == 1 and not args.no_cpu:
            threads = os.cpu_count() or 1
            cb = args.cpu_batch or 16
            ips, spt, _ = time_oracle(build_case(args.workload, cb), 1, 0, threads)
            cpu = {"value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                   "sample": "1 step of the same recipe at batch %d (%.1f s) on %d host threads; the --impl reference "
                             "arm runs the full batch" % (cb, spt, threads)}
        h2d = x_host.numel() * 4 + src_host.numel() * 8 + 3 * (batch * 8 * 4)
        line = {"metric": "srgan_train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {"tcgen05_tf32": "tf32", "tcgen05_bf16": "bf16"}.get(kname, "f32"), "data": "synthetic",
                "config": {"workload": WORKLOADS[args.workload]["desc"], "per_gpu_batch": batch,
                           "global_batch": batch * world, "image": "3x128x128", "domains": 4,
                           "parallelism": "dp%d" % world, "conv_engine": ops.get_conv_engine(),
                           "cuda_graph": use_graph,
                           "l2": "per-step working set (GBs of activations) >> 126 MB L2; no explicit flush"},
                "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches,
                "gpu_launches_note": "C-ABI kernel entry points of one eager step (%d) x steps; the timed steps replay "
                                     "the same kernels from one CUDA graph" % launches_per_step,
                "clocks": clocks, "roofline": roof, "roofline_glue": glue,
                "latent_loss_pair": time_latent_losses(batch * world, dev),
                "cpu_baseline": cpu}
        if pretrained:
            line["config"]["encoder_pretraining"] = pretrained
        line.update(checks)
        print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        if use_graph:
            # a CUDA graph that holds captured NCCL kernels must not outlive its communicator, and tearing the two
            # down in the interpreter's exit order blocks: drop the graph first, then leave without finalisers
            sg.disable_cuda_graph()
            import gc
            gc.collect()
            sys.stderr.flush()
            os._exit(0)
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="srgan_nb03", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=64, help="per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=0,
                    help="batch of the CPU sample (default: 16 for the cpu_baseline leg, --batch for --impl reference)")
    ap.add_argument("--ref-budget", type=float, default=600.0,
                    help="--impl reference: stop after this many seconds of timed steps")
    ap.add_argument("--skip-checks", action="store_true",
                    help="skip the pre-timing parity check (N=1: vs the CPU oracle; N>1: rank / single-GPU invariance)")
    ap.add_argument("--engine", default="bf16", choices=["auto", "fp32", "tf32", "bf16"],
                    help="bf16 (default): bf16 storage in the generator and encoder trunks (tcgen05 kind::f16, stated "
                         "tolerance: losses 1e-2); auto: TF32 on fp32 storage; fp32: exact FFMA engine")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--pretrain-iters", type=int, default=4,
                    help="srgan_nb05: iterations of the notebook-04 classifier job that produce the frozen encoder trunk")
    ap.add_argument("--graph", default="auto", choices=["auto", "on", "off"],
                    help="replay the step as one CUDA graph (sg.enable_cuda_graph) or issue every kernel from Python; "
                         "auto = on")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
        run_ours(args)


if __name__ == "__main__":
    main()
