"""bench.py -- E2ENet hot-path benchmark on B200 (contract: see the task's Measurement section).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-graph] [--no-inference]

Workload at every N: BASELINE.json configs[1] -- E2ENet BTCV-shaped training: batch 2 per GPU
(weak scaling), 1x64x160x160 CT patches, 14 classes, DSFF density 0.2, SGD(nesterov) +
clip + Masking.step() every iteration.  One "step" = one full training iteration of one batch:
`value` replays it as ONE CUDA graph with the batch resident, `e2e` adds the pinned-host -> device copy
of the batch and the loss read-back, `roofline` comes from an eager pass with per-launch CUDA events.
The `inference` object is BASELINE.json configs[2] (sliding-window voxels/s, tiles sharded over the N
ranks).  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PATCH = (64, 160, 160)
BATCH = 2
NCLS = 14
IN_CH = 1
DENSITY = 0.2
# dense algorithmic FLOPs of the conv / tconv / 1x1 GEMMs, config 2, B=2
FWD_GFLOP_B2 = 4536.6
STEP_GFLOP_B2 = 13607.0


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no_samples"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max(int(s[1]) for s in self.samples if s[1].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons}


# ---------------------------------------------------------------------------------------- CPU baseline (oracle port)
def cpu_baseline_patches_per_s(budget_s=20.0, crop=(32, 96, 96)):
    """the oracle's fp32 torch-CPU restatement of the same training iteration (fwd + DS loss + bwd +
    clip + SGD + apply_mask), on a bounded sample: one 1x1x32x96x96 crop = 0.18 patch."""
    import numpy as np
    import torch
    from collections import OrderedDict
    from oracle import masking as omask
    from oracle import network as onet
    from e2enet_medical_b200.training import POOLS
    torch.set_num_threads(os.cpu_count() or 1)
    pools = POOLS["btcv"]
    shapes = onet.param_shapes(IN_CH, 48, NCLS, pools)
    params = onet.det_params(shapes, seed=0)
    plist = [v.requires_grad_(True) for v in params.values()]
    opt = torch.optim.SGD(plist, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    import random
    random.seed(0)
    masks = {k: torch.from_numpy(m) for k, m in omask.init_uniform(shapes, DENSITY).items()}
    with torch.no_grad():
        for k, m in masks.items():
            params[k].mul_(m)
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.rand(1, IN_CH, *crop).astype(np.float32))
    tg, sp = [], np.array(crop)
    for k in range(4):
        tg.append(torch.from_numpy(np.round(rs.rand(1, 1, *sp) * (NCLS - 1)).astype(np.float32)))
        sp = sp // np.array(pools[k])

    def one():
        opt.zero_grad()
        outs = onet.unetpp_forward(params, x, pools)
        l = onet.ds_loss(outs, tg)
        l.backward()
        torch.nn.utils.clip_grad_norm_(plist, 12)
        opt.step()
        with torch.no_grad():
            for k, m in masks.items():
                params[k].mul_(m)
                st = opt.state[params[k]]
                if 'momentum_buffer' in st:
                    st['momentum_buffer'].mul_(m)
        return float(l)

    one()                                             # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        one()
        n += 1
        if time.perf_counter() - t0 > budget_s or n >= 8:
            break
    dt = (time.perf_counter() - t0) / n
    frac = float(np.prod(crop)) / float(np.prod(PATCH))
    return frac / dt, dt, n


def cpu_baseline_config0(budget_s=12.0):
    """BASELINE.json configs[0], unscaled: E2ENet 3d_fullres, density 0.2, fwd + DS loss + bwd on ONE synthetic
    1x1x40x56x40 Hippocampus-shaped patch on the host cores (oracle port, fp32, all threads)."""
    import numpy as np
    import torch
    import random
    from oracle import masking as omask
    from oracle import network as onet
    from e2enet_medical_b200.training import POOLS
    torch.set_num_threads(os.cpu_count() or 1)
    pools, patch, ncls = POOLS["hippo"], (40, 56, 40), 3
    shapes = onet.param_shapes(1, 48, ncls, pools)
    params = onet.det_params(shapes, seed=0)
    random.seed(0)
    for k, m in omask.init_uniform(shapes, DENSITY).items():
        params[k].mul_(torch.from_numpy(m))
    plist = [v.requires_grad_(True) for v in params.values()]
    rs = np.random.RandomState(1)
    x = torch.from_numpy(rs.rand(1, 1, *patch).astype(np.float32))
    tg, sp = [], np.array(patch)
    for k in range(4):
        tg.append(torch.from_numpy(np.round(rs.rand(1, 1, *sp) * (ncls - 1)).astype(np.float32)))
        sp = sp // np.array(pools[k])

    def one():
        for q in plist:
            q.grad = None
        onet.ds_loss(onet.unetpp_forward(params, x, pools), tg).backward()

    one()
    t0, n = time.perf_counter(), 0
    while n < 3 or (time.perf_counter() - t0 < budget_s and n < 20):
        one()
        n += 1
    dt = (time.perf_counter() - t0) / n
    return {"value": 1.0 / dt, "unit": "patches/s", "ms_per_patch": dt * 1e3, "cores": os.cpu_count(), "kind": "port",
            "sample": "BASELINE.json configs[0] unscaled: oracle fwd + DS loss + bwd on one 1x1x40x56x40 patch, fp32, "
                      "%d timed iterations" % n}


def torch_cudnn_baseline(dev, steps=4):
    """the 'honest before' (SURVEY 2.3 / 8d): the reference's stock PyTorch / cuDNN path on the SAME B200.  The
    oracle's functional restatement of the reference network (torch ops: F.pad-free slicing shift, torch.cat,
    cuDNN Conv3d / ConvTranspose3d, InstanceNorm, LeakyReLU, MaxPool3d) runs on CUDA under autocast exactly as
    nnUNetTrainer_simple.run_iteration does (fp16 + GradScaler, :549-564) and under bf16 autocast: config 2
    training iteration (B=2) and the config-3 per-tile forward.  A comparator only: nothing here is shipped."""
    import numpy as np
    import torch
    import random
    from collections import OrderedDict
    from oracle import masking as omask
    from oracle import network as onet
    from e2enet_medical_b200.training import POOLS, synthetic_batch
    pools = POOLS["btcv"]
    out = {"what": "stock torch %s / cuDNN %s on this GPU through oracle.network (functional restatement of the "
                   "reference modules), TF32 allowed as torch defaults" % (torch.__version__, torch.backends.cudnn.version())}
    torch.backends.cudnn.benchmark = True
    shapes = onet.param_shapes(IN_CH, 48, NCLS, pools)
    params = OrderedDict((k, v.to(dev)) for k, v in onet.det_params(shapes, seed=0).items())
    random.seed(0)
    masks = {k: torch.from_numpy(m).to(dev) for k, m in omask.init_uniform(shapes, DENSITY).items()}
    with torch.no_grad():
        for k, m in masks.items():
            params[k].mul_(m)
    plist = [v.requires_grad_(True) for v in params.values()]
    data, targets = synthetic_batch(BATCH, IN_CH, NCLS, PATCH, pools, seed=1)
    x, tg = data.to(dev), [t.to(dev) for t in targets]

    def ev():
        return torch.cuda.Event(enable_timing=True)

    for mode in ("fp16", "bf16"):
        opt = torch.optim.SGD(plist, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
        scaler = torch.amp.GradScaler("cuda", enabled=(mode == "fp16"))
        dt_ = torch.float16 if mode == "fp16" else torch.bfloat16

        def one():
            opt.zero_grad()
            with torch.autocast("cuda", dtype=dt_):
                outs = onet.unetpp_forward(params, x, pools)
                l = onet.ds_loss(outs, tg)
            scaler.scale(l).backward()
            scaler.unscale_(opt)
            torch.nn.utils.clip_grad_norm_(plist, 12)
            scaler.step(opt)
            scaler.update()
            with torch.no_grad():                       # Masking.apply_mask as the reference does it (core_channel.py:427-434)
                for k, m in masks.items():
                    params[k].data = params[k].data * m
                    st = opt.state[params[k]]
                    if 'momentum_buffer' in st:
                        st['momentum_buffer'] = st['momentum_buffer'] * m
            return l

        try:
            for _ in range(3):
                one()
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(steps):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out["train_" + mode] = {"ms_per_step": ms, "patches_per_s": BATCH / (ms / 1e3),
                                    "peak_mem_GB": torch.cuda.max_memory_allocated(dev) / 2 ** 30}
        except Exception as e:                          # noqa: BLE001 -- a comparator must not kill the bench line
            out["train_" + mode] = {"error": repr(e)[:200]}
        del opt
        for q in plist:
            q.grad = None
        torch.cuda.empty_cache()
    # config-3 per-tile forward (B=1 tile, 16 classes), as predict_3D runs it: autocast + no_grad
    shapes16 = onet.param_shapes(1, 48, 16, pools)
    p16 = OrderedDict((k, v.to(dev)) for k, v in onet.det_params(shapes16, seed=0).items())
    tile = torch.randn((1, 1) + PATCH, device=dev)
    for mode in ("fp16", "bf16"):
        dt_ = torch.float16 if mode == "fp16" else torch.bfloat16
        try:
            with torch.no_grad(), torch.autocast("cuda", dtype=dt_):
                for _ in range(3):
                    onet.unetpp_forward(p16, tile, pools, deep_supervision=False)
                torch.cuda.synchronize()
                e0, e1 = ev(), ev()
                e0.record()
                for _ in range(steps):
                    torch.softmax(onet.unetpp_forward(p16, tile, pools, deep_supervision=False), 1)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out["tile_forward_" + mode] = {"ms_per_tile": ms, "voxels_per_s_324_tiles": 78643200.0 / (324 * ms / 1e3)}
        except Exception as e:                          # noqa: BLE001
            out["tile_forward_" + mode] = {"error": repr(e)[:200]}
    torch.backends.cudnn.benchmark = False
    return out


def masking_update_leg(ts, dev):
    """BASELINE.json configs[4] / SURVEY 8(d): wall time and kernel launches of Masking.step() WITH a prune /
    regrow update on the config-2 network (35 masked tensors, 2 072 832 kernels), and without."""
    import torch
    import random
    from e2enet_medical_b200 import _lib
    m = ts.mask
    keep = m.prune_every_k_steps
    res = {}
    for tag, every in (("step_no_update", None), ("step_with_update", 1)):
        m.prune_every_k_steps = every
        random.seed(1)
        m.step()                                       # warm-up (allocations)
        torch.cuda.synchronize()
        n0, t0 = _lib.launch_count(), time.perf_counter()
        reps = 3
        for _ in range(reps):
            m.step()
        torch.cuda.synchronize()
        res[tag] = {"wall_ms": (time.perf_counter() - t0) / reps * 1e3, "launches": (_lib.launch_count() - n0) / reps}
    m.prune_every_k_steps = keep
    res["what"] = ("Masking.step() on the config-2 network: apply_mask + death-rate decay (+ kernel_death for all 35 "
                   "tensors, host random.sample growth, kernel_growth, apply_mask, counts when updating); wall clock "
                   "incl. the host-side sampling and the read-backs the reference's bookkeeping needs")
    return res


def measured_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture of this round (never typed in):
    profiles/*_traffic.json = {"kernel": ..., "dram_read_bytes": ..., "dram_write_bytes": ..., "algorithmic_bytes": ...}"""
    import glob
    cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not cands:
        return None, None
    d = json.load(open(cands[-1]))
    return float(d["dram_read_bytes"]) + float(d["dram_write_bytes"]), \
        "%s (%s; algorithmic bytes %.4g)" % (d.get("kernel"), os.path.basename(cands[-1]), d.get("algorithmic_bytes", 0))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" = the bounded sample below; warm-up folded into cpu_baseline_patches_per_s
    val, dt, n = cpu_baseline_patches_per_s(budget_s=max(10.0, 6.0 * max(args.steps, 1)))
    line = {
        "impl": "reference", "metric": "train patches/s", "value": val, "unit": "patches/s", "n_gpus": args.gpus,
        "steps": n, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "E2ENet BTCV-shaped training: batch 2, 1x64x160x160, 14 classes, density 0.2 "
                               "(CPU sample: 1x1x32x96x96 crop per step, value scaled by voxel count)"},
        "cpu_baseline": {"value": val, "unit": "patches/s", "cores": os.cpu_count(), "kind": "port",
                         "sample": "oracle (torch-CPU fp32 restatement of the reference path) fwd+DS loss+bwd+clip+SGD+"
                                   "apply_mask on one 1x1x32x96x96 crop = 0.18 patch, %d timed steps" % n},
        "e2e": {"value": val, "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ---------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from e2enet_medical_b200 import _lib, ops
    from e2enet_medical_b200.training import POOLS, TrainStep, synthetic_batch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_precision(args.precision)
    _lib.load()
    ops.CONFIG["impl"] = args.kernel_impl
    pools = POOLS["btcv"]
    ts = TrainStep(IN_CH, NCLS, pools, PATCH, DENSITY, 0.5, 1200, dev, world, seed=0)
    data_h, targets_h = synthetic_batch(BATCH, IN_CH, NCLS, PATCH, pools, seed=1 + rank)
    data_h = data_h.pin_memory()
    targets_h = [t.pin_memory() for t in targets_h]
    data_d = data_h.to(dev)
    targets_d = [t.to(dev) for t in targets_h]
    h2d = data_h.numel() * 4 + sum(t.numel() * 4 for t in targets_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, finish=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if finish is not None:
            finish()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def step_resident():
        if ts._graph is not None:
            ts.step(ts._static_data, ts._static_targets)        # inputs already in the graph's static buffers
        else:
            ts.step(data_d, targets_d)

    # two pinned host batches alternate: while step i computes, the copy of step i+1's batch is already in flight
    # (TrainStep.prefetch); every step's inputs still cross PCIe inside the timed region, and every step's loss is
    # read back before the next step starts (like run_iteration's l.detach().cpu().numpy())
    host_batches = [(data_h, targets_h), (data_h.clone().pin_memory(), [t.clone().pin_memory() for t in targets_h])]
    e2e_state = {"i": 0, "pending": None, "losses": []}

    def step_e2e():
        i = e2e_state["i"]
        d_h, t_h = host_batches[i % 2]
        if ts._graph is not None or ts.fused_optimizer:
            if i == 0:
                ts.prefetch(d_h, t_h)
            l = ts.step(d_h, t_h)                  # waits for the staged copy, device-side move into the static inputs, replay
            ts.prefetch(*host_batches[(i + 1) % 2])
        else:
            d = d_h.to(dev, non_blocking=True)
            t = [x.to(dev, non_blocking=True) for x in t_h]
            l = ts.step(d, t)
        e2e_state["i"] = i + 1
        # every step's loss is read back to the host, one iteration deferred: step i's loss is fetched right after
        # step i+1 has been enqueued, so the host work of the next iteration overlaps the device instead of idling it
        # (logging lags by one iteration; nothing in the loop depends on the value)
        if e2e_state["pending"] is not None:
            e2e_state["losses"].append(float(e2e_state["pending"].cpu()))
        e2e_state["pending"] = l

    def finish_e2e():
        if e2e_state["pending"] is not None:
            e2e_state["losses"].append(float(e2e_state["pending"].cpu()))
            e2e_state["pending"] = None

    for _ in range(max(args.warmup, 3)):
        step_resident()
    # (1) eager pass with per-launch CUDA events: kernel-level timing of the GEMM family for the roofline
    # (the side stream of the weight gradients is switched off for this pass: a kernel's own duration is wanted, and
    # events around a launch that shares the GPU with another stream would also count the time it waits for SMs)
    side_default = ops.CONFIG["wgrad_side_stream"]
    ops.CONFIG["wgrad_side_stream"] = False
    ops.PROFILE["enabled"] = True
    ops.PROFILE["records"] = []
    n0 = _lib.launch_count()
    ms_eager = timed(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    recs = ops.PROFILE["records"]
    ops.PROFILE["enabled"] = False
    ops.CONFIG["wgrad_side_stream"] = side_default
    torch.cuda.synchronize()
    gemm_ms = sum(a.elapsed_time(b) for (_, a, b, _) in recs)
    gemm_flops = sum(f for (_, _, _, f) in recs)
    per_kind = {}
    for kind in ("gemm", "wgrad"):
        kms = sum(a.elapsed_time(b) for (k, a, b, _) in recs if k == kind)
        kfl = sum(f for (k, _, _, f) in recs if k == kind)
        per_kind[kind] = {"launches_per_step": sum(1 for r in recs if r[0] == kind) / max(args.steps, 1),
                          "ms_per_step": kms / max(args.steps, 1),
                          "achieved_tflops": (kfl / 1e12) / (kms / 1e3) if kms > 0 else 0.0}
    # (2) the headline: the whole iteration captured in ONE CUDA graph (forward, loss, backward, gradient
    # all-reduce, clip, SGD, apply_mask) and replayed; Masking's host bookkeeping stays eager
    graphed = False
    if not args.no_graph:
        try:
            ts.enable_graph(data_d, targets_d, warmup=2)
            graphed = True
        except Exception as e:                     # keep the eager numbers rather than losing the bench line
            print("CUDA graph capture failed, staying eager: %r" % (e,), file=sys.stderr)
            ts._graph = None
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(2):
        step_resident()
    ms = timed(step_resident, args.steps)
    if graphed:
        launches = ts.graph_launches * args.steps
    ms_e2e = timed(step_e2e, args.steps, finish_e2e)
    assert len(e2e_state["losses"]) == args.steps and all(v == v for v in e2e_state["losses"])
    sampler.stop_flag = True

    if rank != 0:
        if not args.no_inference:
            inference_leg(dev, world, rank, args)
        _finish_ranks(ts, world)
        return
    peaks = measured_peaks()
    value = BATCH * world * args.steps / (ms / 1e3)
    e2e = BATCH * world * args.steps / (ms_e2e / 1e3)
    achieved = (gemm_flops / 1e12) / (gemm_ms / 1e3) if gemm_ms > 0 else 0.0
    line = {
        "metric": "train patches/s", "value": value, "unit": "patches/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": "E2ENet BTCV-shaped training: batch 2 per GPU, 1x64x160x160, 14 classes, density 0.2, "
                               "SGD nesterov + clip 12 + Masking.step() (BASELINE.json configs[1])",
                   "global_batch": BATCH * world, "parallelism": "dp%d" % world,
                   "l2": "activations per step (>10 GB) exceed the 126 MB L2; no explicit flush needed",
                   "kernel_impl": "tcgen05" if args.kernel_impl else "mma.sync"},
        "e2e": {"value": e2e, "unit": "patches/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "how": "TrainStep.step() on pinned host batches: the next batch's H2D copy is prefetched on a copy stream "
                       "while the current step computes; every step's loss is read back, one iteration deferred"},
        "gpu_launches": int(launches), "cuda_graph": graphed, "eager_ms_per_step": ms_eager / args.steps,
        "clocks": sampler.summary(),
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["tf_sustained"],
                     # dram bytes per launch of the family's largest launch, read from the committed ncu capture
                     # (profiles/r*_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep); null if absent
                     "traffic": measured_traffic()[0],
                     "traffic_of": measured_traffic()[1],
                     "kernel": "tcgen05 conv / transposed-conv / 1x1 GEMM launches of one step: conv_tc_kernel (fwd + "
                               "dgrad; key 'gemm') and wgrad_tc_kernel (key 'wgrad'); dense 2MNK FLOPs",
                     "per_kind": {k: dict(v, frac=v["achieved_tflops"] / peaks["tf_sustained"]) for k, v in per_kind.items()},
                     "share_of_step": gemm_ms / ms_eager if ms_eager > 0 else None,
                     "timed_in": "eager, single-stream steps of this run (per-launch CUDA events on the launching stream); `value` "
                                 "is the same iteration replayed as one CUDA graph in which the weight-gradient GEMMs run on "
                                 "a side stream and overlap the InstanceNorm backward kernels", "peak_source": peaks["src"] + ", sustained bf16"},
        "step_tflops": STEP_GFLOP_B2 / 1e3 * args.steps / (ms / 1e3),
    }
    if not args.no_inference:
        line["inference"] = inference_leg(dev, world, rank, args)
    if world == 1 and not args.no_extras:
        try:
            line["masking_update"] = masking_update_leg(ts, dev)
        except Exception as e:                          # noqa: BLE001
            line["masking_update"] = {"error": repr(e)[:200]}
        ts._graph = None
        del ts
        torch.cuda.empty_cache()
        if args.precision == "bf16":
            try:
                line["fp16_mode"] = fp16_mode_leg(dev, data_d, targets_d, args)
            except Exception as e:                      # noqa: BLE001
                line["fp16_mode"] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
        line["torch_cudnn_baseline"] = torch_cudnn_baseline(dev)
        tb = line["torch_cudnn_baseline"].get("train_fp16", {})
        if "patches_per_s" in tb:
            line["speedup_vs_torch_cudnn_fp16"] = {"device": value / tb["patches_per_s"], "e2e": e2e / tb["patches_per_s"]}
        ts = None
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline_config0"] = cpu_baseline_config0()
        v, dt, n = cpu_baseline_patches_per_s()
        line["cpu_baseline"] = {"value": v, "unit": "patches/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "oracle fwd+DS loss+bwd+clip+SGD+apply_mask on one 1x1x32x96x96 crop "
                                          "(0.18 patch), %d steps of %.1f s" % (n, dt)}
    _emit(line)
    if ts is not None:
        _finish_ranks(ts, world)


def fp16_mode_leg(dev, data_d, targets_d, args):
    """the same iteration on the fp16 build of the library (libe2enet_b200_fp16.so: fp16 activations / gradients /
    packed weights = the reference's shipped AMP arithmetic, with the GradScaler as device state of the fused
    optimizer): same kernels, same tcgen05 rate -- reported so that the precision option has a measured cost"""
    import torch
    from e2enet_medical_b200 import ops
    from e2enet_medical_b200.training import POOLS, TrainStep
    ops.set_precision("fp16")
    try:
        ts = TrainStep(IN_CH, NCLS, POOLS["btcv"], PATCH, DENSITY, 0.5, 1200, dev, 1, seed=0)
        for _ in range(3):
            ts.step(data_d, targets_d)
        ts.enable_graph(data_d, targets_d, warmup=2)
        for _ in range(2):
            ts.step(ts._static_data, ts._static_targets)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            l = ts.step(ts._static_data, ts._static_targets)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out = {"ms_per_step": ms, "patches_per_s": BATCH / (ms / 1e3), "loss": float(l),
               "loss_scale": float(ts.optimizer.loss_scale()), "skipped_last_step": bool(float(ts.optimizer._norm_coef[2])),
               "library": "libe2enet_b200_fp16.so", "parity": "profiles/r02_parity_fullsize*.json (fp16 rows)"}
        ts._graph = None
        del ts
        return out
    finally:
        ops.set_precision("bf16")


def _finish_ranks(ts, world):
    """multi-rank teardown: a captured CUDA graph holds NCCL kernels of the process group, and destroying
    the communicator under it can block forever -- drop the graph, synchronise, and leave without the
    NCCL destructor (the processes are exiting anyway)"""
    if world <= 1:
        return
    import torch
    import torch.distributed as dist
    ts._graph = None
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def inference_leg(dev, world, rank, args):
    """BASELINE.json configs[2]: AMOS-CT-shaped sliding-window inference, 300x512x512 volume, 16 classes,
    patch 64x160x160, step 0.5 (324 tiles), Gaussian weighting, no mirroring; tiles sharded over the
    ranks with one NCCL all-reduce of the accumulators.  Timed through predict_3D (NumPy in, NumPy
    out: H2D of the volume and D2H of labels + softmax are inside the timed region)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from e2enet_medical_b200.training import POOLS, build_network
    from e2enet_medical_b200.network_architecture.unetpp_d import softmax_helper
    vol_shape = tuple(args.infer_volume)
    torch.manual_seed(0)
    net = build_network(1, 16, POOLS["btcv"], PATCH, 48, deep_supervision=True).to(dev)
    net.eval()
    net.do_ds = False
    net.inference_apply_nonlin = softmax_helper
    # N > 1: slab ownership -- contiguous tile ranges per rank (each rank uploads only the x-planes its tiles read),
    # neighbour exchange of the overlap planes, every rank finalises its own x-slab, and the finalised slabs are
    # gathered GPU -> GPU to rank 0, whose predict_3D returns the FULL (seg, softmax) like the reference's
    net.set_tile_sharding(rank, world, None, "gather" if world > 1 else None)
    vol = np.random.RandomState(0).randn(1, *vol_shape).astype(np.float32)
    small = vol[:, :64, :160, :320].copy()                       # warm-up: 3 tiles
    net.predict_3D(small, False, (0, 1, 2), True, 0.5, PATCH, None, True, "constant", None, True, False, True)
    # un-timed full-size pass: allocates the pooled pinned result buffers (reused by the timed pass because this
    # pass's results are dropped) and warms NCCL's P2P channels
    r = net.predict_3D(vol, False, (0, 1, 2), True, 0.5, PATCH, None, True, "constant", None, True, False, True)
    del r
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    seg, probs = net.predict_3D(vol, False, (0, 1, 2), True, 0.5, PATCH, None, True, "constant", None, True, False,
                                True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()                     # the volume is done when rank 0 holds the full result
    dt = time.perf_counter() - t0
    if rank == 0:
        assert seg.shape == vol_shape and probs.shape == (16,) + vol_shape, (seg.shape, probs.shape)
    allocs_timed = getattr(net, "_pinned_allocs", 0)
    host_phases = None
    if rank == 0 and world == 1:
        del seg, probs
        net.profile_phases = True          # one more, diagnostic pass: synchronise after every phase and time it
        seg, probs = net.predict_3D(vol, False, (0, 1, 2), True, 0.5, PATCH, None, True, "constant", None, True, False, True)
        net.profile_phases = False
        host_phases = dict(getattr(net, "_last_host_phases", {}), pinned_buffer_allocations=allocs_timed)
    # device-only time of the tile loop (volume resident, results left on the device)
    dev_ms = net._last_tile_loop_ms
    n_tiles = net._last_num_tiles
    pe = net._last_phase_events
    phases = {"tile_loop_ms": pe[0].elapsed_time(pe[1])}
    if len(pe) > 2:
        phases["exchange_ms_incl_wait_for_neighbours"] = pe[1].elapsed_time(pe[2])
    if world > 1:
        t = torch.tensor([dt, dev_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, dev_ms = float(t[0]), float(t[1])
    nvox = float(np.prod(vol_shape))
    gather_how = getattr(net, "_last_gather", None)
    seg_bytes = int(seg.nbytes + probs.nbytes) if seg is not None else 0
    del seg, probs
    net.release_shared_result_segments()
    return {"metric": "inference voxels/s", "value": nvox / (dev_ms / 1e3), "unit": "voxels/s",
            "e2e": {"value": nvox / dt, "unit": "voxels/s", "h2d_bytes_per_step": int(net._last_h2d_bytes),
                    "d2h_bytes_per_step": seg_bytes},
            "ms_per_volume": dev_ms, "e2e_ms_per_volume": dt * 1e3, "tiles": n_tiles, "rank0_phases": phases,
            "rank0_host_phases_diagnostic_pass": host_phases,
            "tiles_per_s": n_tiles / (dev_ms / 1e3),
            "config": {"workload": "E2ENet AMOS-CT-shaped sliding-window inference: %dx%dx%d volume, 16 classes, patch "
                                   "64x160x160, step 0.5, gaussian, no mirroring (BASELINE.json configs[2]); value = "
                                   "tile loop + reduce + finalise on resident data (max over ranks), e2e = predict_3D NumPy -> "
                                   "NumPy on rank 0: FULL (seg int64, softmax fp32) like the reference, into pooled pinned "
                                   "buffers that never alias across calls" % vol_shape, "tiles_sharded_over": world,
                       "exchange": "none" if world == 1 else
                       "slab ownership: every rank uploads only its x-planes, NCCL point-to-point exchange of the overlap "
                       "planes between neighbouring ranks, finalised slabs delivered to rank 0's result arrays: " +
                       str(gather_how)}}


_RESULT_OUT = None


def _emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kernel-impl", type=int, default=1, help="0: mma.sync gather kernels, 1: tcgen05 where available")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"],
                    help="16-bit type of activations / gradients / packed weights (bf16 = north_star's; fp16 = the reference's AMP)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the Masking-update and stock torch/cuDNN comparator legs")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps only (no whole-step CUDA graph)")
    ap.add_argument("--no-inference", action="store_true", help="skip the sliding-window inference leg")
    ap.add_argument("--infer-volume", type=int, nargs=3, default=[300, 512, 512])
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON result): everything the mirrored reference classes print
    # (Masking's density tables, the poly-LR notices) goes to stderr
    global _RESULT_OUT
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
