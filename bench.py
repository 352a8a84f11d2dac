#!/usr/bin/env python
"""Benchmark of the plane-sweep hot path (BASELINE.json metric: cost-volume Mvox/s + depth maps/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload = BASELINE.json configs[1]: MVSNet (variance), 1 reference + 4 source views, 640x512 images ->
32-channel 160x128 feature maps, D = 192 hypotheses (3.93 Mvox per depth map).  One "step" = one pass of the
hot path (fused warp+variance cost volume -> 3-D U-Net regulariser -> softmax/depth/confidence) for ONE
reference view per GPU; with N GPUs every rank processes its own reference view (weak scaling, independent
units), the regression kernel writes the depth map into the rank's slice of a preallocated gather buffer and ONE
in-place NCCL all-gather of the per-view depth maps closes the step -- all of it inside one CUDA graph per step.

`value` is measured with the feature maps already resident in HBM; `e2e` is the same metric through
MVSNet.streamed(...).submit with the features in pinned host memory (H2D + hot path + all-gather + D2H of the gathered
depth maps and the confidence inside the timed region).  Further keys of the line: `roofline` / `kernels` (per-kernel
CUDA-event times against the measured HBM peak), `cpu_baseline` (the unmodified reference on the host cores),
`gpu_reference` (the unmodified reference through stock PyTorch / cuDNN on the same B200: the existing Blackwell path),
`depth_maps_per_s` (full forward from images), `cfg5` (BASELINE configs[4]: 64 Vis-MVSNet reference views sharded over
the N GPUs + one all-gather).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K2_ENGINE_NOTE = {
    "zm": "zm (persistent z-march tcgen05 kind::f16, fp16x2 split after abs-max scaling = fp32-equivalent; tile engine for layers it does not cover)",
    "tc": "tc (tcgen05 kind::tf32, 3xTF32 split = fp32-equivalent)",
    "tc_tf32": "tc_tf32 (tcgen05 kind::tf32 single pass; outside the parity bar)",
    "fp32": "fp32 (CUDA cores)",
}
# library kernels per step: geometry prologue, K1, conv0..conv5, conv6 (two launches: split over input channels),
# conv7, conv9, conv11, prob (C1 head), K3
LAUNCHES_PER_STEP = 15
CFG = {"name": "cfg2", "views": 5, "C": 32, "h": 128, "w": 160, "D": 192}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip gpu_reference, depth_maps_per_s and cfg5 (profiling runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch the kernels of a step one by one instead of replaying a CUDA graph")
    return ap.parse_args()


def voxels():
    return CFG["D"] * CFG["h"] * CFG["w"]


def algorithmic_bytes():
    """Layer-wise compulsory fp32 traffic of one cfg2 pass: every seam tensor written once by its
    producer and read once per consumer, BN/ReLU/skip fused.  Returns (total, per-kernel dict)."""
    V, C, h, w, D = CFG["views"], CFG["C"], CFG["h"], CFG["w"], CFG["D"]
    vox = D * h * w
    per = {"k1_cost_volume": 4 * (V * C * h * w + C * vox)}
    # (name, cin, cout, in_vox_div, out_vox_div, skip)
    layers = [("conv0", 32, 8, 1, 1, 0), ("conv1", 8, 16, 1, 8, 0), ("conv2", 16, 16, 8, 8, 0), ("conv3", 16, 32, 8, 64, 0),
              ("conv4", 32, 32, 64, 64, 0), ("conv5", 32, 64, 64, 512, 0), ("conv6", 64, 64, 512, 512, 0),
              ("conv7", 64, 32, 512, 64, 1), ("conv9", 32, 16, 64, 8, 1), ("conv11", 16, 8, 8, 1, 1), ("prob", 8, 1, 1, 1, 0)]
    for name, cin, cout, di, do, skip in layers:
        per[name] = 4 * (cin * vox // di + cout * vox // do * (1 + skip))
    per["k3_regress"] = 4 * (vox + 2 * h * w)
    return sum(per.values()), per


# ------------------------------------------------------------------------------------------------
# clocks sampler (pynvml): rank 0 only, 100 Hz, plus one sample at either end of every sampled region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index, enabled=True, period_s=0.01):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.period_s = period_s
        self.nv = None
        if enabled:
            try:
                import pynvml
                pynvml.nvmlInit()
                self.nv = pynvml
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
                self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            except Exception:
                self.nv = None
        self.t = None

    def sample(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(self.nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _run(self):
        while not self._stop.wait(self.period_s):
            self.sample()

    def __enter__(self):
        if self.nv:
            self._stop.clear()
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        return self

    def __exit__(self, *a):
        if self.nv:
            # the GPU is still executing the queued steps here: these samples are under load
            self.sample()
            torch.cuda.synchronize()
            self.sample()
            self._stop.set()
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s),
                "sampler": "NVML on rank 0, %.0f Hz, during the timed `value` and `e2e` regions" % (1 / self.period_s)}


# ------------------------------------------------------------------------------------------------
# workload construction (shared by both arms)
# ------------------------------------------------------------------------------------------------
def make_workload(seed):
    """(net, feats [V x [1,C,h,w]], cams (K full-res, R, t, dmin, dmax), projs [1,V,4,4], depth [1,D])."""
    from wild_deep_mvs_b200 import synth
    from wild_deep_mvs_b200.mvsnet import MVSNet, build_proj_matrices
    torch.manual_seed(0)
    net = MVSNet("variance")
    synth.randomize_norm_stats(net, seed=1)
    synth.scale_param(net.cost_regularization.prob.weight, 40.0)
    net.num_depth = CFG["D"]
    net.eval()
    feats = synth.make_features(1, CFG["views"], CFG["C"], CFG["h"], CFG["w"], seed=seed)
    cams = synth.make_cameras(1, CFG["views"], 4 * CFG["h"], 4 * CFG["w"])
    K, R, t, dmin, dmax = cams
    K = K.clone()
    K[:, :, :2] /= 4
    projs = build_proj_matrices(K, R, t)
    D = CFG["D"]
    depth = dmin[:, :1] + (dmax[:, :1] - dmin[:, :1]) / (D - 1) * torch.arange(D).view(1, -1)
    return net, feats, cams, projs, depth


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the UNMODIFIED reference (oracle/_ref) on the host cores
# ------------------------------------------------------------------------------------------------
def load_reference():
    """The unmodified reference (oracle/ref_import.py: /root/reference in the build container, the packed archive
    oracle/_ref/ref_hotpath.zip on the GPU box) or None."""
    try:
        from oracle import ref_import
        if ref_import.reference_location() is None:
            return None
        return ref_import.import_reference(cuda_shim=False)
    except Exception as e:  # noqa: BLE001 -- reported in the line
        sys.stderr.write("reference import failed: %r\n" % (e,))
        return None


def cpu_pass_fn(d_sample):
    """Returns (fn() -> seconds for one pass at D = d_sample hypotheses, kind, description)."""
    net, feats, cams, projs, depth = make_workload(seed=0)
    K, R, t, dmin, dmax = cams
    ref = load_reference()
    if ref is not None:
        from oracle import ref_run
        rnet = ref_run.reference_mvsnet(ref, net.state_dict(), "variance", d_sample)
        # the same hypotheses as the first d_sample planes of the full sweep: depth_max of the truncated sweep
        dmax_s = dmin + (dmax - dmin) / (CFG["D"] - 1) * (d_sample - 1)

        def fn():
            t0 = time.perf_counter()
            ref_run.mvsnet_from_features(rnet, feats, K, R, t, dmin, dmax_s)
            return time.perf_counter() - t0
        return fn, "reference", "unmodified reference (oracle/_ref: models/MVSNet/model.py forward from the feature maps on)"
    from oracle import torch_port as tp
    sd = net.state_dict()
    pl = list(torch.unbind(projs, 1))

    def fn():
        with torch.no_grad():
            t0 = time.perf_counter()
            tp.mvsnet_hot_path(sd, feats, pl, depth[:, :d_sample].contiguous(), "variance", None, {})
            return time.perf_counter() - t0
    return fn, "port", "oracle/torch_port.py (ATen restatement of the reference path; the reference archive is missing)"


def run_cpu_baseline(budget_s, steps, warmup):
    """Times the reference's CPU implementation of the path on all host cores.  The sample is the full cfg2 pass unless
    (steps+warmup) passes would exceed `budget_s`; then the depth axis is cut to a multiple of 8 that fits."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    D = CFG["D"]
    probe, kind, what = cpu_pass_fn(48)
    probe()                                   # warms the allocator / thread pool
    est_full = probe() * D / 48               # a quarter slab, scaled
    d_sample = D
    if est_full * (steps + warmup) > budget_s:
        d_sample = max(8, int(D * budget_s / (est_full * (steps + warmup))) // 8 * 8)
    fn, kind, what = cpu_pass_fn(d_sample)
    for _ in range(warmup):
        fn()
    times = [fn() for _ in range(steps)]
    vox = d_sample * CFG["h"] * CFG["w"]
    mean = sum(times) / len(times)
    return {"value": vox / mean / 1e6, "unit": "Mvox/s", "cores": cores, "kind": kind,
            "sample": "%s, %d timed pass(es) of cfg2 with D=%d of %d hypotheses (%.3f Mvox each), %d threads"
                      % (what, steps, d_sample, D, vox / 1e6, torch.get_num_threads()),
            "ms_per_step": mean * 1e3}, mean, vox


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, mean, vox = run_cpu_baseline(budget_s=150.0, steps=args.steps, warmup=args.warmup)
    line = {"impl": "reference", "metric": "cost-volume Mvox/s (build+regularise+regress)", "value": cb["value"], "unit": "Mvox/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.gpus), "maps_per_s": 1.0 / (mean * CFG["D"] * CFG["h"] * CFG["w"] / vox),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "Mvox/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n):
    return {"workload": "BASELINE cfg2: MVSNet variance, 1 ref + 4 src views, 640x512 images -> 32ch 160x128 features, D=192, "
                        "features->depth+confidence (K1 warp+variance, K2 3-D U-Net, K3 softmax/regress)",
            "views": CFG["views"], "feature_hw": [CFG["h"], CFG["w"]], "D": CFG["D"], "voxels_per_map": voxels(),
            "maps_per_step": n, "k2_engine": K2_ENGINE_NOTE.get(os.environ.get("MVSB200_K2_ENGINE", "zm"), "?"),
            "parallelism": "view-sharded replicas x%d + 1 in-place all-gather of depth maps per step (inside the step's CUDA graph; "
                           "backend: %s)" % (n, "mvsb200_allgather_depth on the library's own NCCL communicator"
                                             if os.environ.get("MVSB200_GATHER", "lib") == "lib" else "torch.distributed all_gather_into_tensor"),
            "l2": "inputs larger than L2: every step streams ~1.4 GB of intermediates (503 MB cost volume, 389 MB activations) "
                  "through the 126 MB L2 before the next step re-reads its 13 MB of feature maps; K steps back to back between one "
                  "pair of CUDA events, L2 flushed once before the first (per-kernel `kernels` timings: flushed before each)"}


# ------------------------------------------------------------------------------------------------
# own arm
# ------------------------------------------------------------------------------------------------
def own_arm(args):
    import torch.distributed as dist
    from wild_deep_mvs_b200 import ops, shard
    from wild_deep_mvs_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback exists for the hot path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL writes its version banner there when NCCL_DEBUG is VERSION or
        # above, so the communicator is created (init + a first collective) with file descriptor 1 pointing at stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    L.load()

    net, feats, cams, projs, depth = make_workload(seed=rank)  # each rank: its own reference view sample
    net = net.to(dev)
    dfeats = [ops.to_nhwc(f.to(dev)) for f in feats]
    dprojs = list(torch.unbind(projs.to(dev), 1))
    ddepth = depth.to(dev)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    h, w = CFG["h"], CFG["w"]

    # K3 writes into gather.local (this rank's slice of the [world,h,w] buffer); the in-place all-gather is captured
    gather = shard.DepthGather(1, (h, w), dev) if world > 1 else None
    graph = None if args.no_graph else net.graphed(dfeats, dprojs, ddepth, gather=gather)   # one cudaGraphLaunch per step

    def step():
        if graph is not None:
            graph()
            return graph.gathered if gather is not None else graph.depth
        d, c = net.depth_from_features(dfeats, dprojs, ddepth, out_depth=gather.local if gather is not None else None)
        return gather.all_gather() if gather is not None else d

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local, enabled=(rank == 0))   # NVML init takes ~0.1 s: BEFORE the barrier, or rank 0 enters the timed
    for _ in range(args.warmup):                        # region late and every other rank bills the wait to its first all-gather
        step()
    barrier()
    # ONE event pair around the K steps, launched back to back.  No L2 flush between steps: a step streams 1.4 GB of
    # intermediates (503 MB cost volume, 389 MB activations, ...) through the 126 MB L2, so nothing a step reads first
    # (13 MB of feature maps, last touched 1.4 GB of traffic earlier) is still resident -- the "inputs larger than L2"
    # case of the timing rules.  (A 512 MiB memset between steps is not neutral at N > 1: it saturates this GPU's HBM
    # while a peer's all-gather is still reading from it over NVLink, which bills ~0.1 ms of benchmark artefact per step.)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    with clocks:
        ea.record()
        for _ in range(args.steps):
            gathered = step()
        eb.record()
    barrier()
    t = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = t.item()
    ms_per_step = total_ms / args.steps
    value = world * voxels() / (ms_per_step * 1e-3) / 1e6

    # ---- content of the gathered maps: slot r everywhere == what rank r computed; and rank 0 recomputes the LAST rank's map
    # from that rank's seeded inputs on its own GPU (a single-GPU run of the same view) --------------------------------------
    gather_check = None
    if world > 1:
        mine = gathered[rank].clone()
        own, _ = net.depth_from_features(dfeats, dprojs, ddepth)
        assert torch.equal(mine, own[0]), "rank %d: its slot of the gathered buffer differs from its own depth map" % rank
        sums = gathered.double().sum(dim=(1, 2))
        every = [torch.empty_like(sums) for _ in range(world)]
        dist.all_gather(every, sums)
        assert all(torch.equal(e, sums) for e in every), "ranks hold different gathered buffers"
        if rank == 0:
            _, f2, _, p2, d2 = make_workload(seed=world - 1)
            other, _ = net.depth_from_features([ops.to_nhwc(f.to(dev)) for f in f2], list(torch.unbind(p2.to(dev), 1)), d2.to(dev))
            gather_check = {"gathered[last rank] == that view computed on rank 0 alone": bool(torch.equal(gathered[world - 1], other[0])),
                            "max_abs_diff": float((gathered[world - 1] - other[0]).abs().max())}
            assert gather_check["max_abs_diff"] <= 1e-4 * float(other.abs().max()), gather_check

    # ---- e2e: pinned host features -> H2D -> hot path (+ all-gather) -> D2H gathered depth maps + confidence ---------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(f.numel() * 4 for f in feats) + projs.numel() * 4 + depth.numel() * 4
        hd = depth.pin_memory()
        host_nhwc = [f.permute(0, 2, 3, 1).contiguous().pin_memory() for f in feats]   # the engine's own layout, packed on the host once
        hprojs = list(torch.unbind(projs.pin_memory(), 1))
        out_d = torch.empty(world, h, w, dtype=torch.float32).pin_memory()
        out_c = torch.empty(1, h, w, dtype=torch.float32).pin_memory()
        streamed = None
        if graph is not None:
            gathers = [shard.DepthGather(1, (h, w), dev) for _ in range(2)] if world > 1 else None
            streamed = net.streamed(dfeats, dprojs, ddepth, gathers=gathers)
        last = {}

        def e2e_step():
            if streamed is not None:
                # pinned host inputs -> (copy stream) H2D into the slot's captured buffers -> one graph launch (K1..K3 +
                # in-place all-gather) -> D2H of the gathered maps + confidence; the copies of this step overlap the
                # previous step's kernels (two slots, round-robin)
                last["d"], last["c"], last["ev"] = streamed.submit(host_nhwc, hprojs, hd)
            else:
                fd = [f.to(dev, non_blocking=True) for f in host_nhwc]
                pj = [q.to(dev, non_blocking=True) for q in hprojs]
                d, c = net.depth_from_features(fd, pj, hd.to(dev, non_blocking=True), out_depth=gather.local if gather is not None else None)
                out_d.copy_(gather.all_gather() if gather is not None else d, non_blocking=True)
                out_c.copy_(c, non_blocking=True)
                last["d"], last["c"] = out_d, out_c

        for _ in range(max(3, args.warmup)):
            e2e_step()
        barrier()
        # ONE event pair around the K steps (no untimed gap a copy could hide in, hence no L2 flush between steps: every
        # step streams 13 MB of fresh inputs and ~1.4 GB of intermediates through the 126 MB L2 anyway)
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        flush.zero_()
        with clocks:
            ea.record()
            for _ in range(args.steps):
                e2e_step()
            eb.record()
        barrier()
        te = torch.tensor([ea.elapsed_time(eb)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_ms = te.item() / args.steps
        # the host copy of the gathered maps holds, in slot `rank`, the map this rank computed from the same inputs
        own, _ = net.depth_from_features(dfeats, dprojs, ddepth)
        torch.cuda.synchronize()
        host_ok = bool(torch.equal(last["d"][rank if world > 1 else 0], own[0].cpu()))
        assert host_ok, "e2e: host copy of the gathered depth maps differs from this rank's depth map"
        if world > 1 and gathered is not None:
            assert torch.equal(last["d"], gathered.cpu()), "e2e: gathered maps differ from the device-resident arm's"
        e2e = {"value": world * voxels() / (e2e_ms * 1e-3) / 1e6, "unit": "Mvox/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": (world + 1) * h * w * 4, "ms_per_step": e2e_ms, "includes_all_gather": world > 1,
               "gathered_maps_checked": host_ok,
               "api": "MVSNet.streamed(...).submit(pinned host NHWC feature maps, projections, hypotheses) -> pinned host depth maps of all "
                      "ranks (gathered) + confidence (two CUDA-graph slots; the H2D copies of step i run on a copy stream during the "
                      "kernels of step i-1)"}

    # ---- per-kernel timing (CUDA events on the launching stream) for the roofline ------------------
    roof, kernels = kernel_roofline(net, dfeats, dprojs, ddepth, flush) if rank == 0 else (None, None)

    # ---- legs that explain the headline: full forward, the existing Blackwell path, cfg5 ------------
    extras = {}
    if not args.no_extras:
        if rank == 0:
            extras["depth_maps"] = full_forward_leg(net, dev)
            if world == 1:
                extras["gpu_reference"] = gpu_reference_leg(net, feats, cams, dev, gathered)
        barrier()
        sys.path.insert(0, os.path.join(ROOT, "profiles"))
        import bench_cfg5
        try:
            extras["cfg5"] = bench_cfg5.run_cfg5(dev, world, rank, hbm_gbs=hbm_peak()[0])
        except AssertionError:
            raise
        except Exception as e:  # noqa: BLE001
            extras["cfg5"] = {"error": repr(e)}
        barrier()

    if rank == 0:
        total_bytes, _ = algorithmic_bytes()
        line = {"metric": "cost-volume Mvox/s (build+regularise+regress)", "value": value, "unit": "Mvox/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(world), "hot_path_maps_per_s": world / (ms_per_step * 1e-3),
                "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": LAUNCHES_PER_STEP * args.steps,
                "roofline": roof, "kernels": kernels,
                "path_hbm": {"algorithmic_bytes_per_map": total_bytes, "achieved_GBps": total_bytes / (ms_per_step * 1e-3) / 1e9,
                             "frac_of_measured_hbm": total_bytes / (ms_per_step * 1e-3) / 1e9 / hbm_peak()[0]},
                "gather_check": gather_check}
        if "depth_maps" in extras:
            line["depth_maps_per_s"] = extras["depth_maps"]["depth_maps_per_s"] * world
            line["depth_maps"] = extras["depth_maps"]
        line["gpu_reference"] = extras.get("gpu_reference")
        line["cfg5"] = extras.get("cfg5")
        if not args.no_cpu_baseline and world == 1:
            cb, _, _ = run_cpu_baseline(budget_s=25.0, steps=1, warmup=1)
            line["cpu_baseline"] = cb
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
        sys.stdout.flush()
    # CUDA graphs that captured NCCL kernels hold references on the communicator: ncclCommDestroy waits for them, so the
    # graphs go first -- and a teardown that still does not return within 20 s must not turn a finished run into a hang
    graph = streamed = None
    import gc
    gc.collect()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

        def teardown():
            shard.DepthGather.destroy_communicators()
            dist.destroy_process_group()
        th = threading.Thread(target=teardown, daemon=True)
        th.start()
        th.join(20.0)
        sys.stderr.flush()
        os._exit(0)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def full_forward_leg(net, dev, reps=30):
    """BASELINE's "depth maps/s": the whole `forward` from IMAGES (2-D FeatureNet on K7 + hot path), one reference view of
    cfg2, replayed as one CUDA graph; and the same call launched eagerly."""
    from wild_deep_mvs_b200 import synth
    s = {k: v.to(dev) for k, v in synth.make_sample(1, CFG["views"], 4 * CFG["h"], 4 * CFG["w"], seed=0).items()}
    call = lambda: net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])

    def timed(fn):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps

    eager_ms = timed(call)
    g = net.graphed_forward(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    graph_ms = timed(g)
    ok = bool(torch.equal(g()["depth"], call()["depth"]))
    return {"depth_maps_per_s": 1e3 / graph_ms, "forward_ms_graphed": graph_ms, "forward_ms_eager": eager_ms,
            "graph_equals_eager": ok,
            "what": "MVSNet.forward(imgs, K, R, t, depth_min, depth_max) from 5 640x512 images resident in HBM, per GPU"}


def gpu_reference_leg(net, feats, cams, dev, ours):
    """The UNMODIFIED reference through stock PyTorch / cuDNN on this B200 (the existing Blackwell path to beat), hot path
    only (features -> depth + confidence), fp32 (TF32 off) and with PyTorch's TF32 switches on; plus the full-size parity
    number: our depth map against the reference's fp32 depth map on the same inputs."""
    ref = load_reference()
    if ref is None:
        return {"unavailable": "oracle/_ref/ref_hotpath.zip not present (python oracle/make_ref.py in the build container)"}
    from oracle import ref_run
    K, R, t, dmin, dmax = [x.to(dev) for x in cams]
    rnet = ref_run.reference_mvsnet(ref, net.state_dict(), "variance", CFG["D"], dev)
    f = [x.to(dev) for x in feats]
    out = {"what": "models/MVSNet/model.py forward from the feature maps on, .cuda(), torch %s, cudnn.benchmark" % torch.__version__}
    for name, tf32 in (("fp32", False), ("tf32", True)):
        try:
            ms, res = ref_run.time_gpu(rnet, f, K, R, t, dmin, dmax, steps=10, warmup=3, tf32=tf32)
        except Exception as e:  # noqa: BLE001
            out[name] = {"error": repr(e)}
            continue
        out[name] = {"ms_per_step": ms, "Mvox_per_s": voxels() / ms / 1e3}
        if ours is not None:
            d = res["depth"]
            out[name]["ours_vs_reference_depth_rel_linf"] = float((ours.to(d.dtype) - d).abs().max() / d.abs().max())
    del rnet
    torch.cuda.empty_cache()
    return out


def kernel_roofline(net, dfeats, dprojs, ddepth, flush, iters=3):
    """Time every kernel of one step with CUDA events on the launching stream and report the dominant one against the
    measured HBM peak."""
    from wild_deep_mvs_b200 import ops
    from wild_deep_mvs_b200 import _lib as L
    reg = net.cost_regularization
    pk = reg._pack()
    _, per_bytes = algorithmic_bytes()
    times = {}

    def timed(name, fn, reps=4):
        # The op writes into a caller-owned buffer (`out=`), so the timed region holds kernel launches only: no
        # allocator traffic.  `reps` back-to-back launches between one pair of events let the host-side cost of a
        # launch (ctypes marshalling) overlap the previous launch instead of being billed to a ~100 us kernel.  L2 is
        # flushed before the first launch; the big layers (K1, conv0, conv11, prob) stream more than L2 holds anyway.
        out = fn(None)
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn(out)
            b.record()
            b.synchronize()
            times.setdefault(name, []).append(a.elapsed_time(b) / reps)
        return out

    am = torch.zeros(16, device=dfeats[0].device)   # abs-max scalars (atomic max only: re-launching into them is harmless)
    amax = {n: am[i:i + 1] for i, n in enumerate(["vol"] + list(pk))}
    cv = lambda x, name, skip=None: (lambda out: ops.conv3d(x, pk[name], skip=skip, out=out, amax=amax[name]))
    warp = ops.mvs_relative_proj(dprojs[0], torch.stack(dprojs[1:], 1))
    vol = timed("k1_cost_volume", lambda out: ops.build_cost_volume(dfeats[0], dfeats[1:], warp, ddepth, CFG["D"], L.GEOM_MVS,
                                                                    L.AGG_VARIANCE, out=out, amax=amax["vol"]))
    c0 = timed("conv0", cv(vol, "conv0"))
    del vol
    c1 = timed("conv1", cv(c0, "conv1"))
    c2 = timed("conv2", cv(c1, "conv2"))
    c3 = timed("conv3", cv(c2, "conv3"))
    c4 = timed("conv4", cv(c3, "conv4"))
    c5 = timed("conv5", cv(c4, "conv5"))
    c6 = timed("conv6", cv(c5, "conv6"))
    x7 = timed("conv7", cv(c6, "conv7", c4))
    x9 = timed("conv9", cv(x7, "conv9", c2))
    x11 = timed("conv11", cv(x9, "conv11", c0))
    s = timed("prob", cv(x11, "prob")).squeeze(-1)
    timed("k3_regress", lambda out: ops.depth_regress(s, ddepth, conf_mode=L.CONF_SUM4))
    peak, how = hbm_peak()
    kernels = []
    for name, ts in times.items():
        ms = sorted(ts)[len(ts) // 2]
        kernels.append({"name": name, "ms": ms, "algorithmic_bytes": per_bytes[name], "GBps": per_bytes[name] / (ms * 1e-3) / 1e9,
                        "frac": per_bytes[name] / (ms * 1e-3) / 1e9 / peak})
    top = max(kernels, key=lambda k: k["ms"])
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("layers", {}).get(top["name"], {}).get("dram_bytes")
    roof = {"kernel": top["name"], "bound": "hbm", "achieved": top["GBps"], "peak": peak, "unit": "GB/s",
            "frac": top["GBps"] / peak, "traffic": traffic, "peak_source": how,
            "share_of_step": top["ms"] / sum(k["ms"] for k in kernels)}
    return roof, kernels


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        reference_arm(a)
    else:
        own_arm(a)
