"""BASELINE cfg5: Vis-MVSNet over a batch of 64 reference views (each 1+4 views, 640x512, 128-interval range, eval
setting depth_nums [64,32,16]) sharded over the GPUs of one box -- contiguous blocks of reference views per rank, no
data-path collective, ONE NCCL all-gather of the depth maps at the end (models/trainer.py:101,246-247).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/bench_cfg5.py [out.json] [--batch G]

`bench.py --gpus N` runs the same leg (run_cfg5) and prints it under the key "cfg5" of its JSON line.

Features are resident in HBM (the hot path: features -> depth; the 2-D extractor is row f1); every rank replays the
three-stage cascade as one CUDA graph per GROUP of G reference views (default 8: the kernels take the views of a group on
their batch axis, so the many small-channel layers of the cascade run on G times the work per launch; --batch 1 = one
replay per view); the last stage's regression kernel writes the depth maps of a group into the rank's slice of the
preallocated gather buffer (shard.DepthGather) and ONE in-place all-gather follows the last group; timed on the device
(CUDA events), max over ranks.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import ops, shard, synth  # noqa: E402
from wild_deep_mvs_b200.vismvsnet import Frontend as Vis  # noqa: E402

N_VIEWS, NUMS, SCALES = 64, [64, 32, 16], [2, 1, 0.5]
VOX = NUMS[0] * 64 * 80 + NUMS[1] * 128 * 160 + NUMS[2] * 256 * 320
# layer-wise compulsory fp32 traffic of one cfg3-eval cascade
ALGORITHMIC_BYTES_PER_VIEW = 5823.0e6


def _inputs(net, dev, seed, G):
    s1 = {k: v.to(dev) for k, v in synth.make_sample(1, 5, 512, 640, seed=seed).items()}
    s = {k: v.expand(G, *v.shape[1:]).contiguous() for k, v in s1.items()}   # a group: G reference views, own inputs
    interval = (s["depth_max"] - s["depth_min"]) / 128
    ref_cam = net.fill_cam_array(s["K"][:, 0], s["R"][:, 0], s["t"][:, 0], s["depth_min"][:, 0], interval[:, 0])
    src_cams = torch.stack([net.fill_cam_array(s["K"][:, i], s["R"][:, i], s["t"][:, i], s["depth_min"][:, i], interval[:, i])
                            for i in range(1, 5)], 1)
    feats1 = [[ops.to_nhwc(f) for f in fv] for fv in ops.map_views(net.model.feat_ext, torch.unbind(s1["imgs"], 1))]
    feats = [[f.expand(G, *f.shape[1:]).contiguous() for f in fv] for fv in feats1]
    return s, interval, ref_cam, src_cams, feats1, feats


def _gain(i):
    return 1.0 + 0.01 * i     # reference view i = the rank's sample with its feature maps scaled by this (checkable)


def run_cfg5(dev, world, rank, G=8, reps=3, hbm_gbs=None):
    """Returns the result dict on rank 0 (None elsewhere).  Must be called by every rank of the process group."""
    torch.manual_seed(0)
    net = Vis()
    synth.randomize_norm_stats(net, seed=2)
    net.depth_nums, net.interval_scales = NUMS, SCALES
    net = net.to(dev).eval()
    a, b = shard.block_partition(N_VIEWS, world, rank)
    n_local = b - a
    assert N_VIEWS % world == 0
    G = max(1, min(G, n_local))
    with torch.no_grad():
        s, interval, ref_cam, src_cams, feats1, feats = _inputs(net, dev, rank, G)
        gather = shard.DepthGather(n_local, (256, 320), dev)
        direct = n_local == G      # one launch group per rank: the regression kernel writes the send slice itself
        g = net.graphed(feats, ref_cam, src_cams, s["depth_min"][:, 0].contiguous(), interval[:, 0].contiguous(), NUMS, SCALES,
                        out_depth=gather.local if direct else None)

        def group_feats(i0, n):
            """Features of reference views i0 .. i0+n-1; a short last group is padded with its first view."""
            gain = torch.tensor([_gain(i0 + min(j, n - 1)) for j in range(G)], device=dev).view(G, 1, 1, 1)
            return [[f * gain for f in fv] for fv in feats]

        def block():
            for i0 in range(a, b, G):
                n = min(G, b - i0)
                ests, _, _ = g(group_feats(i0, n))
                if not direct:
                    gather.local[i0 - a:i0 - a + n].copy_(ests[2][:n])
            return gather.all_gather()

        for _ in range(2):
            out = block()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        for _ in range(reps):
            out = block()
        eb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ea.elapsed_time(eb) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)

        # ---- content of the gathered batch ---------------------------------------------------------------------------
        assert out.shape == (N_VIEWS, 256, 320) and torch.isfinite(out).all()
        # (1) every rank: the checksums of the maps each rank produced, exchanged separately, match its slots everywhere
        sums = out[a:b].double().sum(dim=(1, 2))
        all_sums = torch.empty(N_VIEWS, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_gather_into_tensor(all_sums, sums)
        else:
            all_sums.copy_(sums)
        assert torch.equal(all_sums, out.double().sum(dim=(1, 2))), "gathered maps differ from what their ranks produced"
        # (2) rank 0: a view of ANOTHER rank's block (the last rank's second view), recomputed here on its own from that
        # rank's seeded inputs, equals its slot in the gathered batch; so does view 5 of rank 0's own block.  (The conv
        # engine's power-of-two operand scale is taken per launch group, hence 1e-4 and not bit equality.)
        errs = {}
        if rank == 0:
            last0 = shard.block_partition(N_VIEWS, world, world - 1)[0]
            for view, owner in ((5 % n_local, 0), (last0 + 1 % n_local, world - 1)):
                so, io, rc, sc, f1, _ = _inputs(net, dev, owner, 1)
                one = net.depth_from_features([[f * _gain(view) for f in fv] for fv in f1], rc, sc,
                                              so["depth_min"][:, 0].contiguous(), io[:, 0].contiguous(), NUMS, SCALES)[0][2][0]
                errs["view%d_of_rank%d" % (view, owner)] = ((out[view] - one).abs().max() / one.abs().max()).item()
            assert max(errs.values()) < 1e-4, errs
    if rank != 0:
        return None
    ms = t.item()
    r = {"workload": "BASELINE cfg5: Vis-MVSNet, 64 reference views (1+4 views, 640x512, depth_nums [64,32,16]), contiguous blocks "
                     "of %d views per GPU, ONE in-place all-gather of the 64 depth maps" % n_local,
         "n_gpus": world, "views_per_launch_group": G, "ms_per_64_views": round(ms, 3),
         "depth_maps_per_s": round(N_VIEWS / ms * 1e3, 1), "Mvox_per_s": round(N_VIEWS * VOX / ms / 1e3, 1),
         "gathered_bytes": N_VIEWS * 256 * 320 * 4, "k3_writes_send_slice": bool(direct),
         "gathered_equals_single_gpu_rel_err": errs}
    if hbm_gbs:
        r["frac_of_hbm_roofline"] = round(N_VIEWS * ALGORITHMIC_BYTES_PER_VIEW / (ms * 1e-3) / 1e9 / (world * hbm_gbs), 4)
    return r


def main():
    argv = list(sys.argv[1:])
    G = 8
    if "--batch" in argv:
        i = argv.index("--batch")
        G = int(argv[i + 1])
        del argv[i:i + 2]
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    r = run_cfg5(dev, world, rank, G)
    if rank == 0:
        print(json.dumps(r))
        if argv:
            json.dump(r, open(argv[0], "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
