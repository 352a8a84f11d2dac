"""BASELINE cfg5: Vis-MVSNet over a batch of 64 reference views (each 1+4 views, 640x512, 128-interval range, eval
setting depth_nums [64,32,16]) sharded over the GPUs of one box -- contiguous blocks of reference views per rank, no
data-path collective, ONE NCCL all-gather of the depth maps at the end (models/trainer.py:101,246-247).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/bench_cfg5.py [out.json] [--batch G]

Features are resident in HBM (the hot path: features -> depth; the 2-D extractor is row f1); every rank replays the
three-stage cascade as one CUDA graph per GROUP of G reference views (default 8: the kernels take the views of a group on
their batch axis, so the many small-channel layers of the cascade run on G times the work per launch; --batch 1 = one
replay per view); timed on the device, max over ranks.
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
    torch.manual_seed(0)
    net = Vis()
    synth.randomize_norm_stats(net, seed=2)
    net.depth_nums, net.interval_scales = NUMS, SCALES
    net = net.to(dev).eval()
    a, b = shard.block_partition(N_VIEWS, world, rank)
    G = max(1, min(G, b - a))
    with torch.no_grad():
        s1 = {k: v.to(dev) for k, v in synth.make_sample(1, 5, 512, 640, seed=rank).items()}
        s = {k: v.expand(G, *v.shape[1:]).contiguous() for k, v in s1.items()}   # a group: G reference views, own inputs
        interval = (s["depth_max"] - s["depth_min"]) / 128
        ref_cam = net.fill_cam_array(s["K"][:, 0], s["R"][:, 0], s["t"][:, 0], s["depth_min"][:, 0], interval[:, 0])
        src_cams = torch.stack([net.fill_cam_array(s["K"][:, i], s["R"][:, i], s["t"][:, i], s["depth_min"][:, i], interval[:, i])
                                for i in range(1, 5)], 1)
        feats = [[ops.to_nhwc(f) for f in fv] for fv in ops.map_views(net.model.feat_ext, torch.unbind(s["imgs"], 1))]
        g = net.graphed(feats, ref_cam, src_cams, s["depth_min"][:, 0].contiguous(), interval[:, 0].contiguous(), NUMS, SCALES)
        # every reference view of the block: its own features (here: the sample's maps scaled per view, so that the
        # gathered result can be checked), copied into the graph's buffers, one replay, depth map kept
        maps = torch.empty(b - a, 256, 320, device=dev)
        g1 = g if G == 1 else None

        def group_feats(i0, n):
            """Features of reference views i0 .. i0+n-1 (the sample's maps scaled per view, so results can be checked);
            a short last group is padded with its first view."""
            gain = torch.tensor([1.0 + 0.01 * (i0 + min(j, n - 1)) for j in range(G)], device=dev).view(G, 1, 1, 1)
            return [[f * gain for f in fv] for fv in feats]

        def block():
            for i0 in range(a, b, G):
                n = min(G, b - i0)
                ests, _, _ = g(group_feats(i0, n))
                maps[i0 - a:i0 - a + n].copy_(ests[2][:n])
            return shard.gather_depth_maps(maps, N_VIEWS)

        for _ in range(2):
            out = block()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        ea.record()
        for _ in range(reps):
            out = block()
        eb.record()
        torch.cuda.synchronize()
        t = torch.tensor([ea.elapsed_time(eb) / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # every rank holds all 64 maps, ordered by global view index; view i of any rank equals a single-GPU run of view i
        assert out.shape == (N_VIEWS, 256, 320) and torch.isfinite(out).all()
        same_inputs = world == 1 or rank == 0   # ranks use different samples (seed = rank): only rank 0 owns view 5's inputs
        if same_inputs:
            # view 5 computed on its own (batch of one) against its slot in the gathered batch: the per-sample arithmetic
            # is the same; the conv engine's power-of-two operand scale comes from the abs-max of the whole group
            f1 = [[ops.to_nhwc(f) for f in fv] for fv in ops.map_views(net.model.feat_ext, torch.unbind(s1["imgs"], 1))]
            one = net.depth_from_features([[f * (1.0 + 0.01 * 5) for f in fv] for fv in f1], ref_cam[:1], src_cams[:1],
                                          s1["depth_min"][:, 0].contiguous(), interval[:1, 0].contiguous(), NUMS, SCALES)[0][2][0]
            err = ((out[5] - one).abs().max() / one.abs().max()).item()
            assert err < 1e-4, err
    vox = NUMS[0] * 64 * 80 + NUMS[1] * 128 * 160 + NUMS[2] * 256 * 320
    if rank == 0:
        ms = t.item()
        r = {"config": "cfg5 Vis-MVSNet, 64 reference views (1+4 views, 640x512, depth_nums [64,32,16]), view-sharded + 1 all-gather",
             "n_gpus": world, "views_per_launch_group": G, "ms_per_batch_of_64": round(ms, 2), "depth_maps_per_s": round(N_VIEWS / ms * 1e3, 1),
             "hot_path_Mvox_per_s": round(N_VIEWS * vox / ms / 1e3, 1), "gathered_bytes": N_VIEWS * 256 * 320 * 4}
        print(json.dumps(r))
        if argv:
            json.dump(r, open(argv[0], "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
