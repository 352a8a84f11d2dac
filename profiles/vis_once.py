"""One eager pass of the Vis-MVSNet hot path at cfg3 size (for `ncu --metrics gpu__time_duration.sum` launch lists).

    python profiles/vis_once.py [batch of reference views, default 1] [eval: depth_nums [64,32,16] instead of [32,16,8]]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import ops, synth  # noqa: E402
from wild_deep_mvs_b200.vismvsnet import Frontend as Vis  # noqa: E402

DEV = "cuda:0"
torch.manual_seed(0)
net = Vis()
synth.randomize_norm_stats(net, seed=2)
net = net.to(DEV).eval()
NB = int(sys.argv[1]) if len(sys.argv) > 1 else 1
s = {k: v.to(DEV) for k, v in synth.make_sample(NB, 5, 512, 640, seed=0).items()}
nums, scales = ([64, 32, 16], [2, 1, 0.5]) if "eval" in sys.argv[2:] else ([32, 16, 8], [4, 2, 1])
with torch.no_grad():
    interval = ((s["depth_max"] - s["depth_min"]) / 128)
    ref_cam = net.fill_cam_array(s["K"][:, 0], s["R"][:, 0], s["t"][:, 0], s["depth_min"][:, 0], interval[:, 0])
    src_cams = torch.stack([net.fill_cam_array(s["K"][:, i], s["R"][:, i], s["t"][:, i], s["depth_min"][:, i], interval[:, i])
                            for i in range(1, 5)], 1)
    feats = [[ops.to_nhwc(f) for f in fv] for fv in ops.map_views(net.model.feat_ext, torch.unbind(s["imgs"], 1))]
    args = (feats, ref_cam, src_cams, s["depth_min"][:, 0].contiguous(), interval[:, 0].contiguous(), nums, scales)
    for _ in range(2):
        net.depth_from_features(*args)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("vis_hot_path")
    net.depth_from_features(*args)
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
