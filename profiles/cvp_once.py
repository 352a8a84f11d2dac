"""One eager forward of CVP-MVSNet at cfg4 size (for `ncu --metrics gpu__time_duration.sum` launch lists)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import synth  # noqa: E402
from wild_deep_mvs_b200.cvpmvsnet import Frontend as CVP  # noqa: E402

DEV = "cuda:0"
torch.manual_seed(0)
net = CVP()
synth.randomize_norm_stats(net, seed=3)
net = net.to(DEV).eval()
s = {k: v.to(DEV) for k, v in synth.make_sample(1, 5, 1184, 1600, seed=0).items()}
call = lambda: net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=5)
call()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("cvp_forward")
call()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
