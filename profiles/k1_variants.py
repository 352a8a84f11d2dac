"""K1 experiments at the cfg2 size (debug aid, GPU box): times the kernel under the MVSB200_K1_DEBUG masks
(1 no stores, 2 no re-loads on a cell change, 4 no tap loads at all).  The masks exist only in an experiment build:

    MVSB200_NVCC_EXTRA=-DMVSB200_K1_EXPERIMENTS python -m wild_deep_mvs_b200.build --force && python profiles/k1_variants.py

r3 result (B200, cfg2 volume, 256-bit loads/stores): default 0.322 ms, no stores 0.304, no re-loads 0.290, no loads 0.271,
neither 0.263 -- the kernel is bound by its own instruction stream (geometry, shuffles, packed FMAs at 16 warps / SM),
not by HBM or L2."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import _lib as L, ops, synth  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
import bench  # noqa: E402

net, feats, projs, depth = bench.make_workload(seed=0)
C, H, W, D, B = bench.CFG["C"], bench.CFG["h"], bench.CFG["w"], bench.CFG["D"], 1
dfeats = [ops.to_nhwc(f.to(dev)) for f in feats]
dprojs = list(torch.unbind(projs.to(dev), 1))
depth = depth.to(dev)
ref, srcs = dfeats[0], dfeats[1:]
warp = ops.mvs_relative_proj(dprojs[0], torch.stack(dprojs[1:], 1))


def run(tag, env):
    os.environ.pop("MVSB200_K1_DEBUG", None)
    os.environ.update(env)
    out = torch.empty(B, D, H, W, C, device=dev)
    amax = torch.zeros(1, device=dev)
    for _ in range(3):
        ops.build_cost_volume(ref, srcs, warp, depth, D, L.GEOM_MVS, L.AGG_VARIANCE, out=out, amax=amax)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.build_cost_volume(ref, srcs, warp, depth, D, L.GEOM_MVS, L.AGG_VARIANCE, out=out, amax=amax)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    print("%-28s %.4f ms (min %.4f)" % (tag, ts[len(ts) // 2], ts[0]), flush=True)


run("default", {})
run("no stores", {"MVSB200_K1_DEBUG": "1"})
run("no re-loads", {"MVSB200_K1_DEBUG": "2"})
run("no loads", {"MVSB200_K1_DEBUG": "4"})
run("no loads, no stores", {"MVSB200_K1_DEBUG": "5"})
