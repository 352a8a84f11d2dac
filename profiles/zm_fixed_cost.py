"""Fixed cost of one z-march launch: tiny volumes (one or two tile columns, a few planes) timed back to back.

    python profiles/zm_fixed_cost.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import ops  # noqa: E402

dev = "cuda:0"
torch.manual_seed(0)
CASES = [("16->16 s1, 2 planes, 1 column", 16, 16, 1, False, (1, 2, 7, 16)),
         ("16->16 s1, 8 planes, 1 column", 16, 16, 1, False, (1, 8, 7, 16)),
         ("16->16 s1, 32 planes, 1 column", 16, 16, 1, False, (1, 32, 7, 16)),
         ("16->16 s1, 32 planes, 148 columns", 16, 16, 1, False, (1, 32, 7 * 4, 16 * 37)),
         ("32->64 s1 (conv6 half), cfg2 size", 32, 64, 1, False, (1, 24, 16, 20)),
         ("32->64 s2 (conv5), cfg2 size", 32, 64, 2, False, (1, 48, 32, 40)),
         ("64->32 deconv (conv7), cfg2 size", 64, 32, 2, True, (1, 24, 16, 20))]
for name, cin, cout, stride, tr, dims in CASES:
    x = torch.randn(*dims, cin, device=dev)
    w = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3), device=dev) / (cin * 27) ** 0.5
    layer = ops.PackedConv(w, None, stride=stride, transposed=tr, relu=True)
    y = ops.conv3d(x, layer, engine="zm")
    for _ in range(3):
        ops.conv3d(x, layer, engine="zm", out=y)
    torch.cuda.synchronize()
    res = []
    for reps in (1, 16):
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                ops.conv3d(x, layer, engine="zm", out=y)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) / reps * 1e3)
        res.append(sorted(ts)[2])
    # GPU-side cost without the host: 16 launches captured in one CUDA graph
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.conv3d(x, layer, engine="zm", out=y)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(16):
            ops.conv3d(x, layer, engine="zm", out=y)
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) / 16 * 1e3)
    print("%-40s single launch %.1f us, 16 back to back %.1f us each, 16 in one CUDA graph %.1f us each" % (name, res[0], res[1], sorted(ts)[2]))
