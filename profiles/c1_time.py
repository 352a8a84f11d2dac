"""C1 head (Cout = 1 conv) timing per volume shape (GPU box).  python profiles/c1_time.py [libmvsb200.so] -- L2 flushed, median of 10."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import _lib as L  # noqa: E402
if len(sys.argv) > 1:
    L.SO_PATH = os.path.abspath(sys.argv[1])
from wild_deep_mvs_b200 import ops  # noqa: E402

dev = "cuda:0"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


torch.manual_seed(0)
for name, cin, dims in (("cfg2 prob", 8, (1, 192, 128, 160)), ("vis s3 pairs", 8, (32, 16, 256, 320)), ("vis s3 fuse", 8, (8, 16, 256, 320)),
                        ("vis s1 pairs", 8, (32, 64, 64, 80)), ("cvp l0 prob0", 16, (1, 8, 1184, 1600)), ("cvp l4 prob0", 16, (1, 96, 74, 100))):
    x = torch.randn(*dims, cin, device=dev)
    w = torch.randn(1, cin, 3, 3, 3, device=dev) / (cin * 27) ** 0.5
    layer = ops.PackedConv(w, None, conv_bias=torch.zeros(1, device=dev))
    y = ops.conv3d(x, layer)
    gb = (x.numel() + y.numel()) * 4 / 1e9
    ms = timed(lambda: ops.conv3d(x, layer, out=y))
    print("%-14s %.4f ms  %.0f GB/s" % (name, ms, gb / ms * 1e3), flush=True)
