"""K3 timing per volume shape (GPU box).  python profiles/k3_time.py [libmvsb200.so]  -- L2 flushed before each launch, median of 10."""
import sys, torch
sys.path.insert(0, ".")
from wild_deep_mvs_b200 import _lib as L
if len(sys.argv) > 1:
    import os
    L.SO_PATH = os.path.abspath(sys.argv[1])
from wild_deep_mvs_b200 import ops
dev = "cuda:0"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
def timed(fn):
    for _ in range(3): fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
for name, B, D, H, W in (("vis s3 pairs", 32, 16, 256, 320), ("vis s2 pairs", 32, 32, 128, 160), ("vis s1 pairs", 32, 64, 64, 80), ("vis s3 fuse", 8, 16, 256, 320),
                         ("cvp l0", 1, 8, 1184, 1600), ("cfg1", 1, 48, 128, 160), ("cfg2", 1, 192, 128, 160)):
    score = torch.randn(B, D, H, W, device=dev)
    start = torch.rand(B, H, W, device=dev) + 400
    interval = torch.ones(B, device=dev)
    fn = lambda: ops.depth_regress(score, start, interval=interval, conf_mode=L.CONF_NONE, want_entropy=True)
    try:
        print("%-14s %.4f ms" % (name, timed(fn)), flush=True)
    except Exception as e:
        print(name, "FAILED", repr(e)[:200])
# the bench's call (MVSNet: scalar hypotheses per batch, 4-bin confidence, preallocated outputs), L2 flushed and L2 hot
score = torch.randn(1, 192, 128, 160, device=dev)
values = torch.linspace(425.0, 935.0, 192, device=dev).view(1, 192)
od, oc = torch.empty(1, 128, 160, device=dev), torch.empty(1, 128, 160, device=dev)
fn = lambda: ops.depth_regress(score, values, conf_mode=L.CONF_SUM4, out_depth=od, out_conf=oc)
print("%-14s %.4f ms" % ("cfg2 bench", timed(fn)), flush=True)
fn(); torch.cuda.synchronize()
ts = []
for _ in range(20):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("%-14s %.4f ms" % ("cfg2 bench hot", sorted(ts)[10]), flush=True)
