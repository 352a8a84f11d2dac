"""In-kernel cycle accounting of the z-march conv engine on the cfg2 layers (debug aid; run on the GPU box).

    python profiles/zm_profile.py            # prints, per layer, where the MMA / producer / epilogue warps of CTA 0 wait
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import _lib as L, ops  # noqa: E402

lib = L.load()
lib.mvsb200_debug_zm_profile.restype = ctypes.c_int
lib.mvsb200_debug_zm_profile.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_ulonglong)]
dev = "cuda:0"
torch.manual_seed(0)
LAYERS = [("conv0", 32, 8, 1, False, (1, 192, 128, 160)), ("conv1", 8, 16, 2, False, (1, 192, 128, 160)),
          ("conv2", 16, 16, 1, False, (1, 96, 64, 80)), ("conv3", 16, 32, 2, False, (1, 96, 64, 80)),
          ("conv4", 32, 32, 1, False, (1, 48, 32, 40)), ("conv5", 32, 64, 2, False, (1, 48, 32, 40)),
          ("conv6h", 32, 64, 1, False, (1, 24, 16, 20)), ("conv7", 64, 32, 2, True, (1, 24, 16, 20)),
          ("conv9", 32, 16, 2, True, (1, 48, 32, 40)),
          ("tiny", 16, 16, 1, False, (1, 2, 7, 16)),   # one tile, two planes: the fixed cost of a launch (prologue + one pipeline pass)
          ("conv11", 16, 8, 2, True, (1, 96, 64, 80)),
          # Vis-MVSNet Reg blocks: 4 source views on the batch axis, stage 3 / stage 1 volumes
          ("vis8x8s3", 8, 8, 1, False, (4, 8, 256, 320)), ("vis8x8s1", 8, 8, 1, False, (4, 32, 64, 80)),
          ("vis8x16s3", 8, 16, 2, False, (4, 8, 256, 320)), ("visup-s3", 16, 8, 2, True, (4, 4, 128, 160)),
          ("vis16x8s3", 16, 8, 1, False, (4, 8, 256, 320))]
if len(sys.argv) > 1:
    LAYERS = [l for l in LAYERS if any(a in l[0] for a in sys.argv[1:])]
for name, cin, cout, stride, tr, dims in LAYERS:
    x = torch.randn(*dims, cin, device=dev)
    w = torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3), device=dev) / (cin * 27) ** 0.5
    layer = ops.PackedConv(w, None, stride=stride, transposed=tr, relu=True)
    y = ops.conv3d(x, layer, engine="zm")
    torch.cuda.synchronize()
    lib.mvsb200_debug_zm_profile(1, None)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.conv3d(x, layer, engine="zm", out=y)
    b.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 32)()
    lib.mvsb200_debug_zm_profile(0, buf)
    v = list(buf)
    pct = lambda n, d: 100.0 * n / max(d, 1)
    print("%-7s %.3f ms | mma: total %d cyc, %d stages (%.0f cyc/stage): wait-full %.0f%% wait-acc %.0f%% issue %.0f%% | "
          "producer: %d units (%.0f cyc/unit/group) wait-empty %.0f%% | epilogue: %d planes (%.0f cyc/plane) wait-acc-full %.0f%% barriers %.0f%%"
          % (name, a.elapsed_time(b), v[0], v[4], v[0] / max(v[4], 1), pct(v[1], v[0]), pct(v[2], v[0]), pct(v[3], v[0]),
             v[10], v[8] / max(v[10], 1), pct(v[9], v[8]), v[19], v[16] / max(v[19], 1), pct(v[17], v[16]), pct(v[18], v[16])))
