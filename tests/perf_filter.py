"""K8 timing at DTU size: one 1600x1184 reference depth map against 10 source maps (the reconstruction pipeline's
per-view filtering step, evaluation/filtering.py), device resident, CUDA events; and the numpy oracle on the host for
scale (the reference does the same arithmetic with CPU torch ops)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.filter import geometric_filter as oracle_filter  # noqa: E402
from wild_deep_mvs_b200.filtering import geometric_filter  # noqa: E402

h, w, N = 1184, 1600, 10
rng = np.random.default_rng(0)
K = np.tile(np.array([[2892.33, 0, 800.0], [0, 2883.18, 592.0], [0, 0, 1]], np.float32), (N + 1, 1, 1))
R = np.tile(np.eye(3, dtype=np.float32), (N + 1, 1, 1))
t = np.zeros((N + 1, 3, 1), np.float32)
t[:, 0, 0] = -40.0 * np.arange(N + 1)
depths = [(650 + 20 * rng.standard_normal((h, w))).astype(np.float32) for _ in range(N + 1)]
dev = "cuda:0"
cu = lambda a: torch.as_tensor(a, device=dev)
args = (cu(depths[0]), [cu(d) for d in depths[1:]], cu(K), cu(R), cu(t))
for _ in range(3):
    geometric_filter(*args)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    geometric_filter(*args)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
t0 = time.perf_counter()
oracle_filter(depths[0][:296], [d[:296] for d in depths[1:]], K, R, t)   # a quarter of the rows
cpu_s = (time.perf_counter() - t0) * 4
print(json.dumps({"kernel": "k8_geo_filter", "image": [h, w], "sources": N, "ms": round(ms, 3),
                  "Mpix_sources_per_s": round(h * w * N / ms / 1e3, 1), "numpy_oracle_s_estimated_full": round(cpu_s, 2)}))
