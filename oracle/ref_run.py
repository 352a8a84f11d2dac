"""Drive the UNMODIFIED reference (oracle/ref_import.py) on the hot path: features in -> depth out.

TEST / MEASUREMENT INFRASTRUCTURE ONLY: used by bench.py's reference legs (`--impl reference`, `cpu_baseline`,
`gpu_reference`) and by tests/.  The product never imports this module.

The hot path of SURVEY.md section 8 starts at the feature maps.  The reference has no entry point there, so its own
`forward` (models/MVSNet/model.py:178-218) is called with the 2-D extractor `net.feature` replaced by a lookup that hands
back the pre-computed maps in call order: everything from `build_cost_volume` on -- homo_warping, the variance / softmin
aggregation, CostRegNet, softmax, depth regression, photometric confidence -- is the reference's stock code path.
"""
import time

import torch
import torch.nn as nn


class FeatureLookup(nn.Module):
    """Stands in for MVSNet.feature: returns the stored feature maps, one per call, in view order."""

    def __init__(self, feats):
        super().__init__()
        self.feats, self.next = list(feats), 0

    def forward(self, img):
        f = self.feats[self.next % len(self.feats)]
        self.next += 1
        return f


def reference_mvsnet(ref, state_dict, aggregation, num_depth, device="cpu"):
    """The reference's MVSNet with OUR synthetic weights (same key names: the drop-in keeps the reference's state_dict
    layout, so `strict=True` loads)."""
    net = ref.MVSNet(aggregation)
    net.load_state_dict({k: v.detach().cpu() for k, v in state_dict.items()}, strict=True)
    net.num_depth = num_depth
    return net.to(device).eval()


def mvsnet_from_features(net, feats, K, R, t, depth_min, depth_max):
    """feats: list of V maps [B,C,h,w] (reference layout); K is the FULL-resolution intrinsics (forward divides by 4).
    Returns the reference's output dict."""
    saved = net.feature
    net.feature = FeatureLookup(feats)
    try:
        B, V = K.shape[:2]
        imgs = feats[0].new_zeros(B, V, 1, 1, 1)     # only unbound into V placeholders; the lookup ignores them
        with torch.no_grad():
            return net(imgs, K, R, t, depth_min, depth_max)
    finally:
        net.feature = saved


def time_cpu(net, feats, K, R, t, depth_min, depth_max, steps, warmup):
    """Wall-clock seconds per pass of mvsnet_from_features on the host (mean over `steps` after `warmup`)."""
    for _ in range(warmup):
        mvsnet_from_features(net, feats, K, R, t, depth_min, depth_max)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        mvsnet_from_features(net, feats, K, R, t, depth_min, depth_max)
        ts.append(time.perf_counter() - t0)
    return sum(ts) / len(ts), ts


def time_gpu(net, feats, K, R, t, depth_min, depth_max, steps, warmup, tf32):
    """CUDA-event milliseconds per pass on the current device through stock PyTorch / cuDNN (median over `steps`)."""
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(tf32)
    torch.backends.cudnn.benchmark = True           # let cuDNN pick its best algorithm per layer shape
    try:
        for _ in range(warmup):
            out = mvsnet_from_features(net, feats, K, R, t, depth_min, depth_max)
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = mvsnet_from_features(net, feats, K, R, t, depth_min, depth_max)
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b))
        return sorted(ts)[len(ts) // 2], out
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
