"""Seeded synthetic inputs for tests, goldens and bench.py.

DTU-like pinhole cameras on a small arc: view v is rotated about y by 0.05*v rad and
translated by (-40*v, 0, 0) mm; v=0 is the reference view.  Depth range 425..905 mm
(data/dtu_yao.py:109 of the reference).  No CUDA needed to import this module.
"""
import math

import torch


def make_cameras(batch, views, img_h, img_w, dtype=torch.float32):
    K = torch.zeros(batch, views, 3, 3, dtype=dtype)
    R = torch.zeros(batch, views, 3, 3, dtype=dtype)
    t = torch.zeros(batch, views, 3, 1, dtype=dtype)
    for v in range(views):
        K[:, v] = torch.tensor([[2892.33 * img_w / 1600.0, 0.0, img_w / 2.0],
                                [0.0, 2883.18 * img_h / 1200.0, img_h / 2.0],
                                [0.0, 0.0, 1.0]], dtype=dtype)
        a = 0.05 * v
        R[:, v] = torch.tensor([[math.cos(a), 0.0, math.sin(a)],
                                [0.0, 1.0, 0.0],
                                [-math.sin(a), 0.0, math.cos(a)]], dtype=dtype)
        t[:, v] = torch.tensor([[-40.0 * v], [0.0], [0.0]], dtype=dtype)
    depth_min = torch.full((batch, views), 425.0, dtype=dtype)
    depth_max = torch.full((batch, views), 425.0 + 2.5 * 192, dtype=dtype)
    return K, R, t, depth_min, depth_max


def make_sample(batch, views, img_h, img_w, seed=0):
    """The sample dict the reference data loaders produce (data/MVSDataset.py)."""
    g = torch.Generator().manual_seed(seed)
    imgs = torch.rand(batch, views, 3, img_h, img_w, generator=g)
    K, R, t, dmin, dmax = make_cameras(batch, views, img_h, img_w)
    return {"imgs": imgs, "K": K, "R": R, "t": t, "depth_min": dmin, "depth_max": dmax}


def make_features(batch, views, channels, h, w, seed=0):
    """Kernel-level inputs: N(0,1) feature maps, list of [B,C,h,w]."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(batch, channels, h, w, generator=g) for _ in range(views)]


def randomize_norm_stats(module, seed=0):
    """Give every BatchNorm non-trivial running stats / affine so folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm2d, torch.nn.BatchNorm3d)):
            with torch.no_grad():
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)


def scale_param(param, gain):
    """Multiply a head conv's weight so the softmax over depth is peaked."""
    with torch.no_grad():
        param.mul_(gain)
