"""Geometric-consistency filtering of depth maps on the GPU (row f3 of SURVEY.md 8; K8, mvsb200_geometric_filter).

Mirrors the per-view computation of the reference's evaluation/filtering.py:59-84 -- same inputs (the depth map of a
reference view, the depth maps of its source views, K / R / t of all of them with view 0 the reference), same
thresholds (`--depth_threshold`, `--max_reproj_error`, `--min_tri_angle`, `--num_consistent`,
evaluation/pipeline_utils.py:49-52), same three masks -- without the files and the CPU tensors in between.
"""
import ctypes

import torch

from . import _lib as L
from .ops import _dev_f32, _ptr, _stream


def geometric_filter(depth, src_depths, K, R, t, depth_threshold=0.01, max_reproj_error=1.0, min_tri_angle=1.0,
                     num_consistent=3, want_votes=False):
    """depth [h,w]; src_depths: list of N maps [hi,wi] (sizes may differ); K, R [1+N,3,3]; t [1+N,3,1] or [1+N,3];
    all CUDA float32.  Returns dict(mask_depth, mask_disp, geo_mask) of bool [h,w] (+ votes uint8 [3,h,w])."""
    lib = L.load()
    depth = _dev_f32(depth.contiguous(), "depth")
    h, w = depth.shape
    n = len(src_depths)
    if not 1 <= n <= L.MAX_SRC:
        raise L.Mvsb200Error("number of source views %d not in [1,%d]" % (n, L.MAX_SRC))
    if K.shape[0] != n + 1 or R.shape[0] != n + 1 or t.shape[0] != n + 1:
        raise L.Mvsb200Error("K, R, t must hold 1 + %d views" % n)
    srcs = [_dev_f32(s.contiguous(), "src_depths[%d]" % i) for i, s in enumerate(src_depths)]
    ptrs = (ctypes.c_void_p * n)(*[s.data_ptr() for s in srcs])
    hs = (ctypes.c_int * n)(*[s.shape[0] for s in srcs])
    ws = (ctypes.c_int * n)(*[s.shape[1] for s in srcs])
    Kc, Rc = _dev_f32(K.contiguous(), "K"), _dev_f32(R.contiguous(), "R")
    tc = _dev_f32(t.reshape(n + 1, 3).contiguous(), "t")
    masks = torch.empty(3, h, w, device=depth.device, dtype=torch.uint8)
    votes = torch.empty(3, h, w, device=depth.device, dtype=torch.uint8) if want_votes else None
    L.check(lib.mvsb200_geometric_filter(_ptr(depth), h, w, ptrs, hs, ws, n, _ptr(Kc), _ptr(Rc), _ptr(tc),
                                         ctypes.c_float(depth_threshold), ctypes.c_float(max_reproj_error),
                                         ctypes.c_float(min_tri_angle), num_consistent, _ptr(masks[0]), _ptr(masks[1]),
                                         _ptr(masks[2]), _ptr(votes), _stream()), "mvsb200_geometric_filter")
    out = {"mask_depth": masks[0].bool(), "mask_disp": masks[1].bool(), "geo_mask": masks[2].bool()}
    if want_votes:
        out["votes"] = votes
    return out
