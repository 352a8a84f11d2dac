"""Geometric-consistency filtering of depth maps on the GPU (row f3 of SURVEY.md 8; K8, mvsb200_geometric_filter).

Mirrors the per-view computation of the reference's evaluation/filtering.py:59-84 -- same inputs (the depth map of a
reference view, the depth maps of its source views, K / R / t of all of them with view 0 the reference), same
thresholds (`--depth_threshold`, `--max_reproj_error`, `--min_tri_angle`, `--num_consistent`,
evaluation/pipeline_utils.py:49-52), same three masks -- without the files and the CPU tensors in between.
"""
import ctypes

import torch

from . import _lib as L
from .ops import _dev_f32, _on_operand_device, _ptr, _stream


@_on_operand_device
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


@_on_operand_device
def gathered_masks(ref_depth, gathered, proj_mat, ref_idx, geom_clamping=0.05, imgs=None, want=()):
    """The consumer of the all-gathered depth maps (K9, mvsb200_gathered_masks): the geometric half of the reference's
    `masked_photometricloss` (models/trainer.py:240-278) -- `get_flow_from_depthmap` (:209-219) + the re-projection mask.

    ref_depth [b,h,w]: the depth map this rank computed with view `ref_idx` as its reference; gathered [b,N,h,w]: the
    output of the path's ONE all-gather (`shard.DepthGather.all_gather()` reshaped, or `torch.stack(all_depthmaps, 1)` in
    the reference's terms); proj_mat [b,N,4,4] (`build_proj_matrices`); `geom_clamping` = the reference's
    `--geom_clamping` (train.py:278); imgs [b,N,c,h,w] optional.  Sources are all views but `ref_idx`, ascending.
    Returns dict(masks bool [b,N-1,h,w]) plus what `want` names of: "inside" (bool), "flows" ([b,N-1,h,w,2], the clamped
    normalised grid), "depth_src", "warped_depth" ([b,N-1,h,w]), "warped" ([b,N-1,c,h,w], needs imgs).
    The inverse of the reference view's projection is a 4x4 `torch.inverse` on the device, as in the reference."""
    lib = L.load()
    ref_depth = _dev_f32(ref_depth.contiguous(), "ref_depth")
    gathered = _dev_f32(gathered.contiguous(), "gathered")
    proj_mat = _dev_f32(proj_mat.contiguous(), "proj_mat")
    b, n, h, w = gathered.shape
    if tuple(ref_depth.shape) != (b, h, w) or tuple(proj_mat.shape) != (b, n, 4, 4):
        raise L.Mvsb200Error("gathered_masks: ref_depth %s / proj_mat %s do not match gathered %s"
                             % (tuple(ref_depth.shape), tuple(proj_mat.shape), tuple(gathered.shape)))
    if not 0 <= int(ref_idx) < n:
        raise L.Mvsb200Error("gathered_masks: reference view %d not in [0,%d)" % (ref_idx, n))
    unknown = set(want) - {"inside", "flows", "depth_src", "warped_depth", "warped"}
    if unknown:
        raise L.Mvsb200Error("gathered_masks: unknown outputs %s" % sorted(unknown))
    c = 0
    if imgs is not None:
        imgs = _dev_f32(imgs.contiguous(), "imgs")
        c = imgs.shape[2]
        if tuple(imgs.shape) != (b, n, c, h, w):
            raise L.Mvsb200Error("gathered_masks: imgs %s do not match gathered %s" % (tuple(imgs.shape), tuple(gathered.shape)))
    if "warped" in want and imgs is None:
        raise L.Mvsb200Error("gathered_masks: 'warped' needs imgs")
    inv_ref = torch.inverse(proj_mat[:, ref_idx]).contiguous()
    dev = gathered.device
    new = lambda *shape, dt=torch.float32: torch.empty(shape, device=dev, dtype=dt)
    mask = new(b, n - 1, h, w, dt=torch.uint8)
    inside = new(b, n - 1, h, w, dt=torch.uint8) if "inside" in want else None
    flows = new(b, n - 1, h, w, 2) if "flows" in want else None
    depth_src = new(b, n - 1, h, w) if "depth_src" in want else None
    warped_depth = new(b, n - 1, h, w) if "warped_depth" in want else None
    warped = new(b, n - 1, c, h, w) if "warped" in want else None
    L.check(lib.mvsb200_gathered_masks(b, n, c, h, w, int(ref_idx), _ptr(ref_depth), _ptr(gathered), _ptr(proj_mat), _ptr(inv_ref),
                                       _ptr(imgs), ctypes.c_float(geom_clamping), _ptr(mask), _ptr(inside), _ptr(flows),
                                       _ptr(depth_src), _ptr(warped_depth), _ptr(warped), _stream()), "mvsb200_gathered_masks")
    out = {"masks": mask.bool()}
    for name, t in (("inside", inside), ("flows", flows), ("depth_src", depth_src), ("warped_depth", warped_depth), ("warped", warped)):
        if t is not None:
            out[name] = t.bool() if t.dtype == torch.uint8 else t
    return out
