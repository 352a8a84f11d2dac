"""The callers either side of the depth path, on the reference's own directory layout (rows f3 / f4 of SURVEY.md 8):

    <data_path>/IntRes/depthmaps/<model>_<nviews>/<scene>/<filename>_out.npz            run_depthmaps   (writes)
    <data_path>/IntRes/geometric_filtering/<model>_<nviews>/<scene>/<filename>_out.npz  run_filtering   (writes)
    <point_folder>/{cams/<f>.jpg.P, images/<f>.jpg, 2333__<f>/{disp,normals}.dmb}       mvsnet_to_gipuma (writes)

Same `args` attributes, batch keys, file names, skip / `finished.txt` behaviour as evaluation/run_depthmaps.py:27-75,
evaluation/filtering.py:25-91, evaluation/pipeline_utils.py:83-110 and evaluation/fusibile.py:96-158 -- with the
network `forward` being the libmvsb200 drop-in and the geometric filter running as one kernel (K8) on the GPU instead of
~25 CPU tensor ops per view.  Checkpoint discovery (`load_network`), datasets and the fusibile binary itself stay out of
scope: the caller passes the network and a dataloader-like iterable of batches.
"""
import os
from pathlib import Path

import numpy as np
import torch
from torch.nn import functional as F

from . import formats
from .filtering import geometric_filter


def depth_folder_name(args):
    """evaluation/pipeline_utils.py:83-85."""
    return "%s_%s" % (args.model, args.nviews)


def _depth_dir(args):
    return Path(args.data_path) / "IntRes" / "depthmaps" / depth_folder_name(args) / str(args.scene)


def _filter_dir(args):
    return Path(args.data_path) / "IntRes" / "geometric_filtering" / depth_folder_name(args) / str(args.scene)


def run_depthmaps(dataloader, args, net, device=None):
    """evaluation/run_depthmaps.py:27-75 with the network given (no checkpoint lookup).  Each batch holds `filename`
    (list of B names), `imgs`, `K`, `R`, `t`, `depth_min`, `depth_max`; one `<filename>_out.npz` per view."""
    out = _depth_dir(args)
    out.mkdir(parents=True, exist_ok=True)
    if (out / "finished.txt").exists() and not getattr(args, "override", False):
        return
    if device is None:
        p = next(iter(net.parameters()), None)
        device = p.device if p is not None else torch.device("cpu")
    to_dev = lambda v: [to_dev(u) for u in v] if isinstance(v, (list, tuple)) else v.to(device)
    for sample in dataloader:
        names = sample["filename"]
        if all((out / ("%s_out.npz" % f)).exists() for f in names) and not getattr(args, "override", False):
            continue
        with torch.no_grad():
            res = net(to_dev(sample["imgs"]), to_dev(sample["K"]), to_dev(sample["R"]), to_dev(sample["t"]),
                      to_dev(sample["depth_min"]), to_dev(sample["depth_max"]))
        for f, depth, conf in zip(names, res["depth"], res["photometric_confidence"]):
            formats.save_depth_npz(out / ("%s_out.npz" % f), depth, conf)
        if getattr(args, "debug", False):
            return
    with open(out / "finished.txt", "a") as f:
        f.write(" ")


def run_filtering(dataloader, args, device="cuda"):
    """evaluation/filtering.py:25-91: for every reference view, load its depth map and those of its source views, vote
    the three consistency tests over the sources (K8) and store `mask_depth`, `mask_disp`, `geo_mask`.
    Batches hold `filename` ([name]), `K`, `R` [1,1+N,3,3], `t` [1,1+N,3,1], `src_filenames` (N lists of one name)."""
    out = _filter_dir(args)
    if (out / "finished.txt").exists():
        return
    out.mkdir(parents=True, exist_ok=True)
    depth_dir = _depth_dir(args)
    for batch in dataloader:
        name = batch["filename"][0]
        K = batch["K"][0].to(device=device, dtype=torch.float32).clone()
        R = batch["R"][0].to(device=device, dtype=torch.float32)
        t = batch["t"][0].to(device=device, dtype=torch.float32)
        load = lambda f: torch.from_numpy(formats.load_depth_npz(depth_dir / ("%s_out.npz" % f))[0]).to(device)
        depth = load(name)
        srcs = [load(f[0]) for f in batch["src_filenames"]]
        K[:, :2] /= 1 if args.upsample else args.downscale
        if args.upsample:   # nearest-neighbour, as F.interpolate defaults to (filtering.py:53-57)
            up = lambda d: F.interpolate(d[None, None], scale_factor=args.downscale).squeeze()
            depth, srcs = up(depth), [up(d) for d in srcs]
        m = geometric_filter(depth, srcs, K, R, t, args.depth_threshold, args.max_reproj_error, args.min_tri_angle,
                             args.num_consistent)
        np.savez_compressed(out / ("%s_out.npz" % name), mask_depth=m["mask_depth"].cpu().numpy(),
                            mask_disp=m["mask_disp"].cpu().numpy(), geo_mask=m["geo_mask"].cpu().numpy())
        if getattr(args, "debug", False):
            return
    with open(out / "finished.txt", "a") as f:
        f.write(" ")


def get_mask(args, filename, **kwargs):
    """evaluation/pipeline_utils.py:88-110: pixels to DROP -- confidence below `prob_threshold` (in every channel for
    multi-channel confidences) or, with `args.filter`, not geometrically consistent."""
    if "prob" not in kwargs:
        raise NotImplementedError("Need a probability mask from get_mask")
    prob = kwargs["prob"]
    invalid = (prob < args.prob_threshold).all(axis=0) if prob.ndim > 2 else prob < args.prob_threshold
    if args.filter:
        geo = kwargs["geo_mask"] if "geo_mask" in kwargs else np.load(_filter_dir(args) / ("%s_out.npz" % filename))["geo_mask"]
        invalid = invalid | ~geo
    return invalid


def mvsnet_to_gipuma(args, gipuma_point_folder, dataloader):
    """evaluation/fusibile.py:96-158: cameras (`.P`), down-scaled images and masked depth / fake-normal `.dmb` files
    fusibile consumes.  Batches hold `filename`, `imgs` [1,V,3,h,w], `K`, `R`, `t` (+ optional `degenerate`)."""
    folder = Path(gipuma_point_folder)
    (folder / "cams").mkdir(parents=True, exist_ok=True)
    (folder / "images").mkdir(parents=True, exist_ok=True)
    depth_dir = _depth_dir(args)
    for batch in dataloader:
        name = batch["filename"][0]
        img = batch["imgs"][0, 0]
        h, w = img.shape[1:]
        P = formats.gipuma_projection(batch["K"][0, 0].numpy(), batch["R"][0, 0].numpy(), batch["t"][0, 0].numpy(),
                                      args.downscale)
        formats.write_gipuma_cam(P, folder / "cams" / ("%s.jpg.P" % name))
        try:
            from PIL import Image
            arr = (img.mul(255).byte() if img.is_floating_point() else img).permute(1, 2, 0).cpu().numpy()
            Image.fromarray(arr).resize((w // args.downscale, h // args.downscale), resample=Image.LANCZOS).save(
                folder / "images" / ("%s.jpg" % name))
        except ImportError:   # the images only colour the fused points
            pass
        sub = folder / ("2333__" + name)
        sub.mkdir(exist_ok=True)
        depth, prob = formats.load_depth_npz(depth_dir / ("%s_out.npz" % name))
        if getattr(args, "colmap", False):
            depth, prob = depth[:h, :w], prob[..., :h, :w] if prob.ndim > 2 else prob[:h, :w]
        invalid = np.ones(depth.shape, bool) if batch.get("degenerate", False) else get_mask(args, name, prob=prob)
        formats.export_gipuma_view(os.fspath(sub), depth, invalid)
