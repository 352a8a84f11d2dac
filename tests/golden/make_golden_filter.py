"""Golden masks for the geometric-consistency filter (evaluation/filtering.py:25-91).

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_filter.py

Executes the UNMODIFIED `evaluation.filtering.run` of the reference on a synthetic scene written to a temporary
directory in the reference's own on-disk format (`<filename>_out.npz` with `depthmap`), and stores inputs and the three
masks it writes (tests/golden/geo_filter.npz).  Scene: a slanted plane seen by 1 + 4 pinhole cameras, depth maps rendered
analytically; one region of the reference map and one of a source map are corrupted, one source has a 3 mm baseline (its
triangulation angle fails), one source map has a different size (the reference handles per-source shapes).
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def plane_depth(K, R, t, h, w, n, c):
    """Depth map of the world plane n.X = c for camera x_cam = R X + t."""
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    pix = np.stack([xs, ys, np.ones_like(xs)], -1) @ np.linalg.inv(K).T          # camera-frame rays with z = 1
    d = pix @ R                                                                   # world directions (R^T ray)
    o = -R.T @ t.reshape(3)
    return ((c - n @ o) / (d @ n)).astype(np.float32)


def scene(seed=0):
    rng = np.random.default_rng(seed)
    h, w, V = 48, 64, 5
    K = np.zeros((V, 3, 3), np.float64)
    R = np.zeros((V, 3, 3), np.float64)
    t = np.zeros((V, 3, 1), np.float64)
    base = [0.0, -40.0, 35.0, -3.0, 60.0]                    # view 3: 3 mm baseline -> triangulation angle < 1 degree
    sizes = [(h, w), (h, w), (h, w), (h, w), (40, 56)]       # view 4: a smaller depth map
    for v in range(V):
        hv, wv = sizes[v]
        K[v] = [[520.0 * wv / w, 0, wv / 2.0], [0, 515.0 * hv / h, hv / 2.0], [0, 0, 1]]
        a = 0.04 * v * (1 if v % 2 else -1)
        R[v] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        t[v] = [[base[v]], [2.0 * v], [0.0]]
    n, c = np.array([0.15, -0.1, 1.0]), 600.0
    depths = [plane_depth(K[v], R[v], t[v], sizes[v][0], sizes[v][1], n, c) for v in range(V)]
    depths[0][5:15, 10:30] *= 1.05                           # reference map off by 5 % in a block
    depths[2][20:40, 20:50] += rng.normal(0, 30, (20, 30)).astype(np.float32)   # a noisy block in source 2
    depths[1][:, :6] = 0.0                                   # invalid (zero) depths at a border of source 1
    return K.astype(np.float32), R.astype(np.float32), t.astype(np.float32), depths


def main():
    # evaluation/pipeline_utils.py imports the dataset modules (cv2, h5py): not needed for the filter, stub them
    for name in ("data", "data.dtu_yao_eval", "data.yfcc_scene"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import_reference()
    from evaluation import filtering

    K, R, t, depths = scene()
    names = ["view%d" % v for v in range(len(depths))]
    args = types.SimpleNamespace(model="golden", nviews=5, scene="s0", upsample=False, downscale=1, depth_threshold=0.01,
                                 num_consistent=3, max_reproj_error=1.0, min_tri_angle=1.0, debug=False)
    with tempfile.TemporaryDirectory() as tmp:
        args.data_path = tmp
        folder = os.path.join(tmp, "IntRes", "depthmaps", "%s_%d" % (args.model, args.nviews), args.scene)
        os.makedirs(folder)
        for nme, d in zip(names, depths):
            np.savez(os.path.join(folder, nme + "_out.npz"), depthmap=d)
        batch = {"filename": [names[0]], "K": torch.from_numpy(K)[None].clone(), "R": torch.from_numpy(R)[None],
                 "t": torch.from_numpy(t)[None], "src_filenames": [[nme] for nme in names[1:]]}
        filtering.run([batch], args)
        res = np.load(os.path.join(tmp, "IntRes", "geometric_filtering", "%s_%d" % (args.model, args.nviews), args.scene,
                                   names[0] + "_out.npz"))
        out = {"K": K, "R": R, "t": t, "mask_depth": res["mask_depth"], "mask_disp": res["mask_disp"], "geo_mask": res["geo_mask"],
               "depth_threshold": np.float32(args.depth_threshold), "max_reproj_error": np.float32(args.max_reproj_error),
               "min_tri_angle": np.float32(args.min_tri_angle), "num_consistent": np.int32(args.num_consistent)}
        for v, d in enumerate(depths):
            out["depth%d" % v] = d
    np.savez_compressed(os.path.join(OUT, "geo_filter.npz"), **out)
    print("geo_filter: kept", {k: float(out[k].mean()) for k in ("mask_depth", "mask_disp", "geo_mask")})


if __name__ == "__main__":
    main()
