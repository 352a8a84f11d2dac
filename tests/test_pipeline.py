"""Callers either side of the depth path on the reference's directory layout (wild_deep_mvs_b200/pipeline.py;
evaluation/run_depthmaps.py, filtering.py, pipeline_utils.py, fusibile.py of the reference)."""
import os
import types

import numpy as np
import pytest
import torch

from wild_deep_mvs_b200 import formats, pipeline


def make_args(tmp, **kw):
    a = types.SimpleNamespace(model="golden", nviews=5, scene="s0", data_path=str(tmp), upsample=False, downscale=1,
                              depth_threshold=0.01, num_consistent=3, max_reproj_error=1.0, min_tri_angle=1.0, debug=False,
                              override=False, prob_threshold=0.5, filter=False, colmap=False)
    a.__dict__.update(kw)
    return a


class FakeNet(torch.nn.Module):
    """Stands in for a drop-in model: same call signature and result keys (models/MVSNet/model.py:178,217-218)."""

    def __init__(self):
        super().__init__()
        self.calls = 0

    def forward(self, imgs, K, R, t, depth_min, depth_max):
        self.calls += 1
        B, V, _, H, W = imgs.shape
        depth = depth_min[:, :1, None] + imgs[:, 0].mean(1)[:, ::4, ::4] * (depth_max - depth_min)[:, :1, None]
        return {"depth": depth, "photometric_confidence": imgs[:, 0, 0, ::4, ::4]}


def batches(n, B=2):
    g = torch.Generator().manual_seed(0)
    for i in range(n):
        yield {"filename": ["v%d_%d" % (i, j) for j in range(B)], "imgs": torch.rand(B, 3, 3, 16, 24, generator=g),
               "K": torch.eye(3).repeat(B, 3, 1, 1), "R": torch.eye(3).repeat(B, 3, 1, 1), "t": torch.zeros(B, 3, 3, 1),
               "depth_min": torch.full((B, 3), 425.0), "depth_max": torch.full((B, 3), 905.0)}


def test_run_depthmaps_layout_skip_and_finish(tmp_path):
    args, net = make_args(tmp_path), FakeNet()
    pipeline.run_depthmaps(list(batches(2)), args, net)
    out = tmp_path / "IntRes" / "depthmaps" / "golden_5" / "s0"
    assert sorted(os.listdir(out)) == ["finished.txt", "v0_0_out.npz", "v0_1_out.npz", "v1_0_out.npz", "v1_1_out.npz"]
    d, p = formats.load_depth_npz(out / "v1_1_out.npz")
    assert d.shape == (4, 6) and p.shape == (4, 6) and d.dtype == np.float32 and 425 <= d.min() and d.max() <= 905
    pipeline.run_depthmaps(list(batches(2)), args, net)            # finished.txt: nothing is recomputed
    assert net.calls == 2
    os.remove(out / "finished.txt")
    os.remove(out / "v0_1_out.npz")
    pipeline.run_depthmaps(list(batches(2)), args, net)            # only the batch with a missing view runs again
    assert net.calls == 3 and (out / "v0_1_out.npz").exists() and (out / "finished.txt").exists()
    args.override = True
    pipeline.run_depthmaps(list(batches(2)), args, net)
    assert net.calls == 5


def test_get_mask(tmp_path):
    args = make_args(tmp_path)
    prob = np.array([[0.9, 0.2], [0.5, 0.49]], np.float32)
    assert np.array_equal(pipeline.get_mask(args, "x", prob=prob), [[False, True], [False, True]])
    prob3 = np.stack([prob, prob[::-1], np.full((2, 2), 0.1, np.float32)])
    assert np.array_equal(pipeline.get_mask(args, "x", prob=prob3), [[False, True], [False, True]])   # invalid in EVERY channel
    args.filter = True
    geo = np.array([[True, True], [False, True]])
    assert np.array_equal(pipeline.get_mask(args, "x", prob=prob, geo_mask=geo), [[False, True], [True, True]])
    d = tmp_path / "IntRes" / "geometric_filtering" / "golden_5" / "s0"
    d.mkdir(parents=True)
    np.savez_compressed(d / "x_out.npz", geo_mask=geo)
    assert np.array_equal(pipeline.get_mask(args, "x", prob=prob), [[False, True], [True, True]])   # read from disk
    with pytest.raises(NotImplementedError):
        pipeline.get_mask(args, "x")


def test_mvsnet_to_gipuma(tmp_path):
    args = make_args(tmp_path, downscale=2, prob_threshold=0.3)
    b = next(batches(1, B=1))
    K = torch.tensor([[300.0, 0, 12], [0, 310.0, 8], [0, 0, 1]])
    b["K"] = K.repeat(1, 3, 1, 1)
    b["t"] = torch.tensor([[-40.0], [1.0], [0.5]]).repeat(1, 3, 1, 1)
    depth = np.linspace(500, 700, 8 * 12, dtype=np.float32).reshape(8, 12)
    prob = np.linspace(0, 1, 8 * 12, dtype=np.float32).reshape(8, 12)
    dd = tmp_path / "IntRes" / "depthmaps" / "golden_5" / "s0"
    dd.mkdir(parents=True)
    formats.save_depth_npz(dd / "v0_0_out.npz", depth, prob)
    pts = tmp_path / "pts"
    pipeline.mvsnet_to_gipuma(args, pts, [b])
    rows = [[float(v) for v in l.split()] for l in open(pts / "cams" / "v0_0.jpg.P").read().splitlines() if l.strip()]
    assert np.allclose(rows, [[150, 0, 6, -6000 + 3], [0, 155, 4, 155 + 2], [0, 0, 1, 0.5]])
    got = formats.read_gipuma_dmb(pts / "2333__v0_0" / "disp.dmb")
    assert np.array_equal(got, np.where(prob < 0.3, 0, depth))
    nrm = formats.read_gipuma_dmb(pts / "2333__v0_0" / "normals.dmb")
    assert nrm.shape == (8, 12, 3)
    if (pts / "images" / "v0_0.jpg").exists():
        from PIL import Image
        assert Image.open(pts / "images" / "v0_0.jpg").size == (12, 8)


@pytest.mark.gpu
def test_run_filtering_matches_masks_written_by_the_reference(tmp_path, golden):
    g = golden("geo_filter")
    args = make_args(tmp_path)
    names = ["view%d" % v for v in range(5)]
    dd = tmp_path / "IntRes" / "depthmaps" / "golden_5" / "s0"
    dd.mkdir(parents=True)
    for v, n in enumerate(names):
        np.savez(dd / (n + "_out.npz"), depthmap=g["depth%d" % v])
    batch = {"filename": [names[0]], "K": torch.from_numpy(g["K"])[None].clone(), "R": torch.from_numpy(g["R"])[None],
             "t": torch.from_numpy(g["t"])[None], "src_filenames": [[n] for n in names[1:]]}
    pipeline.run_filtering([batch], args)
    out = tmp_path / "IntRes" / "geometric_filtering" / "golden_5" / "s0"
    assert (out / "finished.txt").exists()
    res = np.load(out / "view0_out.npz")
    for k in ("mask_depth", "mask_disp", "geo_mask"):
        assert res[k].dtype == np.bool_ and res[k].shape == g[k].shape
        assert (res[k] != g[k]).mean() < 2e-3, k
    assert torch.equal(batch["K"], torch.from_numpy(g["K"])[None])     # the caller's intrinsics are left alone
