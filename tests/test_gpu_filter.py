"""GPU: K8, the geometric-consistency filter (row f3) -- against the masks written by the unmodified reference
(tests/golden/geo_filter.npz) and against the numpy oracle on a larger randomised scene."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle.filter import geometric_filter as oracle_filter  # noqa: E402
from wild_deep_mvs_b200.filtering import geometric_filter  # noqa: E402

DEV = "cuda:0"


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


def threshold_margins(seams, depth0, depth_threshold, max_reproj_error):
    """Per pixel, how close the nearest per-source test of the filter is to flipping (relative to its threshold): a vote
    mask is boolean work and may differ between two correct fp32 evaluations only where one of its tests sits on its
    threshold.  Returns (margin of the depth tests, margin of the reprojection tests), each [h,w]."""
    dep, disp = [], []
    for sm in seams:
        drep = sm["depth_reproj"].astype(np.float64)
        lim = np.maximum(drep, depth0) * depth_threshold
        dep.append(np.minimum(np.abs(np.abs(drep - depth0) - lim) / np.maximum(lim, 1e-30),
                              np.minimum(np.abs(drep), np.abs(sm["proj_depth_in_src"].astype(np.float64)))))   # ... or a depth at 0
        disp.append(np.abs(sm["reproj_error"].astype(np.float64) - max_reproj_error) / max_reproj_error)
    return np.min(dep, 0), np.min(disp, 0)


def check_masks(out, want, seams, depth0, depth_threshold, max_reproj_error, eps=2e-3):
    m_dep, m_disp = threshold_margins(seams, np.asarray(depth0, np.float64), depth_threshold, max_reproj_error)
    margin = {"mask_depth": m_dep, "mask_disp": m_disp, "geo_mask": np.minimum(m_dep, m_disp)}
    for k in ("mask_depth", "mask_disp", "geo_mask"):
        got = out[k].cpu().numpy()
        assert got.shape == want[k].shape
        bad = got != want[k]
        off = bad & ~(margin[k] < eps)
        assert not off.any(), "%s: %d of %d differing pixels are not on a threshold (worst margin %.3g)" % (
            k, int(off.sum()), int(bad.sum()), float(margin[k][off].max()))
        assert bad.mean() < 2e-3, (k, bad.mean())
        print("%s: %d of %d pixels differ, all within %.0e (relative) of a threshold" % (k, int(bad.sum()), bad.size, eps))


def test_against_reference_masks(golden):
    g = golden("geo_filter")
    out = geometric_filter(cu(g["depth0"]), [cu(g["depth%d" % v]) for v in range(1, 5)], cu(g["K"]), cu(g["R"]), cu(g["t"]),
                           float(g["depth_threshold"]), float(g["max_reproj_error"]), float(g["min_tri_angle"]), int(g["num_consistent"]))
    # bit-exact up to pixels sitting ON a threshold (fp32 evaluation order differs from ATen's matmuls): asserted per pixel with
    # the float seams of the oracle, which reproduces the reference's masks pixel for pixel (tests/test_oracle_golden.py)
    orc = oracle_filter(g["depth0"], [g["depth%d" % v] for v in range(1, 5)], g["K"], g["R"], g["t"], float(g["depth_threshold"]),
                        float(g["max_reproj_error"]), float(g["min_tri_angle"]), int(g["num_consistent"]))
    check_masks(out, g, orc["seams"], g["depth0"], float(g["depth_threshold"]), float(g["max_reproj_error"]))


def test_against_oracle_on_a_random_scene_with_ragged_sources():
    rng = np.random.default_rng(7)
    h, w, N = 120, 168, 6
    K = np.zeros((N + 1, 3, 3), np.float32)
    R = np.zeros((N + 1, 3, 3), np.float32)
    t = np.zeros((N + 1, 3, 1), np.float32)
    sizes = [(h, w)] + [(h - 8 * (v % 3), w - 16 * (v % 2)) for v in range(1, N + 1)]
    for v in range(N + 1):
        hv, wv = sizes[v]
        K[v] = [[900.0 * wv / w, 0, wv / 2.0], [0, 890.0 * hv / h, hv / 2.0], [0, 0, 1]]
        a = 0.03 * v * (1 if v % 2 else -1)
        R[v] = [[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]]
        t[v] = [[-25.0 * v * (1 if v % 2 else -0.8)], [3.0 * v], [1.0 * v]]
    # a fronto-parallel-ish plane per view (analytic depth) plus noise of a few thresholds' width
    depths = []
    n, c = np.array([0.05, 0.08, 1.0]), 700.0
    for v in range(N + 1):
        hv, wv = sizes[v]
        ys, xs = np.meshgrid(np.arange(hv, dtype=np.float64), np.arange(wv, dtype=np.float64), indexing="ij")
        rays = np.stack([xs, ys, np.ones_like(xs)], -1) @ np.linalg.inv(K[v].astype(np.float64)).T
        d = rays @ R[v].astype(np.float64)
        o = -R[v].astype(np.float64).T @ t[v].astype(np.float64).reshape(3)
        depths.append(((c - n @ o) / (d @ n) * (1 + 0.006 * rng.standard_normal((hv, wv)))).astype(np.float32))
    want = oracle_filter(depths[0], depths[1:], K, R, t, 0.01, 1.0, 1.0, 3)
    out = geometric_filter(cu(depths[0]), [cu(d) for d in depths[1:]], cu(K), cu(R), cu(t), 0.01, 1.0, 1.0, 3, want_votes=True)
    for k in ("mask_depth", "mask_disp", "geo_mask"):
        assert 0.05 < want[k].mean() < 0.95, (k, want[k].mean())     # the noise level makes the vote non-trivial
    check_masks(out, want, want["seams"], depths[0], 0.01, 1.0)
    votes = out["votes"].cpu().numpy()
    assert votes.max() <= N and (votes[2] <= votes[0]).all() and (votes[2] <= votes[1]).all()


def test_errors_are_loud():
    from wild_deep_mvs_b200 import _lib as L
    d = torch.ones(8, 8, device=DEV)
    eye = torch.eye(3, device=DEV).repeat(2, 1, 1)
    with pytest.raises(L.Mvsb200Error):
        geometric_filter(d, [], eye[:1], eye[:1], torch.zeros(1, 3, device=DEV))
    with pytest.raises(L.Mvsb200Error):
        geometric_filter(d, [d], eye, eye, torch.zeros(3, 3, device=DEV))   # 1 + N cameras expected
