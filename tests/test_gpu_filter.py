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


# ---- K9: the consumer of the all-gathered depth maps (masked_photometricloss, models/trainer.py:240-278) ----------------
from oracle.filter import gathered_masks as oracle_gathered  # noqa: E402
from wild_deep_mvs_b200.filtering import gathered_masks  # noqa: E402


def check_gathered(out, orc, geom_clamping, eps=1e-4):
    """Masks are boolean work: the kernel may differ from the oracle only where the pixel sits ON a boundary -- the grid
    within eps of +-1 (inside test) or the relative depth difference within eps (relative) of geom_clamping."""
    got = out["masks"].cpu().numpy()
    assert got.shape == orc["masks"].shape
    g = orc["flows"].astype(np.float64)
    edge = np.min(np.abs(np.abs(g) - 1.0), axis=-1)
    thr = np.abs(orc["reproj_diff"].astype(np.float64) - geom_clamping) / geom_clamping
    margin = np.minimum(edge, thr)
    bad = got != orc["masks"]
    off = bad & ~(margin < eps)
    assert not off.any(), "%d of %d differing pixels are not on a boundary" % (int(off.sum()), int(bad.sum()))
    assert bad.mean() < 1e-3, bad.mean()
    print("gathered masks: %d of %d pixels differ, all within %.0e of a boundary" % (int(bad.sum()), bad.size, eps))
    assert np.abs(out["flows"].cpu().numpy() - orc["flows"]).max() < 2e-5
    assert np.abs(out["depth_src"].cpu().numpy() - orc["depth_src"]).max() < 1e-5 * np.abs(orc["depth_src"]).max()
    ok = np.isfinite(orc["warped_depth"])
    # a sample next to a depth discontinuity (the scenes have 650-unit steps) amplifies the last bits of the grid coordinate
    assert np.abs(out["warped_depth"].cpu().numpy() - orc["warped_depth"])[ok].max() < 2e-4 * np.abs(orc["warped_depth"][ok]).max()
    if orc["warped"] is not None:
        assert np.abs(out["warped"].cpu().numpy() - orc["warped"]).max() < 5e-5
    assert (out["inside"].cpu().numpy() != orc["inside"]).mean() < 1e-3


def test_gathered_masks_against_the_reference(golden):
    g = golden("gathered_masks")
    out = gathered_masks(cu(g["ref_depth"]), cu(g["gathered"]), cu(g["proj"]), int(g["ref_idx"]), float(g["geom_clamping"]),
                         imgs=cu(g["imgs"]), want=("inside", "flows", "depth_src", "warped_depth", "warped"))
    orc = oracle_gathered(g["ref_depth"], g["gathered"], g["proj"], int(g["ref_idx"]), float(g["geom_clamping"]), g["imgs"])
    assert (orc["masks"] != g["masks"]).sum() == 0          # the oracle IS the reference on this scene (also a CPU test)
    check_gathered(out, orc, float(g["geom_clamping"]))
    # what the unmodified reference recorded: the warped source images under its masks
    got = np.clip((out["warped"] * out["masks"][:, :, None]).cpu().numpy(), 0, 1)
    differ = (out["masks"].cpu().numpy() != g["masks"])[:, :, None]
    assert np.abs(got - g["warped_masked"])[~np.broadcast_to(differ, got.shape)].max() < 5e-5


@pytest.mark.parametrize("ref_idx", [0, 2, 4])
def test_gathered_masks_against_oracle_random_scene(ref_idx):
    rng = np.random.default_rng(11 + ref_idx)
    b, N, h, w = 2, 5, 96, 136
    proj = np.zeros((b, N, 4, 4), np.float32)
    gathered = np.zeros((b, N, h, w), np.float32)
    for bi in range(b):
        for v in range(N):
            K = np.array([[700.0, 0, w / 2.0], [0, 690.0, h / 2.0], [0, 0, 1]])
            a = 0.015 * v * (1 if v % 2 else -1)
            R = np.array([[np.cos(a), 0, np.sin(a)], [0, 1, 0], [-np.sin(a), 0, np.cos(a)]])
            t = np.array([[-12.0 * v * (1 if v % 2 else -0.7)], [2.0 * v], [0.5 * v]])
            P = np.eye(4)
            P[:3, :3], P[:3, 3:] = K @ R, K @ t
            proj[bi, v] = P
            ys, xs = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
            gathered[bi, v] = 650 + 0.4 * xs - 0.3 * ys + 30 * np.sin(xs / 9.0 + v) + rng.standard_normal((h, w)) * 6
    gathered[:, :, :6] = 0.0        # a band of zero depth (the clamp of the divisor) ...
    gathered[0, 1, 40:50, 60:80] = -5.0   # ... and a block behind the camera
    ref_depth = gathered[:, ref_idx].copy()
    imgs = rng.random((b, N, 3, h, w)).astype(np.float32)
    out = gathered_masks(cu(ref_depth), cu(gathered), cu(proj), ref_idx, 0.05, imgs=cu(imgs),
                         want=("inside", "flows", "depth_src", "warped_depth", "warped"))
    orc = oracle_gathered(ref_depth, gathered, proj, ref_idx, 0.05, imgs)
    assert 0.05 < orc["masks"].mean() < 0.95
    check_gathered(out, orc, 0.05)


def test_gathered_masks_rejects_bad_arguments():
    from wild_deep_mvs_b200._lib import Mvsb200Error
    d = torch.ones(1, 3, 8, 8, device=DEV)
    P = torch.eye(4, device=DEV).repeat(1, 3, 1, 1)
    with pytest.raises(Mvsb200Error):
        gathered_masks(d[:, 0], d, P, 3)                       # reference index out of range
    with pytest.raises(Mvsb200Error):
        gathered_masks(d[:, 0], d, P[:, :2], 0)                # projections of another view count
    with pytest.raises(Mvsb200Error):
        gathered_masks(d[:, 0], d, P, 0, want=("warped",))     # warped images without images
