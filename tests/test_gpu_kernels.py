"""GPU: every CUDA kernel, called through the C ABI (wild_deep_mvs_b200.ops -> libmvsb200.so), against the
CPU oracle on the same seeded inputs and against the reference-generated goldens.

Tolerances are relative L-infinity (max|a-b| / max|b|).  fp32 arithmetic with a different summation order
gives ~1e-6; the projection inverse is fp64 here and fp32 in the reference, which moves sample positions by
~2e-5 px (measured), hence 1e-4 where the library computes its own relative projection."""
import numpy as np
import pytest
import torch

import oracle as orc
from conftest import assert_mismatches_on_boundary, expected_index_boundary_distance, rel_linf

pytestmark = pytest.mark.gpu

from wild_deep_mvs_b200 import _lib as L  # noqa: E402
from wild_deep_mvs_b200 import ops  # noqa: E402

DEV = "cuda:0"


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


def nhwc(a):  # numpy [C,H,W] -> cuda [1,H,W,C]
    return cu(np.transpose(a, (1, 2, 0))[None])


def ndhwc(a):  # numpy [C,D,H,W] -> cuda [1,D,H,W,C]
    return cu(np.transpose(a, (1, 2, 3, 0))[None])


def from_ndhwc(t):  # cuda [1,D,H,W,C] -> numpy [C,D,H,W]
    return t[0].permute(3, 0, 1, 2).contiguous().cpu().numpy()


def warp16(rot, trans):
    w = np.zeros(16, np.float32)
    w[:9] = np.asarray(rot).reshape(-1)
    w[9:12] = np.asarray(trans).reshape(-1)
    return w


@pytest.fixture(autouse=True, params=["pixel-tile", "depth-marching"])
def k1_kernel(request, monkeypatch):
    """K1 has two kernels and picks one per shape (k1_cost_volume.cu: the depth-marching one for long scalar sweeps, the
    pixel-tile one otherwise).  The K1 parity tests run every case through BOTH (MVSB200_K1 = v1 | m forces one; shapes the
    depth-marching kernel does not cover fall through to the other); every other test runs once."""
    if "k1_" not in request.node.name and "errors_are_loud" not in request.node.name:
        if request.param == "depth-marching":
            pytest.skip("not a K1 test: runs once")
        return None
    monkeypatch.setenv("MVSB200_K1", "v1" if request.param == "pixel-tile" else "m")
    return request.param


# ------------------------------------------------------------------------------------------------
# geometry prologues
# ------------------------------------------------------------------------------------------------
def test_mvs_relative_proj_matches_reference(golden):
    g = golden("mvsnet_variance")
    warp = ops.mvs_relative_proj(cu(g["proj"][:, 0]), cu(g["proj"][:, 1:])).cpu().numpy()
    ref = g["rel_proj"][0]
    for s in range(2):
        assert np.abs(warp[0, s, :9].reshape(3, 3) - ref[s, :3, :3]).max() < 1e-5 * np.abs(ref[s, :3, :3]).max()
        assert np.abs(warp[0, s, 9:12] - ref[s, :3, 3]).max() < 1e-5 * np.abs(ref[s, :3, 3]).max()


def test_singular_projection_gives_nan_not_garbage():
    z = torch.zeros(1, 4, 4, device=DEV)
    warp = ops.mvs_relative_proj(z, torch.eye(4, device=DEV).view(1, 1, 4, 4))
    assert torch.isnan(warp[0, 0, :12]).all()


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("agg", ["variance", "softmin"])
def test_k1_mvs_against_golden_and_oracle(golden, agg):
    g = golden("mvsnet_" + agg)
    feats = [g["feat0"][0], g["feat1"][0], g["feat2"][0]]
    dv = g["depth_values"]
    D = dv.shape[1]
    warp = np.stack([warp16(g["rel_proj"][0, s, :3, :3], g["rel_proj"][0, s, :3, 3]) for s in range(2)])[None]
    temp = cu(g["temp"]) if agg == "softmin" else None
    vol = ops.build_cost_volume(nhwc(feats[0]), [nhwc(feats[1]), nhwc(feats[2])], cu(warp), cu(dv), D, L.GEOM_MVS,
                                L.AGG_VARIANCE if agg == "variance" else L.AGG_SOFTMIN, temp=temp)
    got = from_ndhwc(vol)
    assert rel_linf(got, g["cost_volume"][0]) < 2e-5          # vs the reference itself
    H, W = feats[0].shape[1:]
    warped = [orc.homo_warp_mvs(feats[1 + s], g["rel_proj"][0, s, :3, :3], g["rel_proj"][0, s, :3, 3], dv[0], (H, W))
              for s in range(2)]
    want = orc.variance(feats[0], warped, 0) if agg == "variance" else orc.softmin(feats[0], warped, float(g["temp"][0]))
    assert rel_linf(got, want) < 1e-5                          # vs the oracle


def test_k1_variance_mean_order_and_c16(golden):
    """CVP flavour: 16 channels, (M1/V)^2 rounding order, per-pixel hypotheses [B,D,H,W]."""
    rng = np.random.default_rng(0)
    C, H, W, D, S = 16, 12, 20, 5, 3
    ref = rng.standard_normal((C, H, W)).astype(np.float32)
    srcs = [rng.standard_normal((C, 10 + s, 17 + s)).astype(np.float32) for s in range(S)]
    g = golden("warp_ragged")
    rot, trans = g["rel_proj"][0, :3, :3], g["rel_proj"][0, :3, 3]
    depth = (425 + 480 * rng.random((D, H, W))).astype(np.float32)
    depth[0, :2] = -30.0
    warp = np.stack([warp16(rot, trans * (1 + 0.1 * s)) for s in range(S)])[None]
    vol = ops.build_cost_volume(nhwc(ref), [nhwc(s) for s in srcs], cu(warp), cu(depth[None]), D, L.GEOM_MVS,
                                L.AGG_VARIANCE_MEAN)
    warped = [orc.homo_warp_mvs(srcs[s], rot, trans * np.float32(1 + 0.1 * s), depth, (H, W)) for s in range(S)]
    assert rel_linf(from_ndhwc(vol), orc.variance(ref, warped, 1)) < 1e-5


def test_k1_warp_only_ragged_source(golden):
    """A single source + variance lets the warp itself be recovered: V=2 => M1 = ref + w."""
    g = golden("warp_ragged")
    src = g["src"][0]
    C, H, W, D = 8, 10, 14, 5
    ref = np.zeros((C, H, W), np.float32)
    warp = warp16(g["rel_proj"][0, :3, :3], g["rel_proj"][0, :3, 3])[None, None]
    vol = ops.build_cost_volume(nhwc(ref), [nhwc(src)], cu(warp), cu(g["depth"]), D, L.GEOM_MVS, L.AGG_VARIANCE)
    # ref = 0, V = 2: cost = w^2/2 - w^2/4 = w^2/4
    want = g["warped"][0] ** 2 / 4
    assert rel_linf(from_ndhwc(vol), want) < 2e-5


def _vis_cams(g):
    return g["ref_cam"], np.stack([g["src_cam1"], g["src_cam2"]], 1)


def test_k1_vis_groupcorr_against_golden(golden):
    g = golden("vis")
    ref_cam, src_cams = _vis_cams(g)
    warp = ops.vis_homography_params(cu(ref_cam), cu(src_cams), 1.0 / 8)
    f_ref, f1, f2 = g["feat_v0_s1"][0], g["feat_v1_s1"][0], g["feat_v2_s1"][0]
    interval = np.float32((g["depth_max"][0, 0] - g["depth_min"][0, 0]) / np.float32(128)) * np.float32(4)
    out = ops.build_cost_volume(nhwc(f_ref), [nhwc(f1), nhwc(f2)], warp, cu(g["depth_min"][:, 0]), 8, L.GEOM_VIS,
                                L.AGG_GROUPCORR, interval=cu(np.array([interval])), groups=8)
    assert out.shape[0] == 2
    assert rel_linf(from_ndhwc(out[0]), g["s1_groupcorr_pair0"][0]) < 1e-4


def test_k1_vis_per_pixel_start(golden):
    g = golden("vis")
    ref_cam, src_cams = _vis_cams(g)
    rng = np.random.default_rng(1)
    f_ref, f1 = g["feat_v0_s2"][0], g["feat_v1_s2"][0]
    H, W = f_ref.shape[1:]
    start = (450 + 100 * rng.random((H, W))).astype(np.float32)
    interval = np.float32(3.75)
    warp = ops.vis_homography_params(cu(ref_cam), cu(src_cams[:, :1]), 1.0 / 4)
    out = ops.build_cost_volume(nhwc(f_ref), [nhwc(f1)], warp, cu(start[None]), 4, L.GEOM_VIS, L.AGG_GROUPCORR,
                                interval=cu(np.array([interval])), groups=8)
    from oracle import nets
    warped = orc.vis_warp(f1, nets.scale_cam(ref_cam[0], 0.25), nets.scale_cam(src_cams[0, 0], 0.25), start, interval, 4, (H, W))
    assert rel_linf(from_ndhwc(out[0]), orc.groupcorr(f_ref, warped, 8)) < 1e-4


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # cin, cout, k, stride, transposed, dims
    (32, 8, (3, 3, 3), 1, False, (6, 10, 36)),
    (8, 16, (3, 3, 3), 2, False, (8, 12, 40)),
    (16, 16, (3, 3, 3), 1, False, (4, 9, 33)),
    (64, 64, (3, 3, 3), 1, False, (2, 5, 7)),
    (16, 8, (3, 3, 3), 2, True, (3, 5, 34)),
    (64, 32, (3, 3, 3), 2, True, (2, 3, 5)),
    (8, 16, (1, 1, 1), 2, False, (6, 8, 10)),
    (8, 1, (3, 3, 3), 1, False, (5, 9, 35)),
    (16, 1, (3, 3, 3), 1, False, (3, 4, 5)),
    (8, 8, (1, 3, 3), 1, False, (1, 11, 37)),
    (1, 8, (1, 3, 3), 1, False, (1, 9, 13)),
    (8, 1, (1, 3, 3), 1, False, (1, 9, 13)),
]


# shapes that exercise tile edges of the tensor-core engine (tiles are 2 x 7 x 32 voxels)
TC_CASES = [
    (8, 8, (3, 3, 3), 1, False, (3, 15, 65)),
    (32, 32, (3, 3, 3), 1, False, (4, 8, 40)),
    (16, 32, (3, 3, 3), 2, False, (6, 16, 70)),
    (32, 64, (3, 3, 3), 2, False, (4, 6, 10)),
    (32, 16, (3, 3, 3), 2, True, (3, 8, 35)),
    (24, 1, (3, 3, 3), 1, False, (2, 7, 32)),
]


# shapes that exercise the z-march engine: several (y,x) tiles (7 x 32 / 7 x 16 outputs), several depth segments per
# tile column and more tiles than CTAs (persistent loop), odd sizes at every edge
ZM_CASES = [
    (32, 8, (3, 3, 3), 1, False, (19, 30, 70)),
    (8, 8, (3, 3, 3), 1, False, (40, 50, 131)),
    (16, 16, (3, 3, 3), 1, False, (21, 23, 37)),
    (8, 16, (3, 3, 3), 2, False, (22, 30, 66)),
    (16, 32, (3, 3, 3), 2, False, (9, 15, 35)),
    (16, 8, (3, 3, 3), 2, True, (11, 16, 40)),
    (32, 16, (3, 3, 3), 2, True, (7, 9, 19)),
]


@pytest.mark.parametrize("engine", ["fp32", "zm", "tc", "tc_tf32"])
@pytest.mark.parametrize("cin,cout,k,stride,transposed,dims", CONV_CASES + TC_CASES + ZM_CASES)
def test_k2_conv_against_oracle(cin, cout, k, stride, transposed, dims, engine):
    rng = np.random.default_rng(cin * 131 + cout)
    x = rng.standard_normal((cin,) + dims).astype(np.float32)
    if transposed:
        w = (rng.standard_normal((cin, cout) + k) / np.sqrt(cin * 27)).astype(np.float32)
        want = orc.deconv3d(x, w, None, 2, 1, 1)
    else:
        w = (rng.standard_normal((cout, cin) + k) / np.sqrt(cin * np.prod(k))).astype(np.float32)
        want = orc.conv3d(x, w, None, stride)
    layer = ops.PackedConv(cu(w), None, stride=stride, transposed=transposed)
    got = from_ndhwc(ops.conv3d(ndhwc(x), layer, engine=engine))
    assert got.shape == want.shape
    # fp32 FMA chains, the fp16x2 split (zm) and the 3xTF32 split (tc) agree with the oracle to ~1e-6; plain TF32
    # rounds operands to 11 bits
    assert rel_linf(got, want) < (3e-3 if engine == "tc_tf32" else 1e-5)


@pytest.mark.parametrize("gain", [1e-6, 1.0, 3e5])
def test_k2_zm_scaling_is_range_safe(gain):
    """The fp16 split of the z-march engine works on x * 2^k with k from the tracked abs-max: tiny and huge tensors,
    and tensors with a few large outliers, keep fp32-equivalent accuracy."""
    rng = np.random.default_rng(11)
    dims = (6, 16, 40)
    x = (rng.standard_normal((32,) + dims) * gain).astype(np.float32)
    x[3, 2, 5, 7] = 4000.0 * gain   # outlier: sets the scale, everything else sits 2^-12 below it
    w = (rng.standard_normal((8, 32, 3, 3, 3)) / np.sqrt(32 * 27)).astype(np.float32)
    want = orc.conv3d(x, w, None, 1)
    got = from_ndhwc(ops.conv3d(ndhwc(x), ops.PackedConv(cu(w), None), engine="zm"))
    assert rel_linf(got, want) < 1e-5
    # away from the outlier the error relative to the typical magnitude stays small too
    err = np.abs(got - want)
    err[:, 0:5, 2:9, 4:11] = 0
    assert err.max() < 1e-5 * np.abs(want[:, :, :2]).max()


def test_absmax_and_tracked_amax():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 5, 9, 33, 8)).astype(np.float32)
    t = cu(x)
    assert float(ops.absmax(t)) == float(np.abs(x).max())
    w = (rng.standard_normal((8, 8, 3, 3, 3)) / 15).astype(np.float32)
    y = ops.conv3d(t, ops.PackedConv(cu(w), None), engine="zm")
    assert float(y._mvs_amax) == float(y.abs().max())       # the epilogue tracks max|y| for the next layer


def test_k2_fused_epilogue_and_concat():
    rng = np.random.default_rng(7)
    dims = (4, 6, 34)
    xa = rng.standard_normal((8,) + dims).astype(np.float32)
    xb = rng.standard_normal((8,) + dims).astype(np.float32)
    w = (rng.standard_normal((8, 16, 3, 3, 3)) / 20).astype(np.float32)
    skip = rng.standard_normal((8,) + dims).astype(np.float32)
    bn = torch.nn.BatchNorm3d(8).to(DEV).eval()
    with torch.no_grad():
        bn.weight.copy_(cu(rng.random(8) + 0.5)); bn.bias.copy_(cu(rng.standard_normal(8)))
        bn.running_mean.copy_(cu(rng.standard_normal(8) * 0.1)); bn.running_var.copy_(cu(rng.random(8) + 0.5))
    conv = orc.conv3d(np.concatenate([xa, xb], 0), w, None, 1)
    npy = lambda t: t.detach().cpu().numpy()
    y = orc.bn_relu(conv, npy(bn.weight), npy(bn.bias), npy(bn.running_mean), npy(bn.running_var), bn.eps, relu=False)
    for mode, want in ((L.SKIP_BEFORE_RELU, np.maximum(y + skip, 0)), (L.SKIP_AFTER_RELU, np.maximum(y, 0) + skip)):
        for engine in ("fp32", "tc", "zm"):
            layer = ops.PackedConv(cu(w), bn, relu=True, skip_mode=mode)
            got = from_ndhwc(ops.conv3d(ndhwc(xa), layer, x2=ndhwc(xb), skip=ndhwc(skip), engine=engine))
            assert rel_linf(got, want) < 1e-5, engine


def test_k2_stride1_transposed_conv_is_packed_as_flipped_conv():
    rng = np.random.default_rng(9)
    x = rng.standard_normal((64, 3, 4, 6)).astype(np.float32)
    w = (rng.standard_normal((64, 32, 3, 3, 3)) / 40).astype(np.float32)
    want = orc.deconv3d(x, w, None, 1, 1, 0)
    for engine in ("fp32", "tc", "zm"):
        got = from_ndhwc(ops.conv3d(ndhwc(x), ops.PackedConv(cu(w), None, stride=1, transposed=True), engine=engine))
        assert rel_linf(got, want) < 2e-5, engine   # K = 64*27 = 1728 products per output


def test_k2_mvsnet_costregnet_against_golden(golden):
    from wild_deep_mvs_b200.mvsnet import CostRegNet
    g = golden("mvsnet_variance")
    net = CostRegNet()
    sd = {k[len("cost_regularization."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("cost_regularization.")}
    net.load_state_dict(sd, strict=True)
    net = net.to(DEV).eval()
    out = net(cu(g["cost_volume"]))
    assert out.shape == (1, 1) + g["cost_reg"].shape[1:]
    assert rel_linf(out[0, 0].cpu().numpy(), g["cost_reg"][0]) < 2e-5


# ------------------------------------------------------------------------------------------------
# K3 / K4
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [8, 48, 192])
def test_k3_all_modes_against_oracle(D):
    rng = np.random.default_rng(D)
    H, W = 9, 37
    score = (rng.standard_normal((D, H, W)) * 4).astype(np.float32)
    dv = np.linspace(425, 905, D).astype(np.float32)
    # values + MVSNet confidence
    got = ops.depth_regress(cu(score[None]), cu(dv[None]), conf_mode=L.CONF_SUM4, want_entropy=True, want_prob=True)
    want = orc.softmax_regress(score, 0, dv, conf_mode=1, want_entropy=True, want_prob=True)
    assert rel_linf(got["depth"][0].cpu().numpy(), want["depth"]) < 1e-5
    assert rel_linf(got["prob"][0].cpu().numpy(), want["prob"]) < 1e-5
    assert rel_linf(got["entropy"][0].cpu().numpy(), want["entropy"]) < 1e-4
    # the confidence sums the 4 bins around floor(expected index): index work -- it may differ from the oracle ONLY where the
    # expected index sits on an integer (1e-5 of the index range), and such pixels are rare
    dist = expected_index_boundary_distance(want["prob"])
    n_bad = assert_mismatches_on_boundary(got["conf"][0].cpu().numpy(), want["conf"], dist, 1e-4, 1e-5 * D, "K3 CONF_SUM4")
    assert n_bad <= max(1, int(0.01 * H * W)), n_bad
    # per-pixel hypotheses
    dpp = (dv[:, None, None] + rng.random((D, H, W))).astype(np.float32)
    got = ops.depth_regress(cu(score[None]), cu(dpp[None]))
    assert rel_linf(got["depth"][0].cpu().numpy(), orc.softmax_regress(score, 1, dpp)["depth"]) < 1e-5
    # start (+map) + interval, Vis window confidence
    start = (450 + 50 * rng.random((H, W))).astype(np.float32)
    got = ops.depth_regress(cu(score[None]), cu(start[None]), interval=cu(np.array([2.5])), conf_mode=L.CONF_WINDOW)
    want = orc.softmax_regress(score, 2, start, 2.5, conf_mode=2)
    assert rel_linf(got["depth"][0].cpu().numpy(), want["depth"]) < 1e-5
    n_bad = assert_mismatches_on_boundary(got["conf"][0].cpu().numpy(), want["conf"], dist, 1e-4, 1e-5 * D, "K3 CONF_WINDOW")
    assert n_bad <= max(1, int(0.01 * H * W)), n_bad
    got = ops.depth_regress(cu(score[None]), cu(np.array([425.0])), interval=cu(np.array([2.5])))
    assert rel_linf(got["depth"][0].cpu().numpy(), orc.softmax_regress(score, 2, np.array([425.0]), 2.5)["depth"]) < 1e-5


def test_k3_peaked_logits_pick_the_peak():
    D, H, W = 192, 8, 16
    score = torch.full((1, D, H, W), -30.0, device=DEV)
    idx = torch.randint(2, D - 3, (H, W), device=DEV)
    score[0].scatter_(0, idx[None], 30.0)
    dv = torch.linspace(425, 905, D, device=DEV)[None]
    out = ops.depth_regress(score, dv, conf_mode=L.CONF_SUM4)
    assert torch.allclose(out["depth"][0], dv[0][idx], rtol=1e-6)
    assert torch.allclose(out["conf"], torch.ones_like(out["conf"]), atol=1e-6)


def test_k4_vis_fuse_against_oracle():
    rng = np.random.default_rng(3)
    S, G, D, H, W = 3, 8, 4, 6, 10
    interm = [rng.standard_normal((G, D, H, W)).astype(np.float32) for _ in range(S)]
    unc = [rng.standard_normal((H, W)).astype(np.float32) for _ in range(S)]
    got = ops.vis_fuse([ndhwc(a) for a in interm], [cu(u[None]) for u in unc])
    assert rel_linf(from_ndhwc(got), orc.vis_fuse(interm, unc)) < 1e-5


# ------------------------------------------------------------------------------------------------
# error behaviour
# ------------------------------------------------------------------------------------------------
def test_errors_are_loud():
    x = torch.zeros(1, 4, 4, 4, 8, device=DEV)
    layer = ops.PackedConv(torch.zeros(8, 16, 3, 3, 3, device=DEV))
    with pytest.raises(L.Mvsb200Error):
        ops.conv3d(x, layer)  # channel mismatch
    with pytest.raises(L.Mvsb200Error):
        ops.build_cost_volume(torch.zeros(1, 4, 4, 12, device=DEV), [torch.zeros(1, 4, 4, 12, device=DEV)],
                              torch.zeros(1, 1, 16, device=DEV), torch.ones(1, 2, device=DEV), 2, L.GEOM_MVS, L.AGG_VARIANCE)
    with pytest.raises(L.Mvsb200Error):
        ops.conv3d(torch.zeros(1, 4, 4, 4, 16), layer)  # CPU tensor


def test_k6_uncert_net_against_oracle():
    """Fused UncertNet (conv 1->8 + BN + ReLU, conv 8->8 + BN + ReLU + input, 8->1 head) vs the composed oracle convs,
    on maps whose sizes are not multiples of the 16 x 16 tile."""
    from oracle import nets
    rng = np.random.default_rng(21)
    N, H, W = 3, 37, 53
    ent = rng.random((N, H, W)).astype(np.float32) * 3
    sd = {"u.conv1.0.weight": (rng.standard_normal((8, 1, 3, 3)) / 3).astype(np.float32),
          "u.conv2.0.weight": (rng.standard_normal((8, 8, 3, 3)) / 8).astype(np.float32),
          "u.head_convs.0.weight": (rng.standard_normal((1, 8, 3, 3)) / 8).astype(np.float32)}
    bns = []
    for name in ("u.conv1.1", "u.conv2.1"):
        bn = torch.nn.BatchNorm2d(8).to(DEV).eval()
        with torch.no_grad():
            bn.weight.copy_(cu(rng.random(8) + 0.5)); bn.bias.copy_(cu(rng.standard_normal(8) * 0.1))
            bn.running_mean.copy_(cu(rng.standard_normal(8) * 0.1)); bn.running_var.copy_(cu(rng.random(8) + 0.5))
        for k in ("weight", "bias", "running_mean", "running_var"):
            sd[name + "." + k] = getattr(bn, k).detach().cpu().numpy()
        bns.append(bn)
    params = ops.uncert_net_params(ops.PackedConv(cu(sd["u.conv1.0.weight"]), bns[0], relu=True),
                                   ops.PackedConv(cu(sd["u.conv2.0.weight"]), bns[1], relu=True),
                                   ops.PackedConv(cu(sd["u.head_convs.0.weight"])))
    got = ops.vis_uncert_net(cu(ent), params).cpu().numpy()
    for n in range(N):
        want = nets.vis_uncert_net(sd, "u.", ent[n][None])
        assert rel_linf(got[n], want[0]) < 1e-5


def test_k2_zm_split_over_input_channels_with_fused_epilogue():
    """A 64 -> 64 layer does not fit the z-march engine's resident-weight budget: ops.conv3d runs it as two launches over
    the halves of the input channels, the second adding the first's output in place.  BN, ReLU and a skip connection
    before the activation must come out exactly as for a single launch."""
    rng = np.random.default_rng(13)
    dims = (3, 9, 21)
    x = rng.standard_normal((64,) + dims).astype(np.float32)
    w = (rng.standard_normal((64, 64, 3, 3, 3)) / np.sqrt(64 * 27)).astype(np.float32)
    skip = rng.standard_normal((64,) + dims).astype(np.float32)
    bn = torch.nn.BatchNorm3d(64).to(DEV).eval()
    with torch.no_grad():
        bn.weight.copy_(cu(rng.random(64) + 0.5)); bn.bias.copy_(cu(rng.standard_normal(64)))
        bn.running_mean.copy_(cu(rng.standard_normal(64) * 0.1)); bn.running_var.copy_(cu(rng.random(64) + 0.5))
    npy = lambda t: t.detach().cpu().numpy()
    y = orc.bn_relu(orc.conv3d(x, w, None, 1), npy(bn.weight), npy(bn.bias), npy(bn.running_mean), npy(bn.running_var), bn.eps, relu=False)
    layer = ops.PackedConv(cu(w), bn, relu=True, skip_mode=L.SKIP_BEFORE_RELU)
    got = ops.conv3d(ndhwc(x), layer, skip=ndhwc(skip), engine="zm")
    assert rel_linf(from_ndhwc(got), np.maximum(y + skip, 0)) < 1e-5
    assert float(got._mvs_amax) == float(got.abs().max())
    plain = ops.PackedConv(cu(w), bn, relu=True)
    assert rel_linf(from_ndhwc(ops.conv3d(ndhwc(x), plain, engine="zm")), np.maximum(y, 0)) < 1e-5
