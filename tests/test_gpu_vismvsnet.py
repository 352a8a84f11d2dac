"""GPU: drop-in Vis-MVSNet (wild_deep_mvs_b200.vismvsnet.Frontend) against the reference-generated golden
(tests/golden/vis.npz: 1+2 views, 64x80 image, depth_nums [8,4,4]) and the C oracle."""
import numpy as np
import pytest
import torch

from conftest import assert_mismatches_on_boundary, expected_index_boundary_distance, rel_linf

pytestmark = pytest.mark.gpu

from wild_deep_mvs_b200 import ops, synth  # noqa: E402
from wild_deep_mvs_b200.vismvsnet import Frontend  # noqa: E402

DEV = "cuda:0"
DEPTH_TOL = 1e-3


def _load(g):
    net = Frontend()
    sd = net.state_dict()
    n = 0
    for k, v in g.items():
        if k.startswith("model.stage"):
            assert k in sd, k
            sd[k] = torch.from_numpy(v)
            n += 1
    assert n > 100
    net.load_state_dict(sd, strict=True)
    net.depth_nums, net.interval_scales = [8, 4, 4], [4, 2, 1]
    return net.to(DEV).eval()


def _inputs(g):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    feats = [[ops.to_nhwc(t(g["feat_v%d_s%d" % (v, k)])) for k in (1, 2, 3)] for v in range(3)]
    ref_cam = t(g["ref_cam"])
    src_cams = torch.stack([t(g["src_cam1"]), t(g["src_cam2"])], 1)
    dmin = t(g["depth_min"][:, 0])
    interval = ((t(g["depth_max"]) - t(g["depth_min"])) / 128)[:, 0].contiguous()
    return feats, ref_cam, src_cams, dmin, interval


def test_state_dict_names_match_reference(golden):
    g = golden("vis")
    sd = Frontend().state_dict()
    for k, v in g.items():
        if k.startswith("model."):
            assert k in sd and tuple(sd[k].shape) == v.shape, k
    for k in ("model.feat_ext.init_conv.0.weight", "model.feat_ext.unet.enc_blocks.2d2_0.0.downsample.0.weight",
              "model.feat_ext.unet.dec_blocks.2d16_3.2.0.conv1.weight", "model.feat_ext.final_conv_3.weight",
              "model.stage3.uncert_net.head_convs.0.weight"):
        assert k in sd, k


def test_stage1_seams_against_reference(golden):
    g = golden("vis")
    net = _load(g)
    feats, ref_cam, src_cams, dmin, interval = _inputs(g)
    d, p, pairs = net.model.stage1.run(feats[0][0], [feats[1][0], feats[2][0]], ref_cam, src_cams, 8, dmin,
                                       (interval * 4).contiguous(), 8)
    assert rel_linf(d.cpu().numpy(), g["depth_est_2"]) < DEPTH_TOL
    assert rel_linf(pairs[0][0][:, 0].cpu().numpy(), g["pair_depth_st2_v0"][:, 0]) < DEPTH_TOL
    assert rel_linf(pairs[1][0][:, 0].cpu().numpy(), g["pair_depth_st2_v1"][:, 0]) < DEPTH_TOL
    assert np.abs(pairs[0][1][0][:, 0].cpu().numpy() - g["pair_uncert_st2_v0"][:, 0]).max() < 2e-3


def test_cascade_against_reference_and_oracle(golden):
    from oracle import nets
    g = golden("vis")
    net = _load(g)
    feats, ref_cam, src_cams, dmin, interval = _inputs(g)
    ests, probs, pairs = net.depth_from_features(feats, ref_cam, src_cams, dmin, interval, [8, 4, 4], [4, 2, 1])
    for k in range(3):
        assert rel_linf(ests[2 - k].cpu().numpy(), g["depth_est_%d" % k]) < DEPTH_TOL
    assert rel_linf(ests[2].cpu().numpy(), g["depth"]) < 2e-4   # what fp32 actually achieves
    np_feats = [[g["feat_v%d_s%d" % (v, k)][0] for k in (1, 2, 3)] for v in range(3)]
    want = nets.vis_from_features(g, np_feats, g["ref_cam"][0], [g["src_cam1"][0], g["src_cam2"][0]],
                                  g["depth_min"][0, 0], g["depth_max"][0, 0], [8, 4, 4], [4, 2, 1])
    assert rel_linf(ests[2][0].cpu().numpy(), want["depth"]) < 2e-4
    # probability maps (sum of the softmax over |d - expected index| <= 2, nn_utils.py:463-465) of every stage against the
    # oracle's regulariser output: window membership is index work, so a pixel may differ only where its expected index sits
    # on an integer -- asserted per pixel
    for k in range(3):
        score = want["seams"][k]["fuse_score"].astype(np.float64)                   # [D,h,w]
        pr = np.exp(score - score.max(0, keepdims=True))
        pr /= pr.sum(0, keepdims=True)
        idx = np.arange(pr.shape[0], dtype=np.float64).reshape(-1, 1, 1)
        e = (pr * idx).sum(0)
        want_map = (pr * (np.abs(idx - e) <= 2)).sum(0)
        n_bad = assert_mismatches_on_boundary(probs[k][0].cpu().numpy(), want_map, expected_index_boundary_distance(pr), 1e-4,
                                              1e-5 * pr.shape[0], "Vis stage %d probability map" % (k + 1))
        assert n_bad <= 0.02 * want_map.size, n_bad
    # and the forward's photometric_confidence (the three maps, the coarse ones bilinearly up-sampled) against the reference
    # golden: a boundary pixel of a coarse map spreads over its up-sampled neighbourhood, hence a fraction and not a per-pixel rule
    s = {k: torch.from_numpy(g[k]).to(DEV) for k in ("depth_min", "depth_max")}
    p1 = torch.nn.functional.interpolate(probs[0].unsqueeze(1), scale_factor=4, mode="bilinear", align_corners=False)
    p2 = torch.nn.functional.interpolate(probs[1].unsqueeze(1), scale_factor=2, mode="bilinear", align_corners=False)
    conf = torch.cat([p1, p2, probs[2].unsqueeze(1)], 1).cpu().numpy()
    assert conf.shape == g["conf"].shape
    assert (np.abs(conf - g["conf"]) > 1e-3).mean() < 0.02, (np.abs(conf - g["conf"]) > 1e-3).mean()


def test_forward_api(golden):
    torch.manual_seed(0)
    net = Frontend()
    synth.randomize_norm_stats(net, seed=2)
    net = net.to(DEV).eval()
    s = {k: v.to(DEV) for k, v in synth.make_sample(2, 3, 64, 80, seed=0).items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], depth_nums=[8, 4, 4], interval_scales=[4, 2, 1])
    assert out["depth"].shape == (2, 32, 40)
    assert [tuple(d.shape) for d in out["depth_est_list"]] == [(2, 32, 40), (2, 16, 20), (2, 8, 10)]
    assert out["photometric_confidence"].shape == (2, 3, 32, 40)
    assert len(out["depth_pair_list"]) == 3 and len(out["depth_pair_list"][0]) == 2
    est, heads = out["depth_pair_list"][0][0]
    assert est.shape == (2, 1, 32, 40) and heads[0].shape == (2, 1, 32, 40)
    assert all(torch.isfinite(d).all() for d in out["depth_est_list"])


def test_graphed_cascade_equals_eager(golden):
    """Frontend.graphed: the three stages replayed as one CUDA graph give exactly the eager result, also on new inputs."""
    g = golden("vis")
    net = _load(g)
    feats, ref_cam, src_cams, dmin, interval = _inputs(g)
    args = (net.depth_nums, net.interval_scales)
    graphed = net.graphed(feats, ref_cam, src_cams, dmin, interval, *args)
    want, _, _ = net.depth_from_features(feats, ref_cam, src_cams, dmin, interval, *args)
    got, probs, pairs = graphed()
    assert all(torch.equal(a, b) for a, b in zip(got, want))
    feats2 = [[f.flip(2).contiguous() for f in fv] for fv in feats]
    want2, _, _ = net.depth_from_features(feats2, ref_cam, src_cams, dmin, interval, *args)
    got2, _, _ = graphed(feats2)
    assert all(torch.equal(a, b) for a, b in zip(got2, want2)) and not torch.equal(got2[2], want[2])


def test_featext_on_library_kernels_against_reference_golden(golden):
    """Row f1: FeatExt.run executes the extractor on K7 (stem), the K2 engines over one-plane volumes (U-Net) and the
    pointwise kernel (1x1 shortcuts); checked against the features the reference produced for the same image and weights
    (tests/golden/featext.npz) and against the plain PyTorch modules at a size with several tiles.  (Opt-in in forward:
    the cuDNN modules are still faster for these 64-128 channel 2-D layers.)"""
    g = golden("featext")
    net = Frontend().model.feat_ext
    net.load_state_dict({k[len("model.feat_ext."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("model.feat_ext.")},
                        strict=True)
    net = net.to(DEV).eval()
    with torch.no_grad():
        got = [f.permute(0, 3, 1, 2) for f in net.run(torch.from_numpy(g["img"]).to(DEV))]
    for k in range(3):
        assert got[k].shape == g["feat_s%d" % (k + 1)].shape
        assert rel_linf(got[k].cpu().numpy(), g["feat_s%d" % (k + 1)]) < 2e-5
    x = torch.rand(2, 3, 136, 200, device=DEV)
    with torch.no_grad():
        lib = [f.permute(0, 3, 1, 2) for f in net.run(x)]
    with torch.enable_grad():
        torch.backends.cudnn.allow_tf32, old = False, torch.backends.cudnn.allow_tf32
        plain = [t.detach() for t in net(x)]
        torch.backends.cudnn.allow_tf32 = old
    for a, b in zip(lib, plain):
        assert a.shape == b.shape and rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 1e-4


def test_featext_fused_inference_path_against_reference_golden(golden):
    """Row f1, the DEFAULT inference path of FeatExt: cuDNN convolutions with eval-mode BatchNorm folded into them and
    bias / ReLU / skip fused into the call -- against the features the reference produced (tests/golden/featext.npz) and
    against the plain modules (the reference's own execution) at another size."""
    g = golden("featext")
    net = Frontend().model.feat_ext
    net.load_state_dict({k[len("model.feat_ext."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("model.feat_ext.")},
                        strict=True)
    net = net.to(DEV).eval()
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # the reference golden is fp32 (CPU)
    try:
        with torch.no_grad():
            got = net(torch.from_numpy(g["img"]).to(DEV))
        for k in range(3):
            assert got[k].shape == g["feat_s%d" % (k + 1)].shape
            assert rel_linf(got[k].cpu().numpy(), g["feat_s%d" % (k + 1)]) < 2e-5
        x = torch.rand(2, 3, 136, 200, device=DEV)
        with torch.no_grad():
            fused = net(x)
        with torch.enable_grad():
            plain = [t.detach() for t in net(x)]
        for a, b in zip(fused, plain):
            assert a.shape == b.shape and rel_linf(a.cpu().numpy(), b.cpu().numpy()) < 2e-5
    finally:
        torch.backends.cudnn.allow_tf32 = old


def test_graphed_forward_equals_eager_forward():
    """`Frontend.graphed_forward`: the whole eval-mode forward (images -> dict, same kwargs) as ONE CUDA graph gives the eager
    forward's outputs, and follows new inputs copied into the captured buffers."""
    torch.manual_seed(0)
    net = Frontend()
    synth.randomize_norm_stats(net, seed=2)
    net = net.to(DEV).eval()
    kw = dict(depth_nums=[8, 4, 4], interval_scales=[4, 2, 1])
    s = {k: v.to(DEV) for k, v in synth.make_sample(1, 3, 64, 80, seed=0).items()}
    args = lambda d: (d["imgs"], d["K"], d["R"], d["t"], d["depth_min"], d["depth_max"])
    # (the cascade -- library kernels only -- replays bit for bit, test_graphed_cascade_equals_eager; the 2-D extractor in
    # front of it is cuDNN, whose algorithm choice may differ between the eager call and the capture: tolerance, not equality)
    close = lambda a, b: float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
    want = net(*args(s), **kw)
    g = net.graphed_forward(*args(s), **kw)
    got = g()
    assert close(got["depth"], want["depth"]) and close(got["photometric_confidence"], want["photometric_confidence"])
    s2 = {k: v.to(DEV) for k, v in synth.make_sample(1, 3, 64, 80, seed=5).items()}
    s2["depth_min"], s2["depth_max"] = s2["depth_min"] + 40, s2["depth_max"] + 90       # another sweep range: another depth map
    want2 = net(*args(s2), **kw)
    got2 = g(*args(s2))
    assert close(got2["depth"], want2["depth"]) and not close(want2["depth"], want["depth"])
