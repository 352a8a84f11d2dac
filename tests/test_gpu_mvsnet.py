"""GPU: the drop-in MVSNet module (wild_deep_mvs_b200.mvsnet) end to end -- against the reference-generated
goldens at small size, and against the torch port / size-independent properties at BASELINE cfg2 size."""
import numpy as np
import pytest
import torch

from conftest import assert_mismatches_on_boundary, expected_index_boundary_distance, rel_linf

pytestmark = pytest.mark.gpu

from wild_deep_mvs_b200 import _lib as L  # noqa: E402
from wild_deep_mvs_b200 import ops, synth  # noqa: E402
from wild_deep_mvs_b200.mvsnet import MVSNet, build_proj_matrices  # noqa: E402

DEV = "cuda:0"
DEPTH_TOL = 1e-3  # north_star: depth maps within 1e-3 relative L-inf of the reference PyTorch path


def _load(g, agg):
    net = MVSNet(agg)
    sd = net.state_dict()
    for k, v in g.items():
        if k.startswith("cost_regularization."):
            sd[k] = torch.from_numpy(v)
    if agg == "softmin":
        sd["temp"] = torch.from_numpy(g["temp"])
    net.load_state_dict(sd, strict=True)
    net.num_depth = g["depth_values"].shape[1]
    return net.to(DEV).eval()


@pytest.mark.parametrize("agg", ["variance", "softmin"])
def test_depth_from_features_matches_reference(golden, agg):
    g = golden("mvsnet_" + agg)
    net = _load(g, agg)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    feats = [ops.to_nhwc(t(g["feat%d" % i])) for i in range(3)]
    projs = [t(g["proj"][:, i]) for i in range(3)]
    depth, conf = net.depth_from_features(feats, projs, t(g["depth_values"]))
    assert rel_linf(depth.cpu().numpy(), g["depth"]) < DEPTH_TOL
    assert rel_linf(depth.cpu().numpy(), g["depth"]) < 1e-4   # what fp32 actually achieves
    # confidence = 4 bins around floor(expected index): differs from the reference only where that index sits on an integer
    D = g["prob"].shape[1]
    n_bad = assert_mismatches_on_boundary(conf.cpu().numpy(), g["conf"], expected_index_boundary_distance(g["prob"]), 1e-4, 1e-5 * D,
                                          "MVSNet photometric confidence (%s)" % agg)
    assert n_bad <= 0.01 * conf.numel(), n_bad
    # the seam methods keep the reference's signatures and layouts
    vol = net.build_cost_volume(t(g["feat0"]), [t(g["feat1"]), t(g["feat2"])], projs[0], projs[1:], t(g["depth_values"]))
    assert vol.shape == g["cost_volume"].shape
    assert vol.requires_grad == (agg == "softmin")   # like the reference: `temp` is a Parameter, grad mode is on here
    assert rel_linf(vol.detach().cpu().numpy(), g["cost_volume"]) < 1e-4
    reg = net.cost_regularization(t(g["cost_volume"]))
    assert rel_linf(reg[:, 0].cpu().numpy(), g["cost_reg"]) < 2e-5


def test_forward_api_and_state_dict_names():
    net = MVSNet("softmin")
    names = set(net.state_dict().keys())
    for k in ("temp", "feature.conv0.conv.weight", "feature.conv6.bn.running_var", "feature.feature.bias",
              "cost_regularization.conv0.conv.weight", "cost_regularization.conv6.bn.running_mean",
              "cost_regularization.conv7.0.weight", "cost_regularization.conv11.1.bias",
              "cost_regularization.prob.weight", "cost_regularization.prob.bias"):
        assert k in names, k
    net = net.to(DEV).eval()
    net.num_depth = 16
    s = synth.make_sample(2, 3, 64, 96, seed=1)
    s = {k: v.to(DEV) for k, v in s.items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    assert out["depth"].shape == (2, 16, 24) and out["photometric_confidence"].shape == (2, 16, 24)
    assert out["depth_est_list"][0] is out["depth"] and out["depth_pair_list"] == []
    assert torch.isfinite(out["depth"]).all()
    assert (out["depth"] >= 425).all() and (out["depth"] <= 905).all()
    # list input with a source view of a different size (MegaDepth/YFCC test mode)
    imgs = [s["imgs"][:, 0], s["imgs"][:, 1, :, :56, :80].contiguous(), s["imgs"][:, 2]]
    out2 = net(imgs, s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    assert out2["depth"].shape == (2, 16, 24) and torch.isfinite(out2["depth"]).all()
    out3 = net.train()(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])   # training: tests/test_gpu_backward.py
    assert out3["depth"].requires_grad and out3["depth"].shape == (2, 16, 24)
    with pytest.raises(NotImplementedError):
        MVSNet("nope").to(DEV).eval()(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])


def _cfg2(agg="variance", V=5, D=192, h=128, w=160, seed=0):
    torch.manual_seed(seed)
    net = MVSNet(agg)
    synth.randomize_norm_stats(net, seed=1)
    synth.scale_param(net.cost_regularization.prob.weight, 40.0)
    net = net.to(DEV).eval()
    net.num_depth = D
    feats = [f.to(DEV) for f in synth.make_features(1, V, 32, h, w, seed=seed)]
    K, R, t, dmin, dmax = synth.make_cameras(1, V, 4 * h, 4 * w)
    K = K.clone()
    K[:, :, :2] /= 4
    projs = list(torch.unbind(build_proj_matrices(K, R, t).to(DEV), 1))
    depth = (dmin[:, :1] + (dmax[:, :1] - dmin[:, :1]) / (D - 1) * torch.arange(D).view(1, -1)).to(DEV)
    return net, feats, projs, depth


@pytest.mark.parametrize("agg", ["variance", "softmin"])
def test_full_size_against_torch_port(agg):
    """BASELINE cfg2 shapes (1+4 views, 128x160 features, D=192): every seam against the ATen port run on the GPU
    in full fp32 (TF32 off)."""
    from oracle import torch_port as tp
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net, feats, projs, depth = _cfg2(agg)
    sd = {k: v for k, v in net.state_dict().items()}
    temp = net.temp.detach() if agg == "softmin" else None
    with torch.no_grad():
        vol = net.cost_volume_cl([ops.to_nhwc(f) for f in feats][0], [ops.to_nhwc(f) for f in feats[1:]], projs[0], projs[1:], depth)
        want = tp.mvsnet_cost_volume(feats[0], feats[1:], projs[0], projs[1:], depth, agg, temp)
        assert rel_linf(ops.as_ncdhw(vol).cpu().numpy(), want.cpu().numpy()) < 2e-4
        score = net.cost_regularization.run(vol)
        want_reg = tp.mvsnet_costreg(sd, ops.as_ncdhw(vol).contiguous()).squeeze(1)
        assert rel_linf(score.cpu().numpy(), want_reg.cpu().numpy()) < 1e-4
        out = ops.depth_regress(score, depth, conf_mode=L.CONF_SUM4)
        want_depth, want_conf = tp.mvsnet_head(want_reg, depth)
        assert rel_linf(out["depth"].cpu().numpy(), want_depth.cpu().numpy()) < DEPTH_TOL
        assert ((out["conf"] - want_conf).abs() > 1e-3).float().mean().item() < 0.01
        # whole path in one call
        d2, c2 = net.depth_from_features([ops.to_nhwc(f) for f in feats], projs, depth)
        assert torch.equal(d2, out["depth"])


def test_full_size_properties():
    net, feats, projs, depth = _cfg2("variance")
    nh = [ops.to_nhwc(f) for f in feats]
    # (1) identical views under the identity relative pose have zero variance
    vol = net.cost_volume_cl(nh[0], [nh[0]] * 4, projs[0], [projs[0]] * 4, depth)
    assert vol.abs().max().item() < 1e-4 * (feats[0] ** 2).max().item()
    # (2) determinism: the path has no atomics
    d1, c1 = net.depth_from_features(nh, projs, depth)
    d2, c2 = net.depth_from_features(nh, projs, depth)
    assert torch.equal(d1, d2) and torch.equal(c1, c2)
    # (3) regression outputs live where they must
    assert (d1 >= depth.min()).all() and (d1 <= depth.max()).all()
    assert (c1 >= 0).all() and (c1 <= 1 + 1e-5).all()
    # (4) linearity of the raw convolution under exact power-of-two scaling
    layer = ops.PackedConv(net.cost_regularization.conv0.conv.weight)
    vol = net.cost_volume_cl(nh[0], nh[1:], projs[0], projs[1:], depth)
    assert torch.equal(ops.conv3d(vol * 2, layer), ops.conv3d(vol, layer) * 2)
    # (5) permuting the source views leaves the variance volume unchanged up to fp32 re-association
    vol_p = net.cost_volume_cl(nh[0], nh[:0:-1], projs[0], projs[:0:-1], depth)
    assert rel_linf(vol_p.cpu().numpy(), vol.cpu().numpy()) < 1e-5


def test_cuda_graph_replay_equals_eager_launches(golden):
    """GraphedHotPath captures the ~14 launches of a step once; replays on new inputs must equal eager calls bit for bit."""
    g = golden("mvsnet_variance")
    net = _load(g, "variance")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    feats = [ops.to_nhwc(t(g["feat%d" % i])) for i in range(3)]
    projs = [t(g["proj"][:, i]) for i in range(3)]
    dv = t(g["depth_values"])
    graphed = net.graphed(feats, projs, dv)
    d_eager, c_eager = net.depth_from_features(feats, projs, dv)
    d_graph, c_graph = graphed()
    assert torch.equal(d_graph, d_eager) and torch.equal(c_graph, c_eager)
    feats2 = [f.flip(1).contiguous() for f in feats]          # new inputs through the same graph
    d2_eager, _ = net.depth_from_features(feats2, projs, dv)
    d2_graph, _ = graphed(feats2, projs, dv)
    assert torch.equal(d2_graph, d2_eager) and not torch.equal(d2_graph, d_eager)


def test_streamed_pinned_host_samples_equal_eager(golden):
    """StreamedHotPath: pinned host inputs, H2D on a copy stream overlapping the previous sample's kernels, two graph
    slots round-robin, results copied back to pinned host buffers -- every sample must equal the eager device path."""
    g = golden("mvsnet_variance")
    net = _load(g, "variance")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    feats = [ops.to_nhwc(t(g["feat%d" % i])) for i in range(3)]
    projs = [t(g["proj"][:, i]) for i in range(3)]
    dv = t(g["depth_values"])
    streamed = net.streamed(feats, projs, dv)
    samples = [feats, [f.flip(1).contiguous() for f in feats], [f.flip(2).contiguous() for f in feats],
               [(f * 0.5).contiguous() for f in feats], feats]
    hp = [p.cpu().pin_memory() for p in projs]
    hd = dv.cpu().pin_memory()
    want = [net.depth_from_features(s, projs, dv) for s in samples]
    got = []
    for s in samples:   # submitted back to back: slot reuse after two submissions
        hs = [f.cpu().pin_memory() for f in s]
        d, c, ev = streamed.submit(hs, hp, hd)
        ev.synchronize()
        got.append((d.clone(), c.clone()))
    for (d, c), (wd, wc) in zip(got, want):
        assert torch.equal(d, wd.cpu()) and torch.equal(c, wc.cpu())
    assert not torch.equal(got[0][0], got[1][0])


def test_feature_net_k7_against_reference_golden_and_module_path(golden):
    """FeatureNet (row f1): under no_grad the drop-in runs its eight layers through K7 (mvsb200_conv2d).  Checked against
    the features the reference produced for the same image and weights (tests/golden/featurenet.npz, generated by
    tests/golden/make_golden_features.py), against the CPU oracle, and against the plain PyTorch modules at a size
    with several tiles per image."""
    from oracle import nets
    g = golden("featurenet")
    net = MVSNet("variance").feature
    net.load_state_dict({k[len("feature."):]: torch.from_numpy(v) for k, v in g.items() if k.startswith("feature.")}, strict=True)
    net = net.to(DEV).eval()
    img = torch.from_numpy(g["img"]).to(DEV)
    with torch.no_grad():
        feat = net(img)
    assert feat.shape == g["feat"].shape
    assert rel_linf(feat.cpu().numpy(), g["feat"]) < 1e-5                                    # vs the reference itself
    assert rel_linf(feat[0].cpu().numpy(), nets.mvsnet_featurenet(g, g["img"][0])) < 1e-5   # vs the oracle
    synth.randomize_norm_stats(net, seed=5)
    x = torch.rand(3, 3, 200, 328, device=DEV)
    with torch.no_grad():
        k7 = net(x)
    with torch.enable_grad():
        torch.backends.cudnn.allow_tf32, old = False, torch.backends.cudnn.allow_tf32
        plain = net(x).detach()
        torch.backends.cudnn.allow_tf32 = old
    assert k7.shape == plain.shape == (3, 32, 50, 82)
    assert rel_linf(k7.cpu().numpy(), plain.cpu().numpy()) < 1e-4


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_net_on_second_gpu_while_first_is_current(golden):
    """Every library call runs on the OPERANDS' device (ops._on_operand_device), not on whatever device is current:
    a net moved to cuda:1 gives the same depth map as on cuda:0 while cuda:0 stays the current device."""
    from wild_deep_mvs_b200 import _lib as L
    assert torch.cuda.current_device() == 0
    g = golden("mvsnet_variance")
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        net = MVSNet("variance")
        sd = net.state_dict()
        for k, v in g.items():
            if k.startswith("cost_regularization."):
                sd[k] = torch.from_numpy(v)
        net.load_state_dict(sd)
        net.num_depth = g["depth_values"].shape[1]
        net = net.to(dev).eval()
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        feats = [ops.to_nhwc(t(g["feat%d" % i])) for i in range(3)]
        projs = [t(g["proj"][:, i]) for i in range(3)]
        depth, conf = net.depth_from_features(feats, projs, t(g["depth_values"]))
        assert depth.device == torch.device(dev)
        outs.append(depth.cpu())
    assert torch.cuda.current_device() == 0
    assert torch.equal(outs[0], outs[1])
    with pytest.raises(L.Mvsb200Error):
        ops.depth_regress(torch.zeros(1, 8, 8, 8, device="cuda:0"), torch.zeros(1, 8, device="cuda:1"))
