"""GPU: the drop-in CVP-MVSNet (wild_deep_mvs_b200.cvpmvsnet) -- against the reference-generated golden (2-level
pyramid, every seam of SURVEY.md 8-a14..a19), against the CPU oracle, and at BASELINE cfg4 size through
size-independent properties."""
import numpy as np
import pytest
import torch

from conftest import rel_linf

pytestmark = pytest.mark.gpu

from oracle import nets  # noqa: E402
from wild_deep_mvs_b200 import ops, synth  # noqa: E402
from wild_deep_mvs_b200.cvpmvsnet import Frontend, condition_intrinsics  # noqa: E402

DEV = "cuda:0"
DEPTH_TOL = 1e-3  # north_star: depth maps within 1e-3 relative L-inf of the reference PyTorch path


def _load(g):
    net = Frontend()
    sd = net.state_dict()
    for k, v in g.items():
        if k.startswith("model.cost_reg_refine."):
            assert k in sd, k          # the reference's key names are preserved
            sd[k] = torch.from_numpy(v)
    net.load_state_dict(sd, strict=True)
    return net.to(DEV).eval()


def _golden_inputs(g):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    nv, nl = 3, 2
    fps = [[ops.to_nhwc(t(g["fp_v%d_l%d" % (v, l)])) for l in range(nl)] for v in range(nv)]
    K, R, tt = t(g["K"]), t(g["R"]), t(g["t"])
    img_shape = (1, 3, g["fp_v0_l0"].shape[2], g["fp_v0_l0"].shape[3])
    shapes = [g["fp_v0_l%d" % l].shape for l in range(nl)]
    ref_in = condition_intrinsics(K[:, 0], img_shape, shapes)
    src_in = torch.stack([condition_intrinsics(K[:, v], img_shape, shapes) for v in (1, 2)]).permute(1, 0, 2, 3, 4)
    last = torch.tensor([0., 0., 0., 1.], device=DEV)
    E = torch.cat((torch.cat((R, tt), dim=3), last.view(1, 1, 1, 4).expand(1, nv, 1, 4)), dim=2)
    return fps, ref_in, src_in, E[:, 0], E[:, 1:], t(g["depth_min"][:, 0]), t(g["depth_max"][:, 0])


def test_pyramid_against_reference_and_oracle(golden):
    g = golden("cvp")
    net = _load(g)
    fps, ref_in, src_in, ref_ex, src_ex, dmin, dmax = _golden_inputs(g)
    with torch.no_grad():
        ests, conf, seams = net.model.depth_from_pyramids(fps[0], fps[1:], ref_in, src_in, ref_ex, src_ex, dmin, dmax)
    npy = lambda x: x.cpu().numpy()
    # every seam against the tensors the reference itself produced
    assert rel_linf(npy(seams["reg_out_coarse"]), g["reg_out_coarse"]) < 2e-4      # a14, a15, a18
    assert rel_linf(npy(ests[0]), g["depth_est_1"]) < 1e-4                          # a19 (coarse)
    assert rel_linf(npy(seams["hypos_l0"]), g["hypos_l0"]) < 1e-4                   # a17 + bicubic up-sampling
    assert rel_linf(npy(ests[1]), g["depth"]) < DEPTH_TOL                           # a16, a18, a19 (fine)
    assert (np.abs(npy(conf) - g["conf"][:, 0]) > 1e-3).mean() < 0.02
    # and against the CPU oracle on the same inputs
    ofps = [[g["fp_v%d_l%d" % (v, l)][0] for l in range(2)] for v in range(3)]
    Es = []
    for v in range(3):
        E = np.eye(4, dtype=np.float32)
        E[:3, :3], E[:3, 3:] = g["R"][0, v], g["t"][0, v]
        Es.append(E)
    want = nets.cvp_from_features(g, ofps, [g["K"][0, v] for v in range(3)], Es, g["depth_min"][0, 0], g["depth_max"][0, 0])
    assert rel_linf(npy(seams["reg_out_coarse"])[0], want["seams"]["reg_out_coarse"]) < 2e-4
    assert rel_linf(npy(ests[1])[0], want["depth"]) < DEPTH_TOL


def test_regulariser_seam_keeps_the_reference_signature(golden):
    g = golden("cvp")
    net = _load(g)
    x = torch.from_numpy(g["reg_in_l0"]).to(DEV)
    out = net.model.cost_reg_refine(x)           # [B,16,D,H,W] -> [B,D,H,W]
    assert out.shape == g["reg_out_l0"].shape
    assert rel_linf(out.cpu().numpy(), g["reg_out_l0"]) < 5e-5                      # a18


def test_forward_api(golden):
    g = golden("cvp")
    net = _load(g)
    s = synth.make_sample(1, 3, 32, 48, seed=0)
    s = {k: v.to(DEV) for k, v in s.items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=2)
    assert out["depth"].shape == (1, 32, 48) and out["photometric_confidence"].shape == (1, 1, 32, 48)
    assert [tuple(d.shape) for d in out["depth_est_list"]] == [(1, 32, 48), (1, 16, 24)] and out["depth_pair_list"] == []
    assert torch.isfinite(out["depth"]).all()
    # list input and a non-zero reference frame
    out2 = net(list(torch.unbind(s["imgs"], 1)), s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"],
               reference_frame=1, nscale=2)
    assert out2["depth"].shape == (1, 32, 48) and torch.isfinite(out2["depth"]).all()
    out3 = net.train()(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=2)   # tests/test_gpu_backward.py
    assert out3["depth"].requires_grad and out3["depth"].shape == (1, 32, 48)


def test_full_size_properties():
    """BASELINE cfg4: 1+4 views, 1600x1184, 5 pyramid levels.  No oracle at this size (the CPU path needs ~40 s and
    several GB): depth stays inside the swept range at every level, the run is deterministic, and the level-0 map is
    consistent with its own coarser levels."""
    torch.manual_seed(0)
    net = Frontend()
    synth.randomize_norm_stats(net, seed=3)
    synth.scale_param(net.model.cost_reg_refine.prob0.weight, 30.0)
    net = net.to(DEV).eval()
    s = synth.make_sample(1, 5, 1184, 1600, seed=0)
    s = {k: v.to(DEV) for k, v in s.items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=5)
    shapes = [tuple(d.shape) for d in out["depth_est_list"]]
    assert shapes == [(1, 1184, 1600), (1, 592, 800), (1, 296, 400), (1, 148, 200), (1, 74, 100)]
    lo, hi = 425.0, 905.0
    margin = 0.25 * (hi - lo)
    for d in out["depth_est_list"]:
        assert torch.isfinite(d).all()
        assert (d > lo - margin).all() and (d < hi + margin).all()
    conf = out["photometric_confidence"]
    assert conf.shape == (1, 1, 1184, 1600) and (conf >= 0).all() and (conf <= 1 + 1e-5).all()
    out_b = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=5)
    assert torch.equal(out["depth"], out_b["depth"])
    coarse_up = torch.nn.functional.interpolate(out["depth_est_list"][1][None], scale_factor=2, mode="bilinear")[0]
    assert (out["depth"] - coarse_up).abs().mean() < 0.05 * (hi - lo)


def test_k5_depth_hypotheses_against_oracle():
    """K5 (calDepthHypo): per-pixel fp64 solve + device median vs the numpy restatement of modules.py:131-226, incl. a
    batch of two different cameras, invalid pixels (depth behind the source camera) and the degenerate fallback."""
    from wild_deep_mvs_b200.cvpmvsnet import cal_depth_hypo
    rng = np.random.default_rng(4)
    B, H, W = 2, 37, 53
    K, R, tt, dmin, dmax = synth.make_cameras(B, 3, 4 * H, 4 * W)
    K = K.clone()
    K[:, :, :2] /= 4
    R[1, 1] = R[0, 2]                     # second sample: a different relative pose
    tt[1, 1] = tt[0, 2] * 1.5
    last = torch.tensor([0., 0., 0., 1.])
    E = torch.cat((torch.cat((R, tt), dim=3), last.view(1, 1, 1, 4).expand(B, 3, 1, 4)), dim=2)
    depth = (425 + 480 * rng.random((B, H, W))).astype(np.float32)
    depth[0, :3] = -50.0                  # behind the cameras: fails the z > 1e-8 test, excluded from the median
    d = torch.from_numpy(depth).to(DEV)
    got = cal_depth_hypo(d, K[:, 0].to(DEV), K[:, 1:].to(DEV), E[:, 0].to(DEV), E[:, 1:].to(DEV), dmin[:, 0].to(DEV), dmax[:, 0].to(DEV))
    assert got.dtype == torch.float32 and got.shape == (B, 8, H, W)
    for b in range(B):
        want = nets.cvp_depth_hypos(depth[b], K[b, 0].numpy(), K[b, 1].numpy(), E[b, 0].numpy(), E[b, 1].numpy(),
                                    float(dmin[b, 0]), float(dmax[b, 0]))
        assert rel_linf(got[b].cpu().numpy(), want) < 1e-6
        step = float(got[b, 5, 10, 10] - got[b, 4, 10, 10])
        assert 0.05 < step < 500.0        # a plausible interval: about one pixel of parallax on a 53-pixel-wide map
    # the raw per-pixel values: +inf exactly where the reference's validity test fails
    delta = ops.cvp_depth_delta(d, K[:, 0].to(DEV), K[:, 1].to(DEV), E[:, 0].to(DEV), E[:, 1].to(DEV)).view(B, H, W)
    assert torch.isinf(delta[0, :3]).all() and torch.isfinite(delta[0, 3:]).all() and torch.isfinite(delta[1]).all()
    # identical cameras: no parallax, every pixel invalid -> the (max - min) / 128 fallback (modules.py:211-213)
    got = cal_depth_hypo(d[:1], K[:1, 0].to(DEV), K[:1, :1].to(DEV), E[:1, 0].to(DEV), E[:1, :1].to(DEV), dmin[:1, 0].to(DEV), dmax[:1, 0].to(DEV))
    step = (got[0, 5] - got[0, 4])[5:].cpu().numpy()
    assert np.allclose(step, (905.0 - 425.0) / 128, rtol=1e-4)


def test_bias_act_kernel_and_fused_pyramid():
    """K7b (mvsb200_bias_act): y = act(y * scale + bias (+ residual)) in place over channels-last maps -- against the torch
    expression, bit for bit (one fma / add per element, same order); and the CVP FeaturePyramid through it (cuDNN
    convolution without bias + ONE fused pass) against the plain nn.Conv2d + nn.LeakyReLU modules."""
    from wild_deep_mvs_b200 import ops
    from wild_deep_mvs_b200.cvpmvsnet import FeaturePyramid
    torch.manual_seed(0)
    y = torch.randn(2, 16, 37, 53, device=DEV).contiguous(memory_format=torch.channels_last)
    b, s = torch.randn(16, device=DEV), torch.rand(16, device=DEV) + 0.5
    r = torch.randn_like(y)
    want = torch.nn.functional.leaky_relu(y + b.view(1, -1, 1, 1), 0.1)
    assert torch.equal(ops.bias_act_(y.clone(memory_format=torch.channels_last), b, slope=0.1), want)
    want = torch.relu(torch.addcmul(b.view(1, -1, 1, 1), y, s.view(1, -1, 1, 1)) + r)      # fma(y, s, b) + r
    got = ops.bias_act_(y.clone(memory_format=torch.channels_last), b, scale=s, residual=r, slope=0.0)
    assert (got - want).abs().max() <= 1e-6 * want.abs().max()
    yn = y.permute(0, 2, 3, 1).contiguous()
    assert torch.equal(ops.bias_act_(yn.clone(), b, slope=1.0, nhwc=True), yn + b)
    with pytest.raises(Exception):
        ops.bias_act_(torch.randn(1, 6, 4, 4, device=DEV).contiguous(memory_format=torch.channels_last), torch.zeros(6, device=DEV))
    net = FeaturePyramid().to(DEV).eval()
    img = torch.rand(2, 3, 72, 104, device=DEV)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            fused = net(img, 3)
        with torch.enable_grad():
            plain = [t.detach() for t in net(img, 3)]
    finally:
        torch.backends.cudnn.allow_tf32 = old
    for a, c in zip(fused, plain):
        assert a.shape == c.shape and (a - c).abs().max() <= 2e-6 * c.abs().max()


def test_graphed_forward_equals_eager_forward():
    """`Frontend.graphed_forward`: the whole CVP-MVSNet forward (pyramid, K5 median interval, every level) captured into ONE
    CUDA graph -- legal only because the forward has no host read or host -> device copy left -- equals the eager forward."""
    torch.manual_seed(0)
    net = Frontend()
    synth.randomize_norm_stats(net, seed=3)
    net = net.to(DEV).eval()
    s = {k: v.to(DEV) for k, v in synth.make_sample(1, 3, 128, 160, seed=0).items()}
    a = (s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    want = net(*a, nscale=2)
    g = net.graphed_forward(*a, nscale=2)
    got = g()
    close = lambda x, y: float((x - y).abs().max()) <= 1e-5 * float(y.abs().max())     # (the pyramid's convolutions are cuDNN calls)
    assert close(got["depth"], want["depth"])
    assert ((got["photometric_confidence"] - want["photometric_confidence"]).abs() > 1e-4).float().mean() < 5e-3
