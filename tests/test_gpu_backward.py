"""GPU: the backward kernels (row f2 of SURVEY.md 8) through the C ABI -- K1 backward (mvsb200_build_cost_volume_backward)
and K3 backward (mvsb200_depth_regress_backward) against gradients autograd produced through the UNMODIFIED reference
(tests/golden/backward.npz, mvsnet_train.npz), against autograd through the torch port at a BASELINE size, and through a
size-independent property: for the aggregations that are polynomial in the features (variance: quadratic, group
correlation: bilinear) the central difference of the FORWARD kernel along a random direction equals <gradient, direction>.

Tolerances: relative L-inf 1e-4 on gradients (fp32, atomics in arbitrary order; measured ~1e-6)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_linf

pytestmark = pytest.mark.gpu

from oracle import torch_port as tp  # noqa: E402
from wild_deep_mvs_b200 import _lib as L  # noqa: E402
from wild_deep_mvs_b200 import ops, synth  # noqa: E402
from wild_deep_mvs_b200.mvsnet import MVSNet, build_proj_matrices  # noqa: E402

DEV = "cuda:0"
GRAD_TOL = 1e-4


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=DEV)


def nhwc(a):   # numpy [B,C,H,W] -> cuda [B,H,W,C]
    return cu(np.transpose(a, (0, 2, 3, 1)))


def nchw(t):   # cuda [B,H,W,C] -> numpy [B,C,H,W]
    return t.permute(0, 3, 1, 2).contiguous().cpu().numpy()


@pytest.mark.parametrize("agg", ["variance", "softmin"])
def test_k1_backward_against_reference_gradients(golden, agg):
    g, b = golden("mvsnet_" + agg), golden("backward")
    D = g["depth_values"].shape[1]
    warp = ops.mvs_relative_proj(cu(g["proj"][:, 0]), cu(g["proj"][:, 1:]))
    temp = cu(g["temp"]) if agg == "softmin" else None
    feats = [nhwc(g["feat%d" % i]) for i in range(3)]
    G = cu(np.transpose(b[agg + "_G"], (0, 2, 3, 4, 1)))           # reference layout [B,C,D,H,W] -> [B,D,H,W,C]
    code = L.AGG_VARIANCE if agg == "variance" else L.AGG_SOFTMIN
    g_ref, g_srcs, g_temp = ops.build_cost_volume_backward(G, feats[0], feats[1:], warp, cu(g["depth_values"]), D, L.GEOM_MVS,
                                                           code, temp=temp)
    assert rel_linf(nchw(g_ref), b[agg + "_gfeat0"]) < GRAD_TOL
    for s in range(2):
        assert rel_linf(nchw(g_srcs[s]), b["%s_gfeat%d" % (agg, s + 1)]) < GRAD_TOL
    if agg == "softmin":
        assert rel_linf(g_temp.cpu().numpy(), b["softmin_gtemp"]) < GRAD_TOL
    else:
        assert g_temp is None
        # CVP's rounding order of the same function has the same derivative
        g2 = ops.build_cost_volume_backward(G, feats[0], feats[1:], warp, cu(g["depth_values"]), D, L.GEOM_MVS, L.AGG_VARIANCE_MEAN)
        assert rel_linf(nchw(g2[0]), b[agg + "_gfeat0"]) < GRAD_TOL


def test_k1_backward_groupcorr_against_reference_gradients(golden):
    g, b = golden("vis"), golden("backward")
    warp = ops.vis_homography_params(cu(g["ref_cam"]), cu(np.stack([g["src_cam1"], g["src_cam2"]], 1)), 1.0 / 8)
    feats = [nhwc(g["feat_v%d_s1" % v]) for v in range(3)]
    interval = np.float32((g["depth_max"][0, 0] - g["depth_min"][0, 0]) / np.float32(128)) * np.float32(4)
    G = cu(np.stack([np.transpose(b["vis_G%d" % v], (0, 2, 3, 4, 1)) for v in range(2)]))      # [S,B,D,H,W,8]
    g_ref, g_srcs, _ = ops.build_cost_volume_backward(G, feats[0], feats[1:], warp, cu(g["depth_min"][:, 0]), 8, L.GEOM_VIS,
                                                      L.AGG_GROUPCORR, interval=cu(np.array([interval])), groups=8)
    assert rel_linf(nchw(g_ref), b["vis_gfeat0"]) < GRAD_TOL
    for s in range(2):
        assert rel_linf(nchw(g_srcs[s]), b["vis_gfeat%d" % (s + 1)]) < GRAD_TOL


CASES = [
    # C, agg, geom, per-pixel hypotheses, S
    (16, L.AGG_VARIANCE_MEAN, L.GEOM_MVS, True, 3),      # CVP proj_cost flavour: ragged sources, depth volume
    (32, L.AGG_VARIANCE, L.GEOM_MVS, False, 4),
    (8, L.AGG_VARIANCE, L.GEOM_MVS, False, 1),
    (32, L.AGG_GROUPCORR, L.GEOM_VIS, True, 2),          # Vis stages 2/3: per-pixel depth start
    (16, L.AGG_GROUPCORR, L.GEOM_MVS, False, 2),
]


@pytest.mark.parametrize("C,agg,geom,per_pixel,S", CASES)
def test_k1_backward_is_the_derivative_of_the_forward_kernel(golden, C, agg, geom, per_pixel, S):
    gen = torch.Generator(device="cpu").manual_seed(C + 10 * agg + S)
    B, H, W, D = 2, 19, 27, 6                                     # nothing a multiple of the tile, D not of the chunk
    rnd = lambda *s: torch.randn(*s, generator=gen).to(DEV)
    ref = rnd(B, H, W, C)
    srcs = [rnd(B, H - 2 * s, W + 3 * s, C) for s in range(S)]
    if geom == L.GEOM_MVS:
        K, R, t, _, _ = synth.make_cameras(B, S + 1, 4 * H, 4 * W)
        K = K.clone()
        K[:, :, :2] /= 4
        proj = build_proj_matrices(K, R, t).to(DEV)
        warp = ops.mvs_relative_proj(proj[:, 0].contiguous(), proj[:, 1:].contiguous())
    else:
        v = golden("vis")
        cams = np.stack([v["src_cam1"], v["src_cam2"]], 1)[:, :S]
        warp = ops.vis_homography_params(cu(np.repeat(v["ref_cam"], B, 0)), cu(np.repeat(cams, B, 0)), 1.0 / 8)
    if geom == L.GEOM_VIS or agg == L.AGG_GROUPCORR:
        depth = (450 + 200 * torch.rand(B, H, W, generator=gen)).to(DEV) if per_pixel else torch.tensor([450.0, 500.0], device=DEV)
        interval = torch.tensor([30.0, 22.0], device=DEV)
    else:
        depth = (425 + 480 * torch.rand(*((B, D, H, W) if per_pixel else (B, D)), generator=gen)).to(DEV)
        interval = None
    fwd = lambda r, ss: ops.build_cost_volume(r, ss, warp, depth, D, geom, agg, interval=interval, groups=C // 4)
    out = fwd(ref, srcs)
    G = rnd(*out.shape)
    g_ref, g_srcs, _ = ops.build_cost_volume_backward(G, ref, srcs, warp, depth, D, geom, agg, interval=interval, groups=C // 4)
    assert all(torch.isfinite(x).all() for x in [g_ref] + g_srcs)
    # <grad, direction> against the central difference of sum(G * forward): exact for polynomials of degree <= 2
    d_ref, d_srcs = rnd(*ref.shape), [rnd(*s.shape) for s in srcs]
    eps = 0.5
    lp = (G.double() * fwd(ref + eps * d_ref, [s + eps * d for s, d in zip(srcs, d_srcs)]).double()).sum()
    lm = (G.double() * fwd(ref - eps * d_ref, [s - eps * d for s, d in zip(srcs, d_srcs)]).double()).sum()
    want = float((lp - lm) / (2 * eps))
    got = float((g_ref.double() * d_ref.double()).sum() + sum((g.double() * d.double()).sum() for g, d in zip(g_srcs, d_srcs)))
    scale = float((G.abs().double() * out.abs().double()).sum())
    assert abs(got - want) < 2e-5 * scale, (got, want, scale)
    # and per input: the gradient of a source nobody samples from the given direction is untouched by the others
    lp = (G.double() * fwd(ref + eps * d_ref, srcs).double()).sum()
    lm = (G.double() * fwd(ref - eps * d_ref, srcs).double()).sum()
    assert abs(float((g_ref.double() * d_ref.double()).sum()) - float((lp - lm) / (2 * eps))) < 2e-5 * scale


def test_k1_backward_softmin_at_cfg1_size_against_the_torch_port():
    """BASELINE cfg1 (MVSNet-s: 1+2 views, 32ch 128x160, D=48): autograd through the port's training-flavour graph
    (3 x 126 MB of warped volumes and their squares kept alive) vs one backward kernel."""
    gen = torch.Generator().manual_seed(0)
    B, C, H, W, D, S = 1, 32, 128, 160, 48, 2
    feats = [(0.5 * torch.randn(B, C, H, W, generator=gen)).to(DEV).requires_grad_(True) for _ in range(S + 1)]
    K, R, t, dmin, dmax = synth.make_cameras(B, S + 1, 4 * H, 4 * W)
    K = K.clone()
    K[:, :, :2] /= 4
    proj = build_proj_matrices(K, R, t).to(DEV)
    dv = (dmin[:, :1] + (dmax[:, :1] - dmin[:, :1]) / (D - 1) * torch.arange(D).view(1, -1)).to(DEV)
    temp = torch.tensor([0.37], device=DEV, requires_grad=True)
    vol = tp.mvsnet_cost_volume_train(feats[0], feats[1:], proj[:, 0], [proj[:, 1], proj[:, 2]], dv, "softmin", temp)
    G = torch.randn(vol.shape, generator=gen).to(DEV)
    (vol * G).sum().backward()
    warp = ops.mvs_relative_proj(proj[:, 0].contiguous(), proj[:, 1:].contiguous())
    f = [ops.to_nhwc(x.detach()) for x in feats]
    g_ref, g_srcs, g_temp = ops.build_cost_volume_backward(G.permute(0, 2, 3, 4, 1).contiguous(), f[0], f[1:], warp, dv, D,
                                                           L.GEOM_MVS, L.AGG_SOFTMIN, temp=temp.detach())
    assert rel_linf(nchw(g_ref), feats[0].grad.cpu().numpy()) < GRAD_TOL
    for s in range(S):
        assert rel_linf(nchw(g_srcs[s]), feats[s + 1].grad.cpu().numpy()) < GRAD_TOL
    assert rel_linf(g_temp.cpu().numpy(), temp.grad.cpu().numpy()) < 1e-3      # one scalar summed over 31 M terms


def test_cost_volume_autograd_function():
    """ops.cost_volume: torch.autograd sees K1 forward + backward as one op; inputs that need no gradient get None."""
    gen = torch.Generator().manual_seed(2)
    B, C, H, W, D = 1, 32, 16, 24, 8
    ref = torch.randn(B, H, W, C, generator=gen).to(DEV).requires_grad_(True)
    srcs = [torch.randn(B, H, W, C, generator=gen).to(DEV).requires_grad_(i == 0) for i in range(2)]
    K, R, t, dmin, dmax = synth.make_cameras(B, 3, 4 * H, 4 * W)
    K = K.clone()
    K[:, :, :2] /= 4
    proj = build_proj_matrices(K, R, t).to(DEV)
    warp = ops.mvs_relative_proj(proj[:, 0].contiguous(), proj[:, 1:].contiguous())
    dv = torch.linspace(425, 905, D, device=DEV).view(1, D)
    temp = torch.tensor([0.2], device=DEV, requires_grad=True)
    vol = ops.cost_volume(ref, srcs, warp, dv, D, L.GEOM_MVS, L.AGG_SOFTMIN, temp=temp)
    assert vol.requires_grad and vol.shape == (B, D, H, W, C)
    assert torch.equal(vol.detach(), ops.build_cost_volume(ref.detach(), [s.detach() for s in srcs], warp, dv, D, L.GEOM_MVS,
                                                           L.AGG_SOFTMIN, temp=temp.detach()))
    vol.square().sum().backward()
    assert ref.grad is not None and srcs[0].grad is not None and srcs[1].grad is None and temp.grad.shape == (1,)
    assert ref.grad.abs().max() > 0 and srcs[0].grad.abs().max() > 0


def test_k3_backward_against_reference_gradient_and_autograd(golden):
    b = golden("backward")
    score = cu(b["reg_score"])
    g = ops.depth_regress_backward(cu(b["reg_Gd"]), score, cu(b["reg_dvals"]))
    assert rel_linf(g.cpu().numpy(), b["reg_gscore"]) < GRAD_TOL
    # the autograd op, the three other hypothesis layouts, D not a multiple of the 8 depth lanes
    gen = torch.Generator().manual_seed(1)
    B, D, H, W = 2, 13, 7, 45
    sc = (3 * torch.randn(B, D, H, W, generator=gen)).to(DEV)
    vol = (425 + 480 * torch.rand(B, D, H, W, generator=gen)).to(DEV)
    start, start_map = torch.tensor([430.0, 500.0], device=DEV), (450 + 50 * torch.rand(B, H, W, generator=gen)).to(DEV)
    interval = torch.tensor([2.5, 4.0], device=DEV)
    steps = torch.arange(D, device=DEV, dtype=torch.float32).view(1, D, 1, 1)
    Gd = torch.randn(B, H, W, generator=gen).to(DEV)
    for depth, iv, hyp in ((vol, None, vol), (start, interval, start.view(B, 1, 1, 1) + interval.view(B, 1, 1, 1) * steps),
                           (start_map, interval, start_map.unsqueeze(1) + interval.view(B, 1, 1, 1) * steps)):
        s1 = sc.clone().requires_grad_(True)
        d1, conf = ops.regress_depth(s1, depth, iv, L.CONF_SUM4)
        assert not conf.requires_grad and conf.shape == (B, H, W)
        (d1 * Gd).sum().backward()
        s2 = sc.clone().requires_grad_(True)
        d2 = (torch.softmax(s2, 1) * hyp).sum(1)
        (d2 * Gd).sum().backward()
        assert rel_linf(d1.detach().cpu().numpy(), d2.detach().cpu().numpy()) < 1e-6
        assert rel_linf(s1.grad.cpu().numpy(), s2.grad.cpu().numpy()) < GRAD_TOL


def test_mvsnet_training_step_matches_the_reference(golden, monkeypatch):
    """One training-mode forward + L1 loss + backward of MVSNet-s (the reference's training configuration,
    models/trainer.py:96-206) with the reference's weights: depth map and parameter gradients across the whole model
    (FeatureNet <- K1 backward <- regulariser <- K3 backward)."""
    g = golden("mvsnet_train")
    # the golden was computed in fp32 on the CPU: keep cuDNN (FeatureNet / regulariser modules) off TF32 for the comparison
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    net = MVSNet("softmin")
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}, strict=True)
    net.num_depth = 8
    net = net.to(DEV).train()
    s = {k: v.to(DEV) for k, v in synth.make_sample(2, 3, 64, 96, seed=int(g["seed"])).items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    assert out["depth"].requires_grad and not out["photometric_confidence"].requires_grad
    err = np.abs(out["depth"].detach().cpu().numpy() - g["depth"]) / np.abs(g["depth"]).max()
    assert err.max() < 1e-3       # north-star tolerance; measured 3e-5
    loss = (out["depth"] - cu(g["target"])).abs().mean()
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * float(g["loss"])
    loss.backward()
    params = dict(net.named_parameters())
    for k in [k[5:] for k in g if k.startswith("grad.")]:
        want = g["grad." + k]
        if np.abs(want).max() < 1e-5:          # prob.bias: softmax is shift invariant, the gradient is rounding noise
            assert params[k].grad.abs().max().item() < 1e-4
            continue
        # measured 3e-5 ... 7e-5 (temp, one scalar summed over every voxel: 3e-3)
        assert rel_linf(params[k].grad.cpu().numpy(), want) < (1e-2 if k == "temp" else 1e-3), k
    # eval mode still runs the inference kernels, without a graph
    out = net.eval()(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    assert not out["depth"].requires_grad


def test_k3_backward_gradient_of_per_voxel_hypotheses():
    """CVP's refinement levels regress over hypotheses built from the previous level's depth: dL/dh_d = g p_d."""
    gen = torch.Generator().manual_seed(5)
    B, D, H, W = 2, 8, 11, 37
    sc = (2 * torch.randn(B, D, H, W, generator=gen)).to(DEV)
    base = (500 + 100 * torch.rand(B, H, W, generator=gen)).to(DEV)
    Gd = torch.randn(B, H, W, generator=gen).to(DEV)
    steps = torch.arange(-4, 4, device=DEV, dtype=torch.float32).view(1, D, 1, 1)
    grads = []
    for mine in (True, False):
        s1, b1 = sc.clone().requires_grad_(True), base.clone().requires_grad_(True)
        hyp = b1.unsqueeze(1) + 2.5 * steps
        d = ops.regress_depth(s1, hyp, None, L.CONF_NONE)[0] if mine else (torch.softmax(s1, 1) * hyp).sum(1)
        (d * Gd).sum().backward()
        grads.append((s1.grad, b1.grad))
    assert rel_linf(grads[0][0].cpu().numpy(), grads[1][0].cpu().numpy()) < GRAD_TOL
    assert rel_linf(grads[0][1].cpu().numpy(), grads[1][1].cpu().numpy()) < GRAD_TOL     # = Gd: the softmax sums to one
    with pytest.raises(L.Mvsb200Error):       # a [B,D] table of hypotheses shared by all pixels cannot take a gradient here
        dv = torch.linspace(1, 2, D, device=DEV).view(1, D).repeat(B, 1).requires_grad_(True)
        ops.regress_depth(sc.clone().requires_grad_(True), dv, None, L.CONF_NONE)[0].sum().backward()


@pytest.mark.parametrize("k2", ["cudnn", "lib"])
def test_cvp_training_step_matches_the_reference(golden, monkeypatch, k2):
    """One training-mode forward + L1 loss on both pyramid levels + backward of CVP-MVSNet with the reference's weights
    (net.py:96-229 with self.training: 48 initial hypotheses, fixed refinement intervals): K1 backward in its
    VARIANCE_MEAN / per-pixel-hypotheses mode, K3 backward including the gradient through the refinement hypotheses."""
    from wild_deep_mvs_b200.cvpmvsnet import Frontend
    g = golden("cvp_train")
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # the golden is fp32 on the CPU
    if k2 == "lib":      # the regulariser's 3x3x3 layers forward + backward on the library (K2 engines + wgrad kernel)
        monkeypatch.setenv("MVSB200_TRAIN_K2", "lib")
    else:
        monkeypatch.delenv("MVSB200_TRAIN_K2", raising=False)
    net = Frontend()
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}, strict=True)
    net.model.nscale = 2
    net = net.to(DEV).train()
    s = {k: v.to(DEV) for k, v in synth.make_sample(2, 3, 32, 48, seed=int(g["seed"])).items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=2)
    ests = out["depth_est_list"]
    assert [tuple(e.shape) for e in ests] == [(2, 32, 48), (2, 16, 24)] and all(e.requires_grad for e in ests)
    for i in range(2):
        assert rel_linf(ests[i].detach().cpu().numpy(), g["depth_est_%d" % i]) < 1e-3
    loss = sum((e - cu(g["target_%d" % i])).abs().mean() for i, e in enumerate(ests))
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * float(g["loss"])
    loss.backward()
    params = dict(net.named_parameters())
    for k in [k[5:] for k in g if k.startswith("grad.")]:
        err = rel_linf(params[k].grad.cpu().numpy(), g["grad." + k])
        print(k, "grad rel err %.2e" % err)
        assert err < 1e-2, k


@pytest.mark.parametrize("k2", ["cudnn", "lib"])
def test_vis_training_step_matches_the_reference(golden, monkeypatch, k2):
    """One training-mode forward of Vis-MVSNet (cascade [8,4,4]) with the reference's weights, a loss over every output
    the reference's loss touches (stage depths, pair depths weighted by their uncertainties, the uncertainties, the last
    probability map) and its backward: K1 backward in its group-correlation / start-map mode under the reference's own
    modules.  (Golden: tests/golden/make_golden_backward.py, with its documented shim for UncertNet's in-place add.)"""
    from wild_deep_mvs_b200.vismvsnet import Frontend
    g = golden("vis_train")
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)      # the golden is fp32 on the CPU
    if k2 == "lib":      # the 3x3x3 layers of Reg / RegPair / RegFuse forward + backward on the library
        monkeypatch.setenv("MVSB200_TRAIN_K2", "lib")
    else:
        monkeypatch.delenv("MVSB200_TRAIN_K2", raising=False)
    nums, scales = [8, 4, 4], [4, 2, 1]
    net = Frontend()
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}, strict=True)
    net.depth_nums, net.interval_scales = nums, scales
    net = net.to(DEV).train()
    s = {k: v.to(DEV) for k, v in synth.make_sample(2, 3, 64, 80, seed=int(g["seed"])).items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], depth_nums=nums, interval_scales=scales)
    loss = 0
    for k, e in enumerate(out["depth_est_list"]):
        assert e.requires_grad
        assert rel_linf(e.detach().cpu().numpy(), g["depth_est_%d" % k]) < 1e-3, k
        tgt = cu(g["target_%d" % k])
        loss = loss + (e - tgt).abs().mean()
        for v, (pd, heads) in enumerate(out["depth_pair_list"][k]):
            u = heads[0]
            assert rel_linf(pd.detach().cpu().numpy(), g["pair_depth_%d_%d" % (k, v)]) < 1e-3, (k, v)
            assert np.abs(u.detach().cpu().numpy() - g["pair_uncert_%d_%d" % (k, v)]).max() < 1e-3 * max(1.0, np.abs(g["pair_uncert_%d_%d" % (k, v)]).max())
            loss = loss + ((pd.squeeze(1) - tgt).abs() * (-u.squeeze(1)).exp() + u.squeeze(1)).mean() * 0.5
    loss = loss + out["photometric_confidence"][:, 2].mean()
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    loss.backward()
    params = dict(net.named_parameters())
    errs = {k[5:]: rel_linf(params[k[5:]].grad.cpu().numpy(), g[k]) for k in g if k.startswith("grad.")}
    print({k: "%.1e" % v for k, v in errs.items()})
    for k, err in errs.items():
        # Stages 2 and 3 (and what only they reach): measured 1e-6 ... 1e-4.  Everything stage 1 reaches (its own layers,
        # final_conv_1, the shared first layer) sits on a non-smooth point of the REFERENCE for this sample: scaling the
        # translation vectors by 1 + 2e-6 moves the reference's own CPU gradients by exactly the amounts measured here
        # (final_conv_1 2.7e-2, stage1.reg 2.6e-2, stage1.uncert_net 3.6e-2, init_conv 1.2e-2) while every forward
        # output stays within 2e-3 mm -- one discrete event in the 8x10-pixel stage-1 maps, reached by K1's 2e-5 px
        # difference in sample positions (fp64 relative pose vs the reference's fp32 matrix products).
        stage1 = k.startswith(("model.stage1.", "model.feat_ext.init_conv", "model.feat_ext.final_conv_1"))
        assert err < (6e-2 if stage1 else 1e-3 if k.startswith("model.stage") else 1e-2), (k, err)
    out = net.eval()(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], depth_nums=nums, interval_scales=scales)
    assert not out["depth"].requires_grad


def test_backward_errors_are_loud():
    ref = torch.zeros(1, 8, 8, 32, device=DEV)
    warp = torch.zeros(1, 1, 16, device=DEV)
    dv = torch.ones(1, 4, device=DEV)
    with pytest.raises(L.Mvsb200Error):
        ops.build_cost_volume_backward(torch.zeros(1, 4, 8, 8, 16, device=DEV), ref, [ref], warp, dv, 4, L.GEOM_MVS, L.AGG_VARIANCE)
    with pytest.raises(L.Mvsb200Error):
        ops.depth_regress_backward(torch.zeros(1, 3, 3, device=DEV), torch.zeros(1, 4, 8, 8, device=DEV), dv)


# ------------------------------------------------------------------------------------------------
# K2 in training: forward + input gradient on the K2 engines, weight gradient on mvsb200_conv3d_wgrad
# ------------------------------------------------------------------------------------------------
K2_TRAIN_CASES = [
    # cin, cout, stride, transposed, bias, input dims (D,H,W)
    (32, 8, 1, False, False, (6, 10, 36)),
    (8, 16, 2, False, False, (8, 12, 40)),
    (16, 16, 1, False, False, (4, 9, 33)),
    (64, 64, 1, False, False, (2, 5, 7)),
    (16, 8, 2, True, False, (3, 5, 34)),
    (64, 32, 2, True, False, (2, 3, 5)),
    (64, 32, 1, True, False, (4, 6, 9)),      # CVP conv5: ConvTranspose3d with stride 1
    (8, 1, 1, False, True, (6, 10, 36)),      # the `prob` head, with bias
]


@pytest.mark.parametrize("cin,cout,stride,transposed,bias,dims", K2_TRAIN_CASES)
def test_k2_training_conv_against_torch_autograd(cin, cout, stride, transposed, bias, dims, monkeypatch):
    """ops.conv3d_train (K2 engine forward, K2 engine input gradient with re-packed weights, wgrad kernel) against
    torch.autograd through F.conv3d / F.conv_transpose3d in fp32 -- the calls the reference's modules make."""
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    gen = torch.Generator().manual_seed(cin + cout + stride)
    B = 2
    x = torch.randn(B, cin, *dims, generator=gen).to(DEV)
    w = (torch.randn((cin, cout, 3, 3, 3) if transposed else (cout, cin, 3, 3, 3), generator=gen) / (cin * 27) ** 0.5).to(DEV)
    b = torch.randn(cout, generator=gen).to(DEV) if bias else None
    x1, w1 = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    b1 = b.clone().requires_grad_(True) if bias else None
    y1 = F.conv_transpose3d(x1, w1, b1, stride, 1, stride - 1) if transposed else F.conv3d(x1, w1, b1, stride, 1)
    G = torch.randn(y1.shape, generator=torch.Generator().manual_seed(1)).to(DEV)
    (y1 * G).sum().backward()
    x2, w2 = ops.to_ndhwc(x).requires_grad_(True), w.clone().requires_grad_(True)
    b2 = b.clone().requires_grad_(True) if bias else None
    y2 = ops.conv3d_train(x2, w2, b2, stride, transposed)
    assert rel_linf(ops.as_ncdhw(y2).detach().cpu().numpy(), y1.detach().cpu().numpy()) < 2e-5
    (ops.as_ncdhw(y2) * G).sum().backward()
    assert rel_linf(ops.as_ncdhw(x2.grad).cpu().numpy(), x1.grad.cpu().numpy()) < 2e-5
    assert rel_linf(w2.grad.cpu().numpy(), w1.grad.cpu().numpy()) < 2e-5
    if bias:
        assert rel_linf(b2.grad.cpu().numpy(), b1.grad.cpu().numpy()) < 2e-5


def test_mvsnet_training_step_with_the_regulariser_on_the_library(golden, monkeypatch):
    """The MVSNet-s training step of test_mvsnet_training_step_matches_the_reference with MVSB200_TRAIN_K2=lib: every
    3x3x3 layer of the regulariser forward and backward on the library (K2 engines + wgrad kernel)."""
    monkeypatch.setenv("MVSB200_TRAIN_K2", "lib")
    g = golden("mvsnet_train")
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    net = MVSNet("softmin")
    net.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in g.items() if k.startswith("sd.")}, strict=True)
    net.num_depth = 8
    net = net.to(DEV).train()
    s = {k: v.to(DEV) for k, v in synth.make_sample(2, 3, 64, 96, seed=int(g["seed"])).items()}
    out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    err = np.abs(out["depth"].detach().cpu().numpy() - g["depth"]) / np.abs(g["depth"]).max()
    assert err.max() < 1e-3
    loss = (out["depth"] - cu(g["target"])).abs().mean()
    loss.backward()
    params = dict(net.named_parameters())
    for k in [k[5:] for k in g if k.startswith("grad.")]:
        want = g["grad." + k]
        if np.abs(want).max() < 1e-5:
            continue
        e = rel_linf(params[k].grad.cpu().numpy(), want)
        print(k, "grad rel err %.2e" % e)
        assert e < (1e-2 if k == "temp" else 1e-3), k
