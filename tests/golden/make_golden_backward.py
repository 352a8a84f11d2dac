"""Golden gradients for the backward kernels (row f2 of SURVEY.md 8), produced by autograd through the UNMODIFIED reference.

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_backward.py

tests/golden/backward.npz holds, on the inputs already stored in mvsnet_*.npz / vis.npz:
  * `MVSNet.build_cost_volume` in training mode (models/MVSNet/model.py:109-176, variance and soft-min): the gradient of
    sum(volume * G) for a seeded G with respect to the three feature maps (and `temp`);
  * Vis-MVSNet `SingleStage.build_cost_volume` + `groupwise_correlation` (model_cas.py:176-186, nn_utils.py:473-490):
    the same for the 8-group correlation volumes of both source views;
  * softmax + `depth_regression` (model.py:207-209): gradient of sum(depth * Gd) with respect to the score volume;
  * one training-mode `forward` + L1 loss of MVSNet-s (the reference's training configuration): the depth map and the
    gradients of a few parameters across the whole model (`mvsnet_train.npz`, with the weights);
  * the same for CVP-MVSNet with two pyramid levels (`cvp_train.npz`): training-mode hypotheses (net.py:126,176-182),
    loss on both levels, so the gradient also runs through the refinement level's hypotheses into the coarse level.

  * the same for Vis-MVSNet with a [8,4,4] cascade (`vis_train.npz`).

    python tests/golden/make_golden_backward.py [cvp|vis]      # only cvp_train.npz / vis_train.npz
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402
from wild_deep_mvs_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
T = torch.from_numpy


def cvp_train(ref):
    torch.manual_seed(0)
    net = ref.CVPFrontend()
    synth.randomize_norm_stats(net, seed=3)
    synth.scale_param(net.model.cost_reg_refine.prob0.weight, 30.0)
    net.model.nscale = 2
    sd = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    net.train()
    s = synth.make_sample(2, 3, 32, 48, seed=6)
    res = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=2)
    gen = torch.Generator().manual_seed(9)
    ests = res["depth_est_list"]
    targets = [500 + 300 * torch.rand(e.shape, generator=gen) for e in ests]
    loss = sum((e - t).abs().mean() for e, t in zip(ests, targets))
    loss.backward()
    tr = {"sd." + k: v for k, v in sd.items()}
    tr.update({"depth_est_0": ests[0].detach().numpy(), "depth_est_1": ests[1].detach().numpy(), "target_0": targets[0].numpy(),
               "target_1": targets[1].numpy(), "conf": res["photometric_confidence"].numpy(), "loss": np.float32(loss.item()),
               "seed": np.int32(6)})
    params = dict(net.named_parameters())
    for k in ("model.featurePyramid.conv0aa.0.weight", "model.featurePyramid.conv0bh.0.bias", "model.cost_reg_refine.conv0.conv.weight",
              "model.cost_reg_refine.conv4a.bn.weight", "model.cost_reg_refine.conv6.0.weight", "model.cost_reg_refine.prob0.weight"):
        tr["grad." + k] = params[k].grad.numpy()
        print(k, "grad abs-max", float(params[k].grad.abs().max()))
    np.savez_compressed(os.path.join(OUT, "cvp_train.npz"), **tr)
    print("cvp_train.npz", os.path.getsize(os.path.join(OUT, "cvp_train.npz")) // 1024, "KiB")


def vis_train(ref):
    """Vis-MVSNet, training mode, cascade [8,4,4]: a loss over every output the reference's loss touches (stage depths,
    pair depths weighted by their uncertainties, the uncertainties themselves, the probability map of the last stage)."""
    # Shim 3 (training only): UncertNet.forward adds its input IN PLACE to the output of a ReLU (`out += x`,
    # models/VisMVSNet/model_cas.py:96), which torch >= 1.5 refuses to differentiate (ReLU's backward needs that
    # output; the pinned torch 1.4 let it pass).  The out-of-place form computes the same numbers.
    import models.VisMVSNet.model_cas as model_cas

    def uncert_forward(self, x):
        out = self.conv2(self.conv1(x))
        out = out + x
        return [conv(out) for conv in self.head_convs]

    model_cas.UncertNet.forward = uncert_forward
    torch.manual_seed(0)
    net = ref.VisFrontend()
    synth.randomize_norm_stats(net, seed=2)
    for st in (net.model.stage1, net.model.stage2, net.model.stage3):
        synth.scale_param(st.reg_fuse.final_conv.weight, 30.0)
        synth.scale_param(st.reg_pair.final_conv.weight, 30.0)
    nums, scales = [8, 4, 4], [4, 2, 1]
    net.depth_nums, net.interval_scales = nums, scales
    sd = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    net.train()
    s = synth.make_sample(2, 3, 64, 80, seed=8)
    res = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], depth_nums=nums, interval_scales=scales)
    gen = torch.Generator().manual_seed(9)
    loss = 0
    tr = {}
    for k, e in enumerate(res["depth_est_list"]):
        tgt = 500 + 300 * torch.rand(e.shape, generator=gen)
        tr["target_%d" % k] = tgt.numpy()
        tr["depth_est_%d" % k] = e.detach().numpy()
        loss = loss + (e - tgt).abs().mean()
        for v, (pd, heads) in enumerate(res["depth_pair_list"][k]):
            u = heads[0]
            loss = loss + ((pd.squeeze(1) - tgt).abs() * (-u.squeeze(1)).exp() + u.squeeze(1)).mean() * 0.5
            tr["pair_depth_%d_%d" % (k, v)] = pd.detach().numpy()
            tr["pair_uncert_%d_%d" % (k, v)] = u.detach().numpy()
    loss = loss + res["photometric_confidence"][:, 2].mean()
    loss.backward()
    tr.update({"sd." + k: v for k, v in sd.items()})
    tr.update({"conf": res["photometric_confidence"].detach().numpy(), "loss": np.float32(loss.item()), "seed": np.int32(8)})
    params = dict(net.named_parameters())
    for k in ("model.feat_ext.init_conv.0.weight", "model.feat_ext.final_conv_1.weight", "model.feat_ext.final_conv_3.weight",
              "model.stage1.reg.unet.enc_blocks.reg14_0.0.conv1.weight", "model.stage1.uncert_net.conv1.0.weight",
              "model.stage2.reg_fuse.final_conv.weight", "model.stage3.reg.unet.dec_blocks.reg116_2.1.weight",
              "model.stage3.reg_pair.final_conv.weight", "model.stage3.uncert_net.head_convs.0.weight"):
        tr["grad." + k] = params[k].grad.numpy()
        print(k, "grad abs-max", float(params[k].grad.abs().max()))
    np.savez_compressed(os.path.join(OUT, "vis_train.npz"), **tr)
    print("vis_train.npz", os.path.getsize(os.path.join(OUT, "vis_train.npz")) // 1024, "KiB")


def main():
    ref = import_reference()
    torch.set_num_threads(8)
    if sys.argv[1:] == ["cvp"]:
        return cvp_train(ref)
    if sys.argv[1:] == ["vis"]:
        return vis_train(ref)
    out = {}

    # ---- MVSNet cost volumes ---------------------------------------------------------------------------------------
    for agg in ("variance", "softmin"):
        g = dict(np.load(os.path.join(OUT, "mvsnet_%s.npz" % agg)))
        net = ref.MVSNet(agg).train()
        net.num_depth = g["depth_values"].shape[1]
        if agg == "softmin":
            with torch.no_grad():
                net.temp.copy_(T(g["temp"]))
        feats = [T(g["feat%d" % i]).clone().requires_grad_(True) for i in range(3)]
        proj = T(g["proj"])
        vol = net.build_cost_volume(feats[0], feats[1:], proj[:, 0], [proj[:, 1], proj[:, 2]], T(g["depth_values"]))
        G = torch.randn(vol.shape, generator=torch.Generator().manual_seed(11))
        (vol * G).sum().backward()
        out["%s_G" % agg] = G.numpy()
        for i in range(3):
            out["%s_gfeat%d" % (agg, i)] = feats[i].grad.numpy()
        if agg == "softmin":
            out["softmin_gtemp"] = net.temp.grad.numpy()
        print(agg, "train-mode volume vs eval golden:", float((vol.detach() - T(g["cost_volume"])).abs().max()))

    # ---- Vis-MVSNet group correlation ------------------------------------------------------------------------------
    g = dict(np.load(os.path.join(OUT, "vis.npz")))
    net = ref.VisFrontend().train()
    st1 = net.model.stage1
    feats = [T(g["feat_v%d_s1" % v]).clone().requires_grad_(True) for v in range(3)]
    ref_cam, src_cams = T(g["ref_cam"]), [T(g["src_cam1"]), T(g["src_cam2"])]
    D = 8
    interval = (T(g["depth_max"]) - T(g["depth_min"])) / 128
    ds = ref_cam[:, 1:2, 3:4, 0:1]
    di = interval[:, 0].view(1, 1, 1, 1) * 4
    refvol = feats[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    loss = 0
    for v in range(2):
        warped = st1.build_cost_volume(feats[0], ref_cam, feats[1 + v], src_cams[v], D, ds, di, 8, 1)
        gc = ref.vis_nn.groupwise_correlation(refvol, warped, 8, 1)
        G = torch.randn(gc.shape, generator=torch.Generator().manual_seed(20 + v))
        out["vis_G%d" % v] = G.numpy()
        loss = loss + (gc * G).sum()
        if v == 0:
            print("vis groupcorr vs golden:", float((gc.detach() - T(g["s1_groupcorr_pair0"])).abs().max()))
    loss.backward()
    for v in range(3):
        out["vis_gfeat%d" % v] = feats[v].grad.numpy()

    # ---- softmax + depth regression --------------------------------------------------------------------------------
    gen = torch.Generator().manual_seed(3)
    score = (4 * torch.randn(2, 12, 9, 21, generator=gen)).requires_grad_(True)
    dvals = 425 + 480 * torch.sort(torch.rand(2, 12, generator=gen), 1)[0]
    depth = ref.mvs_module.depth_regression(F.softmax(score, dim=1), dvals)
    Gd = torch.randn(depth.shape, generator=gen)
    (depth * Gd).sum().backward()
    out.update({"reg_score": score.detach().numpy(), "reg_dvals": dvals.numpy(), "reg_Gd": Gd.numpy(),
                "reg_depth": depth.detach().numpy(), "reg_gscore": score.grad.numpy()})
    np.savez_compressed(os.path.join(OUT, "backward.npz"), **out)

    # ---- one training step of MVSNet-s -----------------------------------------------------------------------------
    torch.manual_seed(0)
    net = ref.MVSNet("softmin")
    synth.randomize_norm_stats(net, seed=1)
    synth.scale_param(net.cost_regularization.prob.weight, 40.0)
    with torch.no_grad():
        net.temp.fill_(0.37)
    net.num_depth = 8
    sd = {k: v.detach().clone().numpy() for k, v in net.state_dict().items()}
    net.train()
    s = synth.make_sample(2, 3, 64, 96, seed=4)
    res = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    target = 500 + 300 * torch.rand(res["depth"].shape, generator=torch.Generator().manual_seed(9))
    loss = (res["depth"] - target).abs().mean()
    loss.backward()
    tr = {"sd." + k: v for k, v in sd.items()}
    tr.update({"depth": res["depth"].detach().numpy(), "conf": res["photometric_confidence"].numpy(), "target": target.numpy(),
               "loss": np.float32(loss.item()), "seed": np.int32(4)})
    params = dict(net.named_parameters())
    for k in ("temp", "feature.conv0.conv.weight", "feature.conv6.bn.weight", "feature.feature.bias",
              "cost_regularization.conv0.conv.weight", "cost_regularization.conv6.bn.bias",
              "cost_regularization.conv9.0.weight", "cost_regularization.prob.weight", "cost_regularization.prob.bias"):
        tr["grad." + k] = params[k].grad.numpy()
        print(k, "grad abs-max", float(params[k].grad.abs().max()))
    np.savez_compressed(os.path.join(OUT, "mvsnet_train.npz"), **tr)
    for f in ("backward.npz", "mvsnet_train.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
    cvp_train(ref)
    vis_train(ref)


if __name__ == "__main__":
    main()
