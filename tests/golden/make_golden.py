"""Generate the golden tensors that pin the oracle and the CUDA path.

Run in the BUILD container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the unmodified reference (oracle/ref_import.py), runs it on CPU on the seeded
synthetic inputs of wild_deep_mvs_b200/synth.py with non-trivial BN statistics and a
gained head conv (so the softmax over depth is peaked), and stores inputs,
weights and the tensor at every seam of SURVEY.md section 8(a) as small .npz files.
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


def npd(d):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def sd_np(module, prefix):
    return {prefix + k: v.detach().cpu().numpy() for k, v in module.state_dict().items()}


def gen_mvsnet(ref, aggregation):
    torch.manual_seed(0)
    net = ref.MVSNet(aggregation).eval()
    synth.randomize_norm_stats(net, seed=1)
    synth.scale_param(net.cost_regularization.prob.weight, 40.0)
    D = 8
    net.num_depth = D
    if aggregation == "softmin":
        with torch.no_grad():
            net.temp.fill_(0.37)
    s = synth.make_sample(1, 3, 64, 96, seed=0)
    out = {}
    with torch.no_grad():
        feats = net.extract_features(torch.unbind(s["imgs"], 1))
        # spread the features out a bit (random-init FeatureNet output is nearly constant)
        g = torch.Generator().manual_seed(5)
        feats = [f + 0.5 * torch.randn(f.shape, generator=g) for f in feats]
        from utils.utils_3D import build_proj_matrices
        sk = s["K"].clone()
        sk[:, :, :2] /= 4
        proj = build_proj_matrices(sk, s["R"], s["t"])
        dvals = s["depth_min"][:, 0:1] + (s["depth_max"][:, 0:1] - s["depth_min"][:, 0:1]) / (D - 1) * torch.arange(D).view(1, -1)
        warped0 = ref.mvs_module.homo_warping(feats[1], proj[:, 1], proj[:, 0], dvals, feats[0].shape[-2:])
        cost = net.build_cost_volume(feats[0], feats[1:], proj[:, 0], [proj[:, 1], proj[:, 2]], dvals)
        cost_in = cost.clone()
        reg = net.cost_regularization(cost).squeeze(1)
        prob = F.softmax(reg, dim=1)
        depth = ref.mvs_module.depth_regression(prob, dvals)
        p4 = 4 * F.avg_pool3d(F.pad(prob.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
        idx = ref.mvs_module.depth_regression(prob, torch.arange(D, dtype=torch.float)).long()
        conf = torch.gather(p4, 1, idx.unsqueeze(1)).squeeze(1)
        rel = proj[:, 1:] @ torch.inverse(proj[:, 0:1])
    out.update(npd({"feat0": feats[0], "feat1": feats[1], "feat2": feats[2], "proj": proj, "rel_proj": rel,
                    "depth_values": dvals, "warped_view1": warped0, "cost_volume": cost_in, "cost_reg": reg,
                    "prob": prob, "depth": depth, "conf": conf,
                    "temp": net.temp if aggregation == "softmin" else torch.zeros(1)}))
    out.update(sd_np(net.cost_regularization, "cost_regularization."))
    np.savez(os.path.join(OUT, "mvsnet_%s.npz" % aggregation), **out)
    print("mvsnet", aggregation, "depth range", depth.min().item(), depth.max().item(),
          "conf", conf.min().item(), conf.max().item())


def gen_warp_ragged(ref):
    """homo_warping with a source map of another size, per-pixel depths, and points behind the camera."""
    g = torch.Generator().manual_seed(3)
    src = torch.randn(1, 8, 13, 19, generator=g)
    H, W, D = 10, 14, 5
    K, R, t, _, _ = synth.make_cameras(1, 2, 40, 56)
    from utils.utils_3D import build_proj_matrices
    sk = K.clone()
    sk[:, :, :2] /= 4
    proj = build_proj_matrices(sk, R, t)
    dpp = 425 + 480 * torch.rand(1, D, H, W, generator=g)
    dpp[0, 0, :2] = -50.0  # behind the camera -> (-10,-10) branch, MVSNet/module.py:147-150
    with torch.no_grad():
        w = ref.mvs_module.homo_warping(src, proj[:, 1], proj[:, 0], dpp, (H, W))
        rel = proj[:, 1] @ torch.inverse(proj[:, 0])
    np.savez(os.path.join(OUT, "warp_ragged.npz"), **npd({"src": src, "proj": proj, "rel_proj": rel,
                                                          "depth": dpp, "warped": w}))
    print("warp_ragged nonzero frac", (w != 0).float().mean().item())


def gen_vis(ref):
    torch.manual_seed(0)
    net = ref.VisFrontend().eval()
    synth.randomize_norm_stats(net, seed=2)
    for st in (net.model.stage1, net.model.stage2, net.model.stage3):
        synth.scale_param(st.reg_fuse.final_conv.weight, 30.0)
        synth.scale_param(st.reg_pair.final_conv.weight, 30.0)
    depth_nums, scales = [8, 4, 4], [4, 2, 1]
    net.depth_nums, net.interval_scales = depth_nums, scales
    s = synth.make_sample(1, 3, 64, 80, seed=0)
    captured = {}

    def hook(name):
        def fn(mod, args, output):
            captured.setdefault(name, []).append(output.detach().clone())
        return fn

    hs = [net.model.stage1.reg.register_forward_hook(hook("s1_interm")),
          net.model.stage1.reg_pair.register_forward_hook(hook("s1_pair_score")),
          net.model.stage1.reg_fuse.register_forward_hook(hook("s1_fuse_score")),
          net.model.stage1.reg_fuse.unet.register_forward_hook(hook("s1_fuse_unet")),
          net.model.feat_ext.register_forward_hook(lambda m, a, o: captured.setdefault("feats", []).append([t.detach().clone() for t in o]))]
    with torch.no_grad():
        out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], depth_nums=depth_nums,
                  interval_scales=scales)
    for h in hs:
        h.remove()
    feats = captured["feats"]  # [ref, src1, src2] each (f1, f2, f3)
    # seam tensors of stage 1, pair 0, recomputed by calling the reference pieces directly
    st1 = net.model.stage1
    interval = (s["depth_max"] - s["depth_min"]) / 128
    ref_cam = net.fill_cam_array(s["K"][:, 0], s["R"][:, 0], s["t"][:, 0], s["depth_min"][:, 0], interval[:, 0])
    src_cams = [net.fill_cam_array(s["K"][:, i], s["R"][:, i], s["t"][:, i], s["depth_min"][:, i], interval[:, i]) for i in (1, 2)]
    with torch.no_grad():
        ds = ref_cam[:, 1:2, 3:4, 0:1]
        di = interval[:, 0].view(1, 1, 1, 1) * scales[0]
        warped = st1.build_cost_volume(feats[0][0], ref_cam, feats[1][0], src_cams[0], depth_nums[0], ds, di, 8, 1)
        refvol = feats[0][0].unsqueeze(2).repeat(1, 1, depth_nums[0], 1, 1)
        gc = ref.vis_nn.groupwise_correlation(refvol, warped, 8, 1)
    res = {"imgs_unused": np.zeros(1, np.float32)}
    for v in range(3):
        for k in range(3):
            res["feat_v%d_s%d" % (v, k + 1)] = feats[v][k]
    res.update({"ref_cam": ref_cam, "src_cam1": src_cams[0], "src_cam2": src_cams[1],
                "s1_warped_pair0": warped, "s1_groupcorr_pair0": gc,
                "s1_interm_pair0": captured["s1_interm"][0], "s1_interm_pair1": captured["s1_interm"][1],
                "s1_pair_score0": captured["s1_pair_score"][0], "s1_fuse_score": captured["s1_fuse_score"][0],
                "s1_fuse_unet": captured["s1_fuse_unet"][0],
                "depth": out["depth"], "conf": out["photometric_confidence"],
                "depth_est_0": out["depth_est_list"][0], "depth_est_1": out["depth_est_list"][1],
                "depth_est_2": out["depth_est_list"][2],
                "depth_min": s["depth_min"], "depth_max": s["depth_max"]})
    for k in range(3):
        for v in range(2):
            res["pair_depth_st%d_v%d" % (k, v)] = out["depth_pair_list"][k][v][0]
            res["pair_uncert_st%d_v%d" % (k, v)] = out["depth_pair_list"][k][v][1][0]
    res = npd(res)
    for name, st in (("stage1", net.model.stage1), ("stage2", net.model.stage2), ("stage3", net.model.stage3)):
        res.update(sd_np(st, "model.%s." % name))
    np.savez(os.path.join(OUT, "vis.npz"), **res)
    print("vis depth range", out["depth"].min().item(), out["depth"].max().item(),
          "conf3", out["photometric_confidence"][:, 2].min().item(), out["photometric_confidence"][:, 2].max().item())


def gen_cvp(ref):
    torch.manual_seed(0)
    net = ref.CVPFrontend().eval()
    synth.randomize_norm_stats(net, seed=3)
    synth.scale_param(net.model.cost_reg_refine.prob0.weight, 30.0)
    net.model.nscale = 2
    s = synth.make_sample(1, 3, 32, 48, seed=0)
    captured = {}
    hs = [net.model.featurePyramid.register_forward_hook(lambda m, a, o: captured.setdefault("fp", []).append([t.detach().clone() for t in o])),
          net.model.cost_reg_refine.register_forward_hook(lambda m, a, o: (captured.setdefault("reg_in", []).append(a[0].detach().clone()),
                                                                            captured.setdefault("reg_out", []).append(o.detach().clone())) and None)]
    # capture the per-level hypotheses
    orig = ref.cvp_modules.calDepthHypo
    import models.CVP_MVSNet.models.net as cvp_net

    def spy(*a, **k):
        r = orig(*a, **k)
        captured.setdefault("hypos", []).append(r.detach().clone())
        return r

    cvp_net.calDepthHypo = spy
    with torch.no_grad():
        out = net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], nscale=2)
    cvp_net.calDepthHypo = orig
    for h in hs:
        h.remove()
    res = {"K": s["K"], "R": s["R"], "t": s["t"], "depth_min": s["depth_min"], "depth_max": s["depth_max"],
           "depth": out["depth"], "depth_est_0": out["depth_est_list"][0], "depth_est_1": out["depth_est_list"][1],
           "conf": out["photometric_confidence"], "hypos_l0": captured["hypos"][0],
           "reg_in_coarse": captured["reg_in"][0], "reg_out_coarse": captured["reg_out"][0],
           "reg_in_l0": captured["reg_in"][1], "reg_out_l0": captured["reg_out"][1]}
    for v in range(3):
        for lvl in range(2):
            res["fp_v%d_l%d" % (v, lvl)] = captured["fp"][v][lvl]
    res = npd(res)
    res.update(sd_np(net.model.cost_reg_refine, "model.cost_reg_refine."))
    np.savez(os.path.join(OUT, "cvp.npz"), **res)
    print("cvp depth range", out["depth"].min().item(), out["depth"].max().item())


if __name__ == "__main__":
    ref = import_reference()
    torch.set_num_threads(8)
    gen_mvsnet(ref, "variance")
    gen_mvsnet(ref, "softmin")
    gen_warp_ragged(ref)
    gen_vis(ref)
    gen_cvp(ref)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")
