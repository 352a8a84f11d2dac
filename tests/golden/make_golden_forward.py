"""Golden outputs of the reference's eval-mode `forward(imgs, K, R, t, depth_min, depth_max, ...)` for the three model
families (SURVEY.md 8-a7, a13 and the CVP frontend: the glue around the kernels -- K/4, build_proj_matrices, fill_cam_array,
interval/128, the cascade / pyramid hand-over, the confidence maps).

Run in the BUILD container only (needs /root/reference):   python tests/golden/make_golden_forward.py

Weights are NOT stored: the drop-in model is constructed from a fixed seed (torch CPU init is deterministic), its BatchNorm
statistics randomised and its head convs gained (peaked softmax), and the UNMODIFIED reference loads that state_dict with
strict=True -- the test rebuilds the same drop-in from the same seed.  Stored: the reference's outputs (tests/golden/forward.npz).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle.ref_import import import_reference  # noqa: E402
from wild_deep_mvs_b200 import synth  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "forward.npz")


def build(kind):
    """The drop-in model of `kind` with the weights both the generator and the test use; returns (net, sample, kwargs)."""
    torch.manual_seed(0)
    if kind == "mvsnet":
        from wild_deep_mvs_b200.mvsnet import MVSNet
        net = MVSNet("variance")
        synth.randomize_norm_stats(net, seed=1)
        synth.scale_param(net.cost_regularization.prob.weight, 40.0)
        net.num_depth = 16
        return net.eval(), synth.make_sample(2, 3, 64, 96, seed=3), {}
    if kind == "vis":
        from wild_deep_mvs_b200.vismvsnet import Frontend
        net = Frontend()
        synth.randomize_norm_stats(net, seed=2)
        for st in (net.model.stage1, net.model.stage2, net.model.stage3):
            synth.scale_param(st.reg_fuse.final_conv.weight, 30.0)
            synth.scale_param(st.reg_pair.final_conv.weight, 30.0)
        kw = {"depth_nums": [8, 4, 4], "interval_scales": [4, 2, 1]}
        net.depth_nums, net.interval_scales = kw["depth_nums"], kw["interval_scales"]
        return net.eval(), synth.make_sample(1, 3, 64, 80, seed=4), kw
    from wild_deep_mvs_b200.cvpmvsnet import Frontend
    net = Frontend()
    synth.randomize_norm_stats(net, seed=3)
    synth.scale_param(net.model.cost_reg_refine.prob0.weight, 30.0)
    return net.eval(), synth.make_sample(1, 3, 64, 96, seed=5), {"nscale": 2}


def main():
    ref = import_reference()
    res = {}
    for kind, cls in (("mvsnet", lambda: ref.MVSNet("variance")), ("vis", ref.VisFrontend), ("cvp", ref.CVPFrontend)):
        mine, s, kw = build(kind)
        rnet = cls()
        rnet.load_state_dict(mine.state_dict(), strict=True)
        rnet.eval()
        if kind == "mvsnet":
            rnet.num_depth = mine.num_depth
        if kind == "vis":
            rnet.depth_nums, rnet.interval_scales = kw["depth_nums"], kw["interval_scales"]
        with torch.no_grad():
            out = rnet(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], **kw)
        res[kind + "_depth"] = out["depth"].numpy()
        res[kind + "_conf"] = out["photometric_confidence"].numpy()
        for i, d in enumerate(out["depth_est_list"]):
            res["%s_est%d" % (kind, i)] = d.numpy()
        for k, stage in enumerate(out["depth_pair_list"]):
            for v, (est, heads) in enumerate(stage):
                res["%s_pair%d_%d" % (kind, k, v)] = est.numpy()
                res["%s_uncert%d_%d" % (kind, k, v)] = heads[0].numpy()
        print(kind, "depth", float(out["depth"].min()), float(out["depth"].max()), "conf", float(out["photometric_confidence"].min()),
              float(out["photometric_confidence"].max()))
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
