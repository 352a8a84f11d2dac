"""All BASELINE.json configs on one B200 (bench.py itself measures configs[1] only, as its contract says): full
`forward` of the three drop-in model families at the BASELINE sizes, timed with CUDA events, split into the 2-D feature
extractors (PyTorch / cuDNN, "next" row f1) and the hot path (features -> depth: K1-K4), which is what Mvox/s counts.

    python profiles/bench_configs.py [out.json]

Random-init weights with randomised BN statistics (no checkpoints offline), synthetic DTU-like cameras (synth.py).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wild_deep_mvs_b200 import ops, synth  # noqa: E402
from wild_deep_mvs_b200.cvpmvsnet import Frontend as CVP  # noqa: E402
from wild_deep_mvs_b200.mvsnet import MVSNet  # noqa: E402
from wild_deep_mvs_b200.vismvsnet import Frontend as Vis  # noqa: E402

DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# layer-wise compulsory fp32 traffic of the hot path of each config (bytes), the roofline denominator
ALGORITHMIC_BYTES = {"cfg1": 495.6e6, "cfg2": 1963.6e6, "cfg3 ": 2969.6e6, "cfg3'": 5823.0e6, "cfg4": 20160e6}


def hbm_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def add_roofline(r):
    """hot-path fraction of the measured HBM roofline: algorithmic bytes / hot-path time / measured copy bandwidth."""
    for key, nbytes in ALGORITHMIC_BYTES.items():
        if r["config"].startswith(key):
            r["algorithmic_MB"] = nbytes / 1e6
            r["hot_path_frac_of_hbm_roofline"] = round(nbytes / (r["hot_path_ms"] * 1e-3) / 1e9 / hbm_gbs(), 4)
    return r


def timed(fn, reps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def sample(views, h, w):
    return {k: v.to(DEV) for k, v in synth.make_sample(1, views, h, w, seed=0).items()}


def graphed_forward_ms(name, net, s, out, **kw):
    """The same forward replayed as ONE CUDA graph (net.graphed_forward); None if the model cannot be captured."""
    try:
        with torch.no_grad():
            g = net.graphed_forward(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], **kw)
            got = g()
            torch.cuda.synchronize()
            err = float((got["depth"] - out["depth"]).abs().max() / out["depth"].abs().max())
            assert err < 1e-4, "graphed forward differs from the eager one: %g" % err     # (cuDNN extractors: not bit-exact)
            return timed(lambda: g(), reps=20, warmup=3)
    except Exception as e:  # noqa: BLE001 -- reported in the line
        sys.stderr.write("%s: graphed forward unavailable: %r\n" % (name, e))
        torch.cuda.synchronize()
        return None


def run(name, net, s, vox, feat_fn, **kw):
    net = net.to(DEV).eval()
    call = lambda: net(s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"], **kw)
    out = call()
    assert torch.isfinite(out["depth"]).all(), name
    ms_fwd = timed(call)
    ms_graph = graphed_forward_ms(name, net, s, out, **kw)
    with torch.no_grad():
        ms_feat = timed(lambda: feat_fn(net, s))
    ms_hot = ms_fwd - ms_feat
    r = {"config": name, "voxels": vox, "forward_ms": round(ms_fwd, 3), "features_ms": round(ms_feat, 3),
         "hot_path_ms": round(ms_hot, 3), "hot_path_Mvox_per_s": round(vox / ms_hot / 1e3, 1),
         "depth_maps_per_s": round(1e3 / ms_fwd, 1), "depth_shape": list(out["depth"].shape)}
    if ms_graph is not None:
        r["forward_graphed_ms"] = round(ms_graph, 3)
        r["depth_maps_per_s_graphed"] = round(1e3 / ms_graph, 1)
    print(json.dumps(add_roofline(r)), flush=True)
    return r


def vis_graphed(name, net, s, vox, nums, scales):
    """Hot path of Vis-MVSNet replayed as one CUDA graph (features resident): what the launch-bound eager number hides."""
    net = net.to(DEV).eval()
    with torch.no_grad():
        interval = ((s["depth_max"] - s["depth_min"]) / 128)
        ref_cam = net.fill_cam_array(s["K"][:, 0], s["R"][:, 0], s["t"][:, 0], s["depth_min"][:, 0], interval[:, 0])
        src_cams = torch.stack([net.fill_cam_array(s["K"][:, i], s["R"][:, i], s["t"][:, i], s["depth_min"][:, i], interval[:, i])
                                for i in range(1, s["K"].shape[1])], 1)
        feats = [[ops.to_nhwc(f) for f in fv] for fv in ops.map_views(net.model.feat_ext, torch.unbind(s["imgs"], 1))]
        g = net.graphed(feats, ref_cam, src_cams, s["depth_min"][:, 0].contiguous(), interval[:, 0].contiguous(), nums, scales)
        eager = net.depth_from_features(feats, ref_cam, src_cams, s["depth_min"][:, 0].contiguous(), interval[:, 0].contiguous(), nums, scales)
        assert torch.equal(g()[0][2], eager[0][2])
        ms = timed(lambda: g(), reps=20, warmup=3)
    r = {"config": name + " -- hot path as one CUDA graph", "voxels": vox, "hot_path_ms": round(ms, 3),
         "hot_path_Mvox_per_s": round(vox / ms / 1e3, 1)}
    print(json.dumps(add_roofline(r)), flush=True)
    return r


def mvs_graphed(name, net, s, vox):
    """Hot path of MVSNet / MVSNet-s replayed as one CUDA graph (features resident), as bench.py measures cfg2."""
    from wild_deep_mvs_b200.mvsnet import build_proj_matrices
    net = net.to(DEV).eval()
    with torch.no_grad():
        feats = [ops.to_nhwc(f) for f in net.extract_features(list(torch.unbind(s["imgs"], 1)))]
        K = s["K"].clone()
        K[:, :, :2] /= 4
        projs = list(torch.unbind(build_proj_matrices(K, s["R"], s["t"]), 1))
        D = net.num_depth
        depth = s["depth_min"][:, :1] + (s["depth_max"][:, :1] - s["depth_min"][:, :1]) / (D - 1) * torch.arange(D, device=DEV).view(1, -1)
        g = net.graphed(feats, projs, depth)
        eager = net.depth_from_features(feats, projs, depth)
        assert torch.equal(g()[0], eager[0])
        ms = timed(lambda: g(), reps=50, warmup=5)
    r = {"config": name + " -- hot path as one CUDA graph", "voxels": vox, "hot_path_ms": round(ms, 3),
         "hot_path_Mvox_per_s": round(vox / ms / 1e3, 1)}
    print(json.dumps(add_roofline(r)), flush=True)
    return r


def main():
    torch.manual_seed(0)
    res = []
    # cfg1 / cfg2: MVSNet-s (softmin, D=48, 1+2 views) and MVSNet (variance, D=192, 1+4 views), 640x512
    feat_mvs = lambda net, s: net.extract_features(list(torch.unbind(s["imgs"], 1)))
    for name, agg, views, D in (("cfg1 MVSNet-s 1+2 views 640x512 D=48", "softmin", 3, 48),
                                ("cfg2 MVSNet 1+4 views 640x512 D=192", "variance", 5, 192)):
        net = MVSNet(agg)
        synth.randomize_norm_stats(net, seed=1)
        net.num_depth = D
        res.append(run(name, net, sample(views, 512, 640), D * 128 * 160, feat_mvs))
        res.append(mvs_graphed(name, net, sample(views, 512, 640), D * 128 * 160))
    # cfg3: Vis-MVSNet, 1+4 views, 640x512, default [32,16,8] and eval [64,32,16] hypotheses per stage
    feat_vis = lambda net, s: ops.map_views(net.model.feat_ext, torch.unbind(s["imgs"], 1))
    for name, nums, scales in (("cfg3 Vis-MVSNet 1+4 views 640x512 depth_nums [32,16,8]", [32, 16, 8], [4, 2, 1]),
                               ("cfg3' Vis-MVSNet eval setting depth_nums [64,32,16]", [64, 32, 16], [2, 1, 0.5])):
        net = Vis()
        synth.randomize_norm_stats(net, seed=2)
        net.depth_nums, net.interval_scales = nums, scales
        vox = nums[0] * 64 * 80 + nums[1] * 128 * 160 + nums[2] * 256 * 320
        res.append(run(name, net, sample(5, 512, 640), vox, feat_vis, depth_nums=nums, interval_scales=scales))
        res.append(vis_graphed(name, net, sample(5, 512, 640), vox, nums, scales))
    # cfg4: CVP-MVSNet, 1+4 views, 1600x1184, 5 pyramid levels (eval: 96 coarse hypotheses, 8 per refinement level)
    feat_cvp = lambda net, s: ops.map_views(lambda im: net.model.featurePyramid(im, 5), torch.unbind(s["imgs"], 1))
    net = CVP()
    synth.randomize_norm_stats(net, seed=3)
    vox = 96 * 74 * 100 + 8 * (148 * 200 + 296 * 400 + 592 * 800 + 1184 * 1600)
    res.append(run("cfg4 CVP-MVSNet 1+4 views 1600x1184 nscale=5", net, sample(5, 1184, 1600), vox, feat_cvp, nscale=5))
    out = sys.argv[1] if len(sys.argv) > 1 else None
    if out:
        json.dump({"engine": ops.DEFAULT_ENGINE, "gpu": torch.cuda.get_device_name(0), "results": res}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
