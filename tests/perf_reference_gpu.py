"""The UNMODIFIED reference through stock PyTorch / cuDNN on the B200 -- the existing Blackwell path to beat (SURVEY.md 2.2,
8-d) -- timed for every BASELINE config next to the drop-in (measurement script, GPU box; lives under tests/ because only
tests and the bench's reference legs may touch oracle/).

    python tests/perf_reference_gpu.py [out.json]

Per config: the reference's full `forward(imgs, K, R, t, depth_min, depth_max, ...)` in fp32 (TF32 off) and with TF32 allowed
(cudnn.benchmark on, median of 5 after 2 warm-up calls, CUDA events), and the drop-in's eager forward on the same inputs and
weights.  The reference is oracle/_ref/ref_hotpath.zip (byte-for-byte files of /root/reference, oracle/make_ref.py).
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import, ref_run  # noqa: E402
from wild_deep_mvs_b200 import synth  # noqa: E402
from wild_deep_mvs_b200.cvpmvsnet import Frontend as CVP  # noqa: E402
from wild_deep_mvs_b200.mvsnet import MVSNet  # noqa: E402
from wild_deep_mvs_b200.vismvsnet import Frontend as Vis  # noqa: E402

DEV = "cuda:0"


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


def tf32(on):
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = on


def compare(name, net, rnet, s, vox, **kw):
    a = (s["imgs"], s["K"], s["R"], s["t"], s["depth_min"], s["depth_max"])
    r = {"config": name, "voxels": vox}
    with torch.no_grad():
        tf32(False)
        want = rnet(*a, **kw)["depth"].clone()
        r["reference_fp32_ms"] = round(timed(lambda: rnet(*a, **kw)), 3)
        tf32(True)
        got_tf32 = rnet(*a, **kw)["depth"].clone()
        r["reference_tf32_ms"] = round(timed(lambda: rnet(*a, **kw)), 3)
        r["reference_tf32_vs_fp32_depth_rel_linf"] = float((got_tf32 - want).abs().max() / want.abs().max())
        del got_tf32
        torch.cuda.empty_cache()
        tf32(False)     # the drop-in's cuDNN feature extractors in fp32 too: a like-for-like accuracy statement
        got = net(*a, **kw)["depth"]
        r["ours_vs_reference_fp32_depth_rel_linf"] = float((got - want).abs().max() / want.abs().max())
        r["ours_fp32_features_ms"] = round(timed(lambda: net(*a, **kw)), 3)
        tf32(True)      # torch's default (what bench_configs.py measures)
        r["ours_ms"] = round(timed(lambda: net(*a, **kw)), 3)
    r["speedup_vs_reference_fp32"] = round(r["reference_fp32_ms"] / r["ours_ms"], 1)
    r["speedup_vs_reference_tf32"] = round(r["reference_tf32_ms"] / r["ours_ms"], 1)
    print(json.dumps(r), flush=True)
    return r


def main():
    ref = ref_import.import_reference(cuda_shim=False)
    torch.backends.cudnn.benchmark = True
    sample = lambda v, h, w: {k: t.to(DEV) for k, t in synth.make_sample(1, v, h, w, seed=0).items()}
    res = []
    for name, agg, views, D in (("cfg1 MVSNet-s 1+2 views 640x512 D=48", "softmin", 3, 48),
                                ("cfg2 MVSNet 1+4 views 640x512 D=192", "variance", 5, 192)):
        torch.manual_seed(0)
        net = MVSNet(agg)
        synth.randomize_norm_stats(net, seed=1)
        synth.scale_param(net.cost_regularization.prob.weight, 40.0)
        net.num_depth = D
        net = net.to(DEV).eval()
        rnet = ref_run.reference_mvsnet(ref, net.state_dict(), agg, D, DEV)
        res.append(compare(name, net, rnet, sample(views, 512, 640), D * 128 * 160))
        del net, rnet
    for name, nums, scales in (("cfg3 Vis-MVSNet 1+4 views 640x512 depth_nums [32,16,8]", [32, 16, 8], [4, 2, 1]),
                               ("cfg3' Vis-MVSNet eval setting depth_nums [64,32,16]", [64, 32, 16], [2, 1, 0.5])):
        torch.manual_seed(0)
        net = Vis()
        synth.randomize_norm_stats(net, seed=2)
        for st in (net.model.stage1, net.model.stage2, net.model.stage3):
            synth.scale_param(st.reg_fuse.final_conv.weight, 30.0)
            synth.scale_param(st.reg_pair.final_conv.weight, 30.0)
        net = net.to(DEV).eval()
        rnet = ref.VisFrontend()
        rnet.load_state_dict(net.state_dict(), strict=True)
        rnet = rnet.to(DEV).eval()
        for m in (net, rnet):
            m.depth_nums, m.interval_scales = nums, scales
        vox = nums[0] * 64 * 80 + nums[1] * 128 * 160 + nums[2] * 256 * 320
        res.append(compare(name, net, rnet, sample(5, 512, 640), vox, depth_nums=nums, interval_scales=scales))
        del net, rnet
    torch.manual_seed(0)
    net = CVP()
    synth.randomize_norm_stats(net, seed=3)
    synth.scale_param(net.model.cost_reg_refine.prob0.weight, 30.0)
    net = net.to(DEV).eval()
    rnet = ref.CVPFrontend()
    rnet.load_state_dict(net.state_dict(), strict=True)
    rnet = rnet.to(DEV).eval()
    vox = 96 * 74 * 100 + 8 * (148 * 200 + 296 * 400 + 592 * 800 + 1184 * 1600)
    res.append(compare("cfg4 CVP-MVSNet 1+4 views 1600x1184 nscale=5", net, rnet, sample(5, 1184, 1600), vox, nscale=5))
    if len(sys.argv) > 1:
        json.dump({"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "results": res}, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
