"""A/B timing of K1 builds on the GPU box: every `--so` library is loaded in a fresh process and times K1 at

  * cfg2      MVS geometry, variance, C = 32, 4 source views, D = 192, 128 x 160 (scalar hypotheses)
  * cvp-like  MVS geometry, variance-mean, C = 16, 4 source views, D = 8, 592 x 800, per-pixel hypotheses
  * vis-like  VIS geometry, 8-group correlation, C = 32, 4 source views, D = 32, 128 x 160, start map + interval

    python profiles/k1_ab.py wild_deep_mvs_b200/libmvsb200.so wild_deep_mvs_b200/build/libmvsb200_mb5.so

(L2 flushed before every timed launch; median and min of 20.)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child(so):
    import torch
    sys.path.insert(0, ROOT)
    from wild_deep_mvs_b200 import _lib as L
    L.SO_PATH = os.path.abspath(so)
    from wild_deep_mvs_b200 import ops, synth
    from wild_deep_mvs_b200.mvsnet import build_proj_matrices
    dev = "cuda:0"
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(20):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    def mvs_case(C, h, w, D, agg, per_pixel, views=5):
        V = views
        feats = [ops.to_nhwc(f.to(dev)) for f in synth.make_features(1, V, C, h, w, seed=0)]
        K, R, t, dmin, dmax = synth.make_cameras(1, V, 4 * h, 4 * w)
        K = K.clone()
        K[:, :, :2] /= 4
        projs = build_proj_matrices(K, R, t).to(dev)
        warp = ops.mvs_relative_proj(projs[:, 0], projs[:, 1:])
        depth = (dmin[:, :1] + (dmax[:, :1] - dmin[:, :1]) / (D - 1) * torch.arange(D).view(1, -1)).to(dev)
        if per_pixel:
            depth = (depth.view(1, D, 1, 1) + torch.rand(1, 1, h, w, device=dev)).expand(1, D, h, w).contiguous()
        out = torch.empty(1, D, h, w, C, device=dev)
        amax = torch.zeros(1, device=dev)
        temp = torch.ones(1, device=dev) if agg == L.AGG_SOFTMIN else None
        return lambda: ops.build_cost_volume(feats[0], feats[1:], warp, depth, D, L.GEOM_MVS, agg, temp=temp, out=out, amax=amax)

    def vis_case(h, w, D, B=1, start_map=True, scale=1.0):
        V, C = 5, 32
        feats = [ops.to_nhwc(f.to(dev)).expand(B, -1, -1, -1).contiguous() for f in synth.make_features(1, V, C, h, w, seed=0)]
        K, R, t, dmin, dmax = synth.make_cameras(1, V, 4 * h, 4 * w)
        cams = torch.zeros(1, V, 2, 4, 4)
        cams[:, :, 0, :3, :3] = R
        cams[:, :, 0, :3, 3:] = t
        cams[:, :, 0, 3, 3] = 1
        cams[:, :, 1, :3, :3] = K
        cams = cams.to(dev).expand(B, -1, -1, -1, -1).contiguous()
        warp = ops.vis_homography_params(cams[:, 0], cams[:, 1:], 0.25)
        if start_map:
            start = (dmin[:, :1].view(1, 1, 1) + 40 * torch.rand(B, h, w)).to(dev)
        else:
            start = dmin[:, :1].view(1).expand(B).contiguous().to(dev)
        interval = torch.full((B,), 2.5 * 192 / 128 * scale, device=dev)
        out = torch.empty(4, B, D, h, w, 8, device=dev)
        amax = torch.zeros(1, device=dev)
        return lambda: ops.build_cost_volume(feats[0], feats[1:], warp, start, D, L.GEOM_VIS, L.AGG_GROUPCORR, interval=interval,
                                             out=out, amax=amax)

    cases = [("cfg2", lambda: mvs_case(32, 128, 160, 192, L.AGG_VARIANCE, False)),
             ("cvp-like", lambda: mvs_case(16, 592, 800, 8, L.AGG_VARIANCE_MEAN, True)),
             ("vis-like", lambda: vis_case(128, 160, 32)),
             ("vis-s1x8", lambda: vis_case(64, 80, 64, B=8, start_map=False, scale=2.0)),     # cfg5 stage 1: scalar start, 8 views per launch
             ("vis-s2x8", lambda: vis_case(128, 160, 32, B=8, scale=1.0)),
             ("vis-s3x8", lambda: vis_case(256, 320, 16, B=8, scale=0.5)),
             ("cvp-l4", lambda: mvs_case(16, 74, 100, 96, L.AGG_VARIANCE_MEAN, False)),      # cfg4 coarsest level: scalar hypotheses
             ("cvp-l0", lambda: mvs_case(16, 1184, 1600, 8, L.AGG_VARIANCE_MEAN, True)),
             ("cfg1", lambda: mvs_case(32, 128, 160, 48, L.AGG_SOFTMIN, False, views=3))]
    only = os.environ.get("K1AB_CASES")     # comma-separated case names (default: all)
    for name, mk in cases:
        if only and name not in only.split(","):
            continue
        try:
            med, mn = timed(mk())
            print("%-40s %-9s %.4f ms (min %.4f)" % (os.path.basename(so), name, med, mn), flush=True)
        except Exception as e:  # noqa: BLE001 -- an experiment script: report and go on
            print("%-40s %-9s FAILED %r" % (os.path.basename(so), name, e), flush=True)


if __name__ == "__main__":
    if len(sys.argv) >= 3 and sys.argv[1] == "--child":
        child(sys.argv[2])
    else:
        for so in sys.argv[1:]:
            subprocess.run([sys.executable, os.path.abspath(__file__), "--child", so], check=False)
