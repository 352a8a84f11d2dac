"""K1 / K3 backward timings next to torch.autograd through the torch port on the same GPU (the existing Blackwell path of
the reference: grid_sample forward + grid_sampler_2d_backward and the materialised per-view volumes).  GPU box only.

    python tests/perf_backward.py [out.json]
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import torch_port as tp  # noqa: E402  (baseline leg; lives under tests/ because it times the oracle next to the kernels)
from wild_deep_mvs_b200 import _lib as L, ops, synth  # noqa: E402
from wild_deep_mvs_b200.mvsnet import build_proj_matrices  # noqa: E402

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


def case(name, agg, V, D, H=128, W=160, C=32):
    gen = torch.Generator().manual_seed(0)
    feats = [(0.5 * torch.randn(1, C, H, W, generator=gen)).to(DEV) for _ in range(V)]
    K, R, t, dmin, dmax = synth.make_cameras(1, V, 4 * H, 4 * W)
    K = K.clone()
    K[:, :, :2] /= 4
    proj = build_proj_matrices(K, R, t).to(DEV)
    dv = (dmin[:, :1] + (dmax[:, :1] - dmin[:, :1]) / (D - 1) * torch.arange(D).view(1, -1)).to(DEV)
    temp = torch.tensor([0.37], device=DEV)
    warp = ops.mvs_relative_proj(proj[:, 0].contiguous(), proj[:, 1:].contiguous())
    f = [ops.to_nhwc(x) for x in feats]
    code = L.AGG_SOFTMIN if agg == "softmin" else L.AGG_VARIANCE
    G = torch.randn(1, D, H, W, C, generator=gen).to(DEV)
    fwd = timed(lambda: ops.build_cost_volume(f[0], f[1:], warp, dv, D, L.GEOM_MVS, code, temp=temp))
    bwd = timed(lambda: ops.build_cost_volume_backward(G, f[0], f[1:], warp, dv, D, L.GEOM_MVS, code, temp=temp))
    res = {"case": name, "voxels": D * H * W, "k1_forward_ms": round(fwd, 3), "k1_backward_ms": round(bwd, 3)}

    def torch_step():
        fr = [x.detach().requires_grad_(True) for x in feats]
        tt = temp.detach().requires_grad_(True)
        vol = tp.mvsnet_cost_volume_train(fr[0], fr[1:], proj[:, 0], [proj[:, i] for i in range(1, V)], dv, agg, tt)
        vol.backward(G.permute(0, 4, 1, 2, 3))
    try:
        torch.cuda.reset_peak_memory_stats()
        res["torch_autograd_fwd_bwd_ms"] = round(timed(torch_step, reps=3, warmup=1), 3)
        res["torch_autograd_peak_GB"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    except torch.OutOfMemoryError:
        res["torch_autograd_fwd_bwd_ms"] = None
    score = torch.randn(1, D, H, W, generator=gen).to(DEV)
    gd = torch.randn(1, H, W, generator=gen).to(DEV)
    res["k3_backward_ms"] = round(timed(lambda: ops.depth_regress_backward(gd, score, dv)), 3)
    print(json.dumps(res), flush=True)
    return res


def regulariser_step(D=48, H=128, W=160):
    """Forward + backward of MVSNet's CostRegNet in training mode at the cfg1 volume: the PyTorch modules on cuDNN against
    MVSB200_TRAIN_K2=lib (K2 engines forward / input gradient, mvsb200_conv3d_wgrad)."""
    from wild_deep_mvs_b200.mvsnet import CostRegNet
    torch.manual_seed(0)
    net = CostRegNet().to(DEV).train()
    x = torch.randn(1, D, H, W, 32, device=DEV).permute(0, 4, 1, 2, 3)      # channels-last volume, reference view [B,C,D,H,W]
    res = {"case": "CostRegNet training step, volume %dx%dx%dx32" % (D, H, W)}
    for name, env, tf32 in (("cudnn_tf32_ms", None, True), ("cudnn_fp32_ms", None, False), ("lib_ms", "lib", False)):
        os.environ.pop("MVSB200_TRAIN_K2", None)
        if env:
            os.environ["MVSB200_TRAIN_K2"] = env
        torch.backends.cudnn.allow_tf32 = tf32

        def step():
            net.zero_grad(set_to_none=True)
            xi = x.detach().requires_grad_(True)
            net(xi).square().mean().backward()
        res[name] = round(timed(step, reps=3, warmup=2), 3)
    os.environ.pop("MVSB200_TRAIN_K2", None)
    print(json.dumps(res), flush=True)
    return res


if __name__ == "__main__":
    out = [case("cfg1 MVSNet-s soft-min 1+2 views D=48", "softmin", 3, 48), case("cfg2 MVSNet variance 1+4 views D=192", "variance", 5, 192),
           regulariser_step()]
    if len(sys.argv) > 1:
        json.dump(out, open(sys.argv[1], "w"), indent=1)
