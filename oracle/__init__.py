"""CPU oracle for the plane-sweep hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
may import this package.  The product (wild_deep_mvs_b200) never does.

Two layers:
  * oracle/mvs_oracle.c   plain-C fp32 restatement of each reference operator (ctypes
                          wrappers below, numpy in / numpy out, reference layouts NCHW/NCDHW)
  * oracle/nets.py        the three regularisation networks + model forwards composed from
                          those primitives and a state_dict of numpy arrays
  * oracle/torch_port.py  the same path restated with the ATen calls the reference makes
                          (multi-threaded CPU baseline timed by bench.py)

Parity pinning: the reference ships no tests or golden vectors, so the oracle
is pinned against tensors produced by importing the reference itself in the build
container (tests/golden/make_golden.py -> tests/golden/*.npz, checked by
tests/test_oracle_golden.py).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmvs_oracle.so")
_lib = None

_f32p = ctypes.POINTER(ctypes.c_float)


def build(force=False):
    src = os.path.join(_HERE, "mvs_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(_f32p) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _parr(arrs):
    return (_f32p * len(arrs))(*[_p(a) for a in arrs])


def relative_proj(src_proj, ref_proj):
    """proj = src_proj @ inv(ref_proj) -> (rot[3,3], trans[3]); MVSNet/module.py:128."""
    rot = np.empty((3, 3), np.float32)
    trans = np.empty(3, np.float32)
    rc = lib().orc_relative_proj(_p(_c(src_proj)), _p(_c(ref_proj)), _p(rot), _p(trans))
    if rc:
        raise ValueError("singular reference projection")
    return rot, trans


def homo_warp_mvs(src, rot, trans, depth, ref_hw):
    """src [C,Hs,Ws]; depth [D] or [D,H,W] -> [C,D,H,W]; MVSNet/module.py:111-169."""
    src = _c(src)
    depth = _c(depth)
    C, Hs, Ws = src.shape
    H, W = ref_hw
    D = depth.shape[0]
    out = np.empty((C, D, H, W), np.float32)
    lib().orc_homo_warp_mvs(_p(src), C, Hs, Ws, _p(_c(rot)), _p(_c(trans)), _p(depth),
                            int(depth.ndim == 3), D, H, W, _p(out))
    return out


def variance(ref, warped, order=0):
    """ref [C,H,W]; warped list of [C,D,H,W]; MVSNet/model.py:113-139 (order 0), CVP (order 1)."""
    ref = _c(ref)
    warped = [_c(w) for w in warped]
    C, D, H, W = warped[0].shape
    out = np.empty((C, D, H, W), np.float32)
    lib().orc_variance(_p(ref), _parr(warped), len(warped), C, D, H, W, order, _p(out))
    return out


def softmin(ref, warped, temp):
    """MVSNet-s aggregation; MVSNet/model.py:141-173."""
    ref = _c(ref)
    warped = [_c(w) for w in warped]
    C, D, H, W = warped[0].shape
    out = np.empty((C, D, H, W), np.float32)
    lib().orc_softmin(_p(ref), _parr(warped), len(warped), C, D, H, W, ctypes.c_float(temp), _p(out))
    return out


def vis_warp(src, ref_cam, src_cam, depth_start, depth_interval, D, ref_hw):
    """Vis-MVSNet homography warp; homography.py:23-121. cams [2,4,4] already scaled."""
    src = _c(src)
    C, Hs, Ws = src.shape
    H, W = ref_hw
    ds = _c(depth_start).reshape(-1)
    per_pixel = int(ds.size > 1)
    if per_pixel:
        assert ds.size == H * W
    out = np.empty((C, D, H, W), np.float32)
    rc = lib().orc_vis_warp(_p(src), C, Hs, Ws, _p(_c(ref_cam)), _p(_c(src_cam)), _p(ds), per_pixel,
                            ctypes.c_float(depth_interval), D, H, W, _p(out))
    if rc == -2:
        raise Exception("Nan")  # homography.py:71-72
    if rc:
        raise ValueError("singular intrinsics")
    return out


def groupcorr(ref, warped, groups=8):
    """nn_utils.py:473-490."""
    ref = _c(ref)
    warped = _c(warped)
    C, D, H, W = warped.shape
    out = np.empty((groups, D, H, W), np.float32)
    lib().orc_groupcorr(_p(ref), _p(warped), C, groups, D, H, W, _p(out))
    return out


def conv3d(x, w, bias=None, stride=1, pad=None):
    """nn.Conv3d, x [Cin,D,H,W], w [Cout,Cin,kd,kh,kw]."""
    x = _c(x)
    w = _c(w)
    Cin, D, H, W = x.shape
    Cout, _, kd, kh, kw = w.shape
    if pad is None:
        pad = (kd // 2, kh // 2, kw // 2)
    elif isinstance(pad, int):
        pad = (pad, pad, pad)
    Do = (D + 2 * pad[0] - kd) // stride + 1
    Ho = (H + 2 * pad[1] - kh) // stride + 1
    Wo = (W + 2 * pad[2] - kw) // stride + 1
    out = np.empty((Cout, Do, Ho, Wo), np.float32)
    b = _c(bias) if bias is not None else None
    lib().orc_conv3d(_p(x), Cin, D, H, W, _p(w), _p(b), Cout, kd, kh, kw, stride, pad[0], pad[1],
                     pad[2], _p(out))
    return out


def deconv3d(x, w, bias=None, stride=2, pad=1, outpad=1):
    """nn.ConvTranspose3d, x [Cin,D,H,W], w [Cin,Cout,k,k,k]."""
    x = _c(x)
    w = _c(w)
    Cin, D, H, W = x.shape
    _, Cout, k, _, _ = w.shape
    Do, Ho, Wo = [(n - 1) * stride - 2 * pad + k + outpad for n in (D, H, W)]
    out = np.empty((Cout, Do, Ho, Wo), np.float32)
    b = _c(bias) if bias is not None else None
    lib().orc_deconv3d(_p(x), Cin, D, H, W, _p(w), _p(b), Cout, k, stride, pad, outpad, _p(out))
    return out


def bn_relu(x, gamma, beta, mean, var, eps=1e-5, relu=True):
    """eval-mode BatchNorm (+ReLU); returns a new array."""
    x = _c(x).copy()
    C = x.shape[0]
    n = x.size // C
    lib().orc_bn_relu(_p(x), C, ctypes.c_size_t(n), _p(_c(gamma)), _p(_c(beta)), _p(_c(mean)),
                      _p(_c(var)), ctypes.c_float(eps), int(relu))
    return x


def softmax_regress(score, depth_mode, dvals, interval=0.0, conf_mode=0, want_entropy=False,
                    want_prob=False):
    """score [D,H,W] -> dict(depth, conf, entropy, prob); see mvs_oracle.c for the modes."""
    score = _c(score)
    D, H, W = score.shape
    dv = _c(dvals).reshape(-1)
    per_pixel = int(depth_mode == 2 and dv.size > 1)
    depth = np.empty((H, W), np.float32)
    conf = np.empty((H, W), np.float32) if conf_mode else None
    ent = np.empty((H, W), np.float32) if want_entropy else None
    prob = np.empty((D, H, W), np.float32) if want_prob else None
    lib().orc_softmax_regress(_p(score), D, H, W, depth_mode, _p(dv), per_pixel,
                              ctypes.c_float(interval), conf_mode, _p(depth), _p(conf), _p(ent),
                              _p(prob))
    return {"depth": depth, "conf": conf, "entropy": ent, "prob": prob}


def vis_fuse(interm, uncert):
    """interm list of [G,D,H,W], uncert list of [H,W]; model_cas.py:354-357,385-386."""
    interm = [_c(a) for a in interm]
    uncert = [_c(a) for a in uncert]
    G, D, H, W = interm[0].shape
    out = np.empty((G, D, H, W), np.float32)
    lib().orc_vis_fuse(_parr(interm), _parr(uncert), len(interm), G, D, H, W, _p(out))
    return out
