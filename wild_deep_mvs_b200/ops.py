"""Tensor <-> pointer marshalling for libmvsb200.so.  PyTorch provides device memory and the
current stream; every computation below happens inside the library's CUDA kernels.

Layouts (include/mvsb200.h): feature maps [B,H,W,C], volumes [B,D,H,W,C], single-channel volumes
[B,D,H,W].  `to_nhwc` / `to_ndhwc` turn reference-layout tensors (NCHW / NCDHW) into these without a
copy when the tensor is already channels-last in memory.
"""
import ctypes
import functools
import os

import torch

from . import _lib as L

# K2 engine:
#   "zm"      = z-march tcgen05 kernel (kind::f16, error-compensated fp16 split after abs-max scaling: fp32-equivalent),
#               the parity-tested default; layers it does not cover fall through to "tc", then to "fp32";
#   "tc"      = tile-at-a-time tcgen05 kernel with 3xTF32 split products (fp32-equivalent);
#   "tc_tf32" = the same kernel, single-pass TF32 (outside the parity bar);
#   "fp32"    = the CUDA-core kernel (also runs every 1x1x1 / 2-D / Cin % 8 != 0 layer).
ENGINES = ("zm", "tc", "tc_tf32", "fp32")
DEFAULT_ENGINE = os.environ.get("MVSB200_K2_ENGINE", "zm")
# programmatic dependent launch of the z-march engine (MVSB200_PDL=0 switches it off: A/B runs)
PDL = os.environ.get("MVSB200_PDL", "1") != "0"


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _tensors(objs):
    for o in objs:
        if isinstance(o, torch.Tensor):
            yield o
        elif isinstance(o, (list, tuple)):
            yield from _tensors(o)


def _on_operand_device(fn):
    """Every library call launches on the CURRENT device's current stream (and `cudaGetDevice` picks the kernel
    attributes / grid sizes).  This guard makes the operands' device current for the duration of the call and refuses
    operands that live on different devices, so a net moved to cuda:1 works while cuda:0 is current."""
    @functools.wraps(fn)
    def guarded(*args, **kwargs):
        dev = None
        for t in _tensors(list(args) + list(kwargs.values())):
            if t.is_cuda:
                if dev is None:
                    dev = t.device
                elif t.device != dev:
                    raise L.Mvsb200Error("%s: operands on different devices (%s and %s)" % (fn.__name__, dev, t.device))
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return guarded


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _dev_f32(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
        raise L.Mvsb200Error("%s must be a contiguous float32 CUDA tensor" % name)
    return t


class AmaxPool:
    """Zero-initialised device scalars for the abs-max tracking of the z-march engine: ONE fill per pool instead of
    one per layer.  take() hands out the next 1-element view."""

    def __init__(self, device, n=32):
        self.buf = torch.zeros(n, device=device, dtype=torch.float32)
        self.next = 0

    def take(self):
        if self.next >= self.buf.numel():
            self.buf = torch.zeros_like(self.buf)
            self.next = 0
        v = self.buf[self.next:self.next + 1]
        self.next += 1
        return v


@_on_operand_device
def absmax(t):
    """max|t| as a 1-element device tensor: the value its producer tracked (attribute `_mvs_amax`, set by
    build_cost_volume / conv3d / vis_fuse) or, for tensors from elsewhere, one reduction pass (mvsb200_absmax).
    The tracked value is tied to the tensor's autograd version counter (`_mvs_amax_version`): an in-place torch write
    (`vol.copy_(new)`) after the value was taken invalidates it and the maximum is recomputed."""
    a = getattr(t, "_mvs_amax", None)
    if a is not None and getattr(t, "_mvs_amax_version", t._version) == t._version:
        return a
    a = torch.empty(1, device=t.device, dtype=torch.float32)
    L.check(L.load().mvsb200_absmax(_ptr(t), t.numel(), _ptr(a), _stream()), "mvsb200_absmax")
    t._mvs_amax = a
    t._mvs_amax_version = t._version
    return a


def set_absmax(t, a):
    """Attach a producer-tracked abs-max scalar to `t` (valid until the next in-place torch write to `t`)."""
    t._mvs_amax = a
    t._mvs_amax_version = t._version
    return t


def map_views(extract, imgs):
    """Run a 2-D feature extractor over the views of a sample.  The reference calls it once per view
    (models/MVSNet/model.py:195, VisMVSNet/frontend.py:59-62, CVP_MVSNet/models/net.py:110-114); when the views have the
    same size they go through as ONE batch (same weights, per-sample arithmetic unchanged), which cuts the launch count
    of these small CNNs by the number of views.  `extract` returns a tensor or a list/tuple of tensors; the result is a
    list with one such item per view."""
    imgs = list(imgs)
    if len(imgs) > 1 and all(im.shape == imgs[0].shape for im in imgs[1:]):
        b = imgs[0].shape[0]
        out = extract(torch.cat(imgs, 0))
        if isinstance(out, torch.Tensor):
            return list(torch.split(out, b, 0))
        per_level = [torch.split(o, b, 0) for o in out]
        return [type(out)(lvl[v] for lvl in per_level) if isinstance(out, tuple) else [lvl[v] for lvl in per_level]
                for v in range(len(imgs))]
    return [extract(im) for im in imgs]


def to_nhwc(x):
    """[B,C,H,W] (any strides) -> contiguous [B,H,W,C]."""
    return x.permute(0, 2, 3, 1).contiguous()


def to_ndhwc(x):
    """[B,C,D,H,W] (any strides) -> contiguous [B,D,H,W,C]."""
    return x.permute(0, 2, 3, 4, 1).contiguous()


def as_ncdhw(v):
    """[B,D,H,W,C] -> reference-shaped view [B,C,D,H,W] (no copy)."""
    return v.permute(0, 4, 1, 2, 3)


# ------------------------------------------------------------------------------------------------
# geometry prologues
# ------------------------------------------------------------------------------------------------
@_on_operand_device
def mvs_relative_proj(ref_proj, src_projs):
    """ref_proj [B,4,4], src_projs [B,S,4,4] -> warp [B,S,16]; replaces MVSNet/module.py:128."""
    ref_proj = _dev_f32(ref_proj.contiguous(), "ref_proj")
    src_projs = _dev_f32(src_projs.contiguous(), "src_projs")
    B, S = src_projs.shape[:2]
    warp = torch.empty(B, S, 16, device=ref_proj.device, dtype=torch.float32)
    L.check(L.load().mvsb200_mvs_relative_proj(_ptr(ref_proj), _ptr(src_projs), _ptr(warp), B, S, _stream()),
            "mvsb200_mvs_relative_proj")
    return warp


@_on_operand_device
def vis_homography_params(ref_cam, src_cams, scale):
    """ref_cam [B,2,4,4], src_cams [B,S,2,4,4] -> warp [B,S,16]; replaces VisMVSNet/homography.py:23-74."""
    ref_cam = _dev_f32(ref_cam.contiguous(), "ref_cam")
    src_cams = _dev_f32(src_cams.contiguous(), "src_cams")
    B, S = src_cams.shape[:2]
    warp = torch.empty(B, S, 16, device=ref_cam.device, dtype=torch.float32)
    L.check(L.load().mvsb200_vis_homography_params(_ptr(ref_cam), _ptr(src_cams), ctypes.c_float(scale), _ptr(warp),
                                                   B, S, _stream()), "mvsb200_vis_homography_params")
    return warp


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
def _depth_mode(depth, interval, B, D, H, W):
    if interval is None:
        if depth.dim() == 2:
            assert depth.shape == (B, D), depth.shape
            return L.DEPTH_VALUES
        assert depth.shape == (B, D, H, W), depth.shape
        return L.DEPTH_VOLUME
    assert interval.numel() == B
    if depth.numel() == B:
        return L.DEPTH_START
    assert depth.numel() == B * H * W, depth.shape
    return L.DEPTH_START_MAP


@_on_operand_device
def build_cost_volume(ref, srcs, warp, depth, D, geom, agg, interval=None, temp=None, groups=8, out=None, amax=None):
    """ref [B,H,W,C]; srcs list of [B,Hs,Ws,C]; warp [B,S,16]; depth/interval see mvsb200.h.
    Returns [B,D,H,W,C] or, for AGG_GROUPCORR, [S,B,D,H,W,groups].  `amax` (zeroed device scalar, optional) receives
    max|out| for the z-march conv engine; `out` lets the caller own the output buffer."""
    lib = L.load()
    ref = _dev_f32(ref, "ref")
    B, H, W, C = ref.shape
    S = len(srcs)
    if not 1 <= S <= L.MAX_SRC:
        raise L.Mvsb200Error("number of source views %d not in [1,%d]" % (S, L.MAX_SRC))
    depth = _dev_f32(depth.contiguous(), "depth")
    if interval is not None:
        interval = _dev_f32(interval.contiguous().view(-1), "interval")
    desc = L.CostVolumeDesc()
    desc.geom, desc.agg = geom, agg
    desc.depth_mode = _depth_mode(depth, interval, B, D, H, W)
    desc.B, desc.S, desc.C, desc.D, desc.H, desc.W = B, S, C, D, H, W
    desc.groups = groups if agg == L.AGG_GROUPCORR else 0
    ptrs = (ctypes.c_void_p * S)()
    for i, s in enumerate(srcs):
        _dev_f32(s, "src[%d]" % i)
        if s.shape[0] != B or s.shape[3] != C:
            raise L.Mvsb200Error("src[%d] has shape %s, expected [%d,*,*,%d]" % (i, tuple(s.shape), B, C))
        ptrs[i] = s.data_ptr()
        desc.src_h[i], desc.src_w[i] = s.shape[1], s.shape[2]
    shape = (S, B, D, H, W, groups) if agg == L.AGG_GROUPCORR else (B, D, H, W, C)
    desc.out_view_stride = B * D * H * W * groups if agg == L.AGG_GROUPCORR else 0
    if out is None:
        out = torch.empty(shape, device=ref.device, dtype=torch.float32)
    elif tuple(_dev_f32(out, "out").shape) != shape:
        raise L.Mvsb200Error("build_cost_volume: out has shape %s, expected %s" % (tuple(out.shape), shape))
    if temp is not None:
        temp = _dev_f32(temp.detach().contiguous(), "temp")
    if amax is None:
        amax = torch.zeros(1, device=ref.device, dtype=torch.float32)
    L.check(lib.mvsb200_build_cost_volume(ctypes.byref(desc), _ptr(ref), ptrs, _ptr(_dev_f32(warp, "warp")), _ptr(depth),
                                          _ptr(interval), _ptr(temp), _ptr(out), _ptr(amax), _stream()),
            "mvsb200_build_cost_volume")
    set_absmax(out, amax)
    return out


def _cost_volume_desc(ref, srcs, D, geom, agg, groups, depth, interval):
    B, H, W, C = ref.shape
    S = len(srcs)
    if not 1 <= S <= L.MAX_SRC:
        raise L.Mvsb200Error("number of source views %d not in [1,%d]" % (S, L.MAX_SRC))
    desc = L.CostVolumeDesc()
    desc.geom, desc.agg = geom, agg
    desc.depth_mode = _depth_mode(depth, interval, B, D, H, W)
    desc.B, desc.S, desc.C, desc.D, desc.H, desc.W = B, S, C, D, H, W
    desc.groups = groups if agg == L.AGG_GROUPCORR else 0
    ptrs = (ctypes.c_void_p * S)()
    for i, s in enumerate(srcs):
        _dev_f32(s, "src[%d]" % i)
        if s.shape[0] != B or s.shape[3] != C:
            raise L.Mvsb200Error("src[%d] has shape %s, expected [%d,*,*,%d]" % (i, tuple(s.shape), B, C))
        ptrs[i] = s.data_ptr()
        desc.src_h[i], desc.src_w[i] = s.shape[1], s.shape[2]
    desc.out_view_stride = B * D * H * W * groups if agg == L.AGG_GROUPCORR else 0
    return desc, ptrs


@_on_operand_device
def build_cost_volume_backward(grad_out, ref, srcs, warp, depth, D, geom, agg, interval=None, temp=None, groups=8):
    """Gradient of build_cost_volume with respect to the feature maps (and `temp` for AGG_SOFTMIN): K1 backward
    (mvsb200_build_cost_volume_backward).  Returns (grad_ref [B,H,W,C], [grad_src_s], grad_temp or None)."""
    lib = L.load()
    ref = _dev_f32(ref, "ref")
    depth = _dev_f32(depth.contiguous(), "depth")
    if interval is not None:
        interval = _dev_f32(interval.contiguous().view(-1), "interval")
    desc, ptrs = _cost_volume_desc(ref, srcs, D, geom, agg, groups, depth, interval)
    B, H, W, C = ref.shape
    shape = (len(srcs), B, D, H, W, groups) if agg == L.AGG_GROUPCORR else (B, D, H, W, C)
    grad_out = grad_out.contiguous()
    if tuple(_dev_f32(grad_out, "grad_out").shape) != shape:
        raise L.Mvsb200Error("build_cost_volume_backward: grad_out has shape %s, expected %s" % (tuple(grad_out.shape), shape))
    g_ref = torch.zeros_like(ref)
    g_srcs = [torch.zeros_like(s) for s in srcs]
    gptrs = (ctypes.c_void_p * len(srcs))(*[g.data_ptr() for g in g_srcs])
    g_temp = None
    if agg == L.AGG_SOFTMIN:
        temp = _dev_f32(temp.detach().contiguous(), "temp")
        g_temp = torch.zeros_like(temp)
    L.check(lib.mvsb200_build_cost_volume_backward(ctypes.byref(desc), _ptr(ref), ptrs, _ptr(_dev_f32(warp, "warp")), _ptr(depth),
                                                   _ptr(interval), _ptr(temp), _ptr(grad_out), _ptr(g_ref), gptrs, _ptr(g_temp),
                                                   _stream()), "mvsb200_build_cost_volume_backward")
    return g_ref, g_srcs, g_temp


class _CostVolumeFn(torch.autograd.Function):
    """build_cost_volume as a differentiable op: K1 forward, K1 backward.  Like the reference (sampling grid under
    torch.no_grad(): models/MVSNet/module.py:127, VisMVSNet/homography.py:25,110) the gradient flows to the feature maps
    (and the soft-min temperature) only."""

    @staticmethod
    def forward(ctx, cfg, warp, depth, interval, temp, ref, *srcs):
        D, geom, agg, groups = cfg
        ref = ref.contiguous()
        srcs = [s.contiguous() for s in srcs]
        out = build_cost_volume(ref, srcs, warp, depth, D, geom, agg, interval=interval, temp=temp, groups=groups)
        ctx.cfg = cfg
        ctx.save_for_backward(warp, depth, interval, temp, ref, *srcs)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        warp, depth, interval, temp, ref, *srcs = ctx.saved_tensors
        D, geom, agg, groups = ctx.cfg
        g_ref, g_srcs, g_temp = build_cost_volume_backward(grad_out, ref, list(srcs), warp, depth, D, geom, agg,
                                                           interval=interval, temp=temp, groups=groups)
        if g_temp is not None:
            g_temp = g_temp.view(ctx.saved_tensors[3].shape)
        return (None, None, None, None, g_temp, g_ref, *g_srcs)


def cost_volume(ref, srcs, warp, depth, D, geom, agg, interval=None, temp=None, groups=8):
    """Differentiable build_cost_volume (training, row f2): same arguments and result, autograd-aware."""
    return _CostVolumeFn.apply((D, geom, agg, groups), warp, depth, interval, temp, ref, *srcs)


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
class PackedConv:
    """A conv / transposed conv layer packed for mvsb200_conv3d: tap-major weights [taps][Cin][Cout],
    folded eval-mode BatchNorm as (scale, bias), and the fused epilogue flags."""

    def __init__(self, weight, bn=None, conv_bias=None, stride=1, transposed=False, relu=False,
                 skip_mode=L.SKIP_NONE):
        w = weight.detach().float()
        if w.dim() == 4:  # 2-D conv (UncertNet): [Cout,Cin,kh,kw] -> kd = 1
            w = w.unsqueeze(2)
        if transposed and stride == 1:
            # ConvTranspose3d(k=3, stride=1, padding=1) == Conv3d with the kernel flipped and channels swapped
            # (CVP conv5, CVP_MVSNet/models/net.py:62-65)
            w = w.flip(2, 3, 4).transpose(0, 1)
            transposed = False
        if transposed:  # [Cin,Cout,k,k,k]
            self.cin, self.cout = w.shape[0], w.shape[1]
            packed = w.permute(2, 3, 4, 0, 1)
        else:           # [Cout,Cin,kd,kh,kw]
            self.cout, self.cin = w.shape[0], w.shape[1]
            packed = w.permute(2, 3, 4, 1, 0)
        self.k = tuple(w.shape[2:])
        self.w = packed.contiguous()
        self.stride, self.transposed, self.relu, self.skip_mode = stride, int(transposed), int(relu), skip_mode
        if bn is not None:
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
            if conv_bias is not None:
                bias = bias + conv_bias.detach().float() * scale
            self.scale, self.bias = scale.contiguous(), bias.contiguous()
        else:
            self.scale = None
            self.bias = conv_bias.detach().float().contiguous() if conv_bias is not None else None
        self._tc_packed = None
        self._zm_packed = None
        self._c1_host = None
        self._halves = None
        self._launches = 0   # z-march launches so far: from the second on, the packed parameters predate the stream's tail

    def halves(self):
        """The layer split over its input channels into two layers whose sum is this layer (z-march engine, layers whose
        weights do not fit its shared memory): A = conv(x[:h]) * scale + bias (+ skip before the activation),
        B = act(conv(x[h:]) * scale + A).  Only layers without a skip after the activation can be split this way."""
        if self._halves is None:
            h = self.cin // 2
            a, b = object.__new__(PackedConv), object.__new__(PackedConv)
            for part, sl in ((a, slice(0, h)), (b, slice(h, self.cin))):
                part.w = self.w[..., sl, :].contiguous()
                part.cin, part.cout, part.k = h, self.cout, self.k
                part.stride, part.transposed = self.stride, self.transposed
                part.scale = self.scale
                part._tc_packed = part._zm_packed = part._c1_host = part._halves = None
                part._launches = 0
            a.bias, a.relu, a.skip_mode = self.bias, 0, self.skip_mode
            b.bias, b.relu, b.skip_mode = None, self.relu, L.SKIP_BEFORE_RELU
            self._halves = (a, b)
        return self._halves

    def c1_host(self):
        """Host copies for the single-output-channel head kernel: (ctypes float array [27*Cin], scale, bias)."""
        if self._c1_host is None:
            w = self.w.reshape(-1).cpu()
            arr = (ctypes.c_float * w.numel())(*w.tolist())
            scale = float(self.scale.cpu()[0]) if self.scale is not None else 1.0
            bias = float(self.bias.cpu()[0]) if self.bias is not None else 0.0
            self._c1_host = (arr, scale, bias)
        return self._c1_host

    def zm_packed(self, desc):
        """Weights scaled by a power of two, split into two fp16 pieces and laid out as kind::f16 B operands for the
        z-march engine (packed once, on the device)."""
        if self._zm_packed is None:
            lib = L.load()
            n = lib.mvsb200_conv3d_zm_packed_bytes(ctypes.byref(desc))
            buf = torch.empty((n + 3) // 4, device=self.w.device, dtype=torch.float32)
            L.check(lib.mvsb200_conv3d_zm_pack(ctypes.byref(desc), _ptr(self.w), _ptr(buf), _stream()), "mvsb200_conv3d_zm_pack")
            self._zm_packed = buf
        return self._zm_packed

    def tc_packed(self, desc):
        """Weights split into tf32 hi/lo and laid out as tcgen05 B operands (packed once, on the device)."""
        if self._tc_packed is None:
            lib = L.load()
            n = lib.mvsb200_conv3d_tc_packed_floats(ctypes.byref(desc))
            buf = torch.empty(n, device=self.w.device, dtype=torch.float32)
            L.check(lib.mvsb200_conv3d_tc_pack(ctypes.byref(desc), _ptr(self.w), _ptr(buf), _stream()), "mvsb200_conv3d_tc_pack")
            self._tc_packed = buf
        return self._tc_packed


@_on_operand_device
def conv3d(x, layer, x2=None, skip=None, engine=None, out=None, amax=None):
    """x [B,D,H,W,Cin] (+ x2 [B,D,H,W,Cin2] concatenated on channels) -> [B,Do,Ho,Wo,Cout].
    `out`: caller-owned output buffer; `amax`: zeroed device scalar that receives max|y| (z-march engine)."""
    lib = L.load()
    engine = engine or DEFAULT_ENGINE
    if engine not in ENGINES:
        raise L.Mvsb200Error("conv3d: unknown engine %r (expected one of %s)" % (engine, ", ".join(ENGINES)))
    x = _dev_f32(x, "x")
    B, D, H, W, C1 = x.shape
    C2 = 0
    if x2 is not None:
        _dev_f32(x2, "x2")
        assert x2.shape[:4] == x.shape[:4]
        C2 = x2.shape[4]
    if C1 + C2 != layer.cin:
        raise L.Mvsb200Error("conv3d: input has %d channels, layer expects %d" % (C1 + C2, layer.cin))
    desc = L.Conv3dDesc()
    desc.B, desc.D, desc.H, desc.W = B, D, H, W
    desc.Cin, desc.Cin2, desc.Cout = C1, C2, layer.cout
    desc.kd, desc.kh, desc.kw = layer.k
    desc.stride, desc.transposed, desc.relu = layer.stride, layer.transposed, layer.relu
    desc.skip_mode = layer.skip_mode if skip is not None else L.SKIP_NONE
    if layer.skip_mode != L.SKIP_NONE and skip is None:
        raise L.Mvsb200Error("conv3d: layer fuses a skip connection but none was given")
    do, ho, wo = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    L.check(lib.mvsb200_conv3d_out_shape(ctypes.byref(desc), ctypes.byref(do), ctypes.byref(ho), ctypes.byref(wo)),
            "mvsb200_conv3d_out_shape")
    yshape = (B, do.value, ho.value, wo.value, layer.cout)
    if out is None:
        y = torch.empty(yshape, device=x.device, dtype=torch.float32)
    else:
        y = _dev_f32(out, "out")
        if tuple(y.shape) != yshape:
            raise L.Mvsb200Error("conv3d: out has shape %s, expected %s" % (tuple(y.shape), yshape))
    if skip is not None:
        _dev_f32(skip, "skip")
        assert skip.shape == y.shape, (skip.shape, y.shape)
    if engine in ("zm", "tc") and x2 is None and skip is None and lib.mvsb200_conv3d_c1_supported(ctypes.byref(desc)):
        # Cout == 1 heads: CUDA cores, fp32, weights as launch parameters
        w_host, scale, bias = layer.c1_host()
        L.check(lib.mvsb200_conv3d_c1(ctypes.byref(desc), _ptr(x), w_host, ctypes.c_float(scale), ctypes.c_float(bias), _ptr(y),
                                      _stream()), "mvsb200_conv3d_c1")
        return y
    if engine == "zm" and x.device == layer.w.device and lib.mvsb200_conv3d_zm_supported(ctypes.byref(desc)):
        if amax is None:
            amax = torch.zeros(1, device=x.device, dtype=torch.float32)
        # parameters packed by an earlier call are not outputs of the kernels in front of this launch: the engine may
        # overlap its prologue with them (programmatic dependent launch, see mvsb200.h)
        desc.static_params = 1 if (PDL and layer._launches > 0) else 0
        layer._launches += 1
        L.check(lib.mvsb200_conv3d_zm(ctypes.byref(desc), _ptr(x), _ptr(x2), _ptr(layer.zm_packed(desc)), _ptr(layer.scale),
                                      _ptr(layer.bias), _ptr(skip), _ptr(y), _ptr(absmax(x)),
                                      _ptr(absmax(x2)) if x2 is not None else None, _ptr(amax), _stream()), "mvsb200_conv3d_zm")
        set_absmax(y, amax)
        return y
    if engine == "zm" and x.device == layer.w.device and x2 is None and C1 % 16 == 0 and layer.k == (3, 3, 3):
        # weights too large to stay resident: two launches over the halves of the input channels (see PackedConv.halves)
        half = L.Conv3dDesc.from_buffer_copy(desc)
        half.Cin = C1 // 2
        if lib.mvsb200_conv3d_zm_supported(ctypes.byref(half)):
            a, b = layer.halves()
            if amax is None:
                amax = torch.zeros(1, device=x.device, dtype=torch.float32)
            xa = _ptr(absmax(x))
            half.static_params = 1 if (PDL and a._launches > 0 and b._launches > 0) else 0
            a._launches += 1
            b._launches += 1
            # a skip that is added AFTER the activation cannot ride on either launch (the second one's skip operand is
            # the first one's partial sum): it is one in-place pass over the output afterwards (K7b)
            after = layer.skip_mode == L.SKIP_AFTER_RELU and skip is not None
            half.relu, half.skip_mode = a.relu, (a.skip_mode if (skip is not None and not after) else L.SKIP_NONE)
            L.check(lib.mvsb200_conv3d_zm_slice(ctypes.byref(half), _ptr(x), C1, 0, None, _ptr(a.zm_packed(half)), _ptr(a.scale),
                                                _ptr(a.bias), None if after else _ptr(skip), _ptr(y), xa, None, None, _stream()),
                    "mvsb200_conv3d_zm_slice")
            half.relu, half.skip_mode = b.relu, L.SKIP_BEFORE_RELU
            L.check(lib.mvsb200_conv3d_zm_slice(ctypes.byref(half), _ptr(x), C1, C1 // 2, None, _ptr(b.zm_packed(half)), _ptr(b.scale),
                                                None, _ptr(y), _ptr(y), xa, None, None if after else _ptr(amax), _stream()),
                    "mvsb200_conv3d_zm_slice")
            if after:
                L.check(lib.mvsb200_bias_act(_ptr(y), y.numel() // y.shape[-1], y.shape[-1], None, None, _ptr(skip), ctypes.c_float(1.0),
                                             _ptr(amax), _stream()), "mvsb200_bias_act")
            set_absmax(y, amax)
            return y
    if engine != "fp32" and x.device == layer.w.device and lib.mvsb200_conv3d_tc_supported(ctypes.byref(desc)):
        prec = L.PRECISION_TF32 if engine == "tc_tf32" else L.PRECISION_3XTF32
        L.check(lib.mvsb200_conv3d_tc(ctypes.byref(desc), _ptr(x), _ptr(x2), _ptr(layer.tc_packed(desc)), _ptr(layer.scale),
                                      _ptr(layer.bias), _ptr(skip), _ptr(y), prec, _stream()), "mvsb200_conv3d_tc")
        return y
    L.check(lib.mvsb200_conv3d(ctypes.byref(desc), _ptr(x), _ptr(x2), _ptr(layer.w), _ptr(layer.scale), _ptr(layer.bias),
                               _ptr(skip), _ptr(y), _stream()), "mvsb200_conv3d")
    return y


# ---- K2 in training (row f2): forward and input gradient on the K2 engines, weight gradient on mvsb200_conv3d_wgrad ----
@_on_operand_device
def conv3d_wgrad(a, b, stride):
    """R[ca][cb][3,3,3] = sum_v a[v,ca] * b[stride*v + tap - 1, cb]   (a [B,Da,Ha,Wa,Ca] the low-resolution side, b
    [B,Db,Hb,Wb,Cb] the high-resolution side).  Conv3d: a = grad_out, b = input -> dW [Cout,Cin,3,3,3];
    ConvTranspose3d: a = input, b = grad_out -> dW [Cin,Cout,3,3,3]."""
    a, b = _dev_f32(a.contiguous(), "a"), _dev_f32(b.contiguous(), "b")
    B, Da, Ha, Wa, Ca = a.shape
    Bb, Db, Hb, Wb, Cb = b.shape
    if B != Bb:
        raise L.Mvsb200Error("conv3d_wgrad: batch sizes %d and %d" % (B, Bb))
    r = torch.zeros(Ca, Cb, 3, 3, 3, device=a.device, dtype=torch.float32)
    L.check(L.load().mvsb200_conv3d_wgrad(_ptr(a), _ptr(b), B, Da, Ha, Wa, Ca, Db, Hb, Wb, Cb, stride, _ptr(r), _stream()),
            "mvsb200_conv3d_wgrad")
    return r


def conv3d_input_grad(grad_out, weight, stride, transposed):
    """Input gradient of a 3x3x3 layer as a FORWARD call of the K2 engines with the weights re-packed:
      Conv3d stride 1            -> conv with the kernel flipped and the channel axes swapped
      Conv3d stride 2            -> ConvTranspose3d(stride 2, output_padding 1) with the same weight tensor
      ConvTranspose3d stride 1/2 -> Conv3d of that stride with the same weight tensor ([Cin,Cout] read as [out,in])
    grad_out [B,Do,Ho,Wo,Cout_of_the_layer] channels-last; weight in its nn.Module layout."""
    w = weight.detach()
    if transposed:
        layer = PackedConv(w, None, stride=stride)
    elif stride == 1:
        layer = PackedConv(w.flip(2, 3, 4).transpose(0, 1), None)
    else:
        layer = PackedConv(w, None, stride=stride, transposed=True)
    return conv3d(grad_out.contiguous(), layer)


class _Conv3dFn(torch.autograd.Function):
    """A 3x3x3 Conv3d / ConvTranspose3d (+ bias) of the regularisers as a differentiable op on channels-last volumes:
    forward and input gradient on the K2 engines (tcgen05), weight gradient on the wgrad kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, transposed):
        x = x.contiguous()
        y = conv3d(x, PackedConv(weight, None, conv_bias=bias, stride=stride, transposed=transposed))
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, transposed, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        stride, transposed, has_bias = ctx.cfg
        gy = gy.contiguous()
        gx = conv3d_input_grad(gy, weight, stride, transposed) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            gw = conv3d_wgrad(x, gy, stride) if transposed else conv3d_wgrad(gy, x, stride)
        gb = gy.sum(dim=(0, 1, 2, 3)) if has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None


def conv3d_train(x, weight, bias=None, stride=1, transposed=False):
    """Differentiable 3x3x3 conv on a channels-last volume x [B,D,H,W,Cin] (training, row f2) -> [B,Do,Ho,Wo,Cout]."""
    return _Conv3dFn.apply(x, weight, bias, stride, transposed)


# ------------------------------------------------------------------------------------------------
# K3 / K4
# ------------------------------------------------------------------------------------------------
@_on_operand_device
def depth_regress(score, depth, interval=None, conf_mode=L.CONF_NONE, want_entropy=False, want_prob=False,
                  out_depth=None, out_conf=None):
    """score [B,D,H,W] -> dict(depth [B,H,W], conf, entropy, prob).  `out_depth` / `out_conf`: caller-owned [B,H,W]
    buffers the kernel writes into -- e.g. this rank's slice of a gather buffer (shard.DepthGather), so that the depth
    map is born where the all-gather sends it from."""
    lib = L.load()
    score = _dev_f32(score, "score")
    B, D, H, W = score.shape
    depth = _dev_f32(depth.contiguous(), "depth")
    if interval is not None:
        interval = _dev_f32(interval.contiguous().view(-1), "interval")
    mode = _depth_mode(depth, interval, B, D, H, W)
    new = lambda *s: torch.empty(*s, device=score.device, dtype=torch.float32)
    for name, t in (("out_depth", out_depth), ("out_conf", out_conf)):
        if t is not None and (tuple(_dev_f32(t, name).shape) != (B, H, W) or t.device != score.device):
            raise L.Mvsb200Error("depth_regress: %s has shape %s on %s, expected %s on %s"
                                 % (name, tuple(t.shape), t.device, (B, H, W), score.device))
    out = {"depth": out_depth if out_depth is not None else new(B, H, W),
           "conf": (out_conf if out_conf is not None else new(B, H, W)) if conf_mode else None,
           "entropy": new(B, H, W) if want_entropy else None, "prob": new(B, D, H, W) if want_prob else None}
    L.check(lib.mvsb200_depth_regress(_ptr(score), B, D, H, W, mode, _ptr(depth), _ptr(interval), conf_mode,
                                      _ptr(out["depth"]), _ptr(out["conf"]), _ptr(out["entropy"]), _ptr(out["prob"]),
                                      _stream()), "mvsb200_depth_regress")
    return out


@_on_operand_device
def depth_regress_backward(grad_depth, score, depth, interval=None, want_grad_hyp=False):
    """grad_score [B,D,H,W] of depth_regress()["depth"] (K3 backward, mvsb200_depth_regress_backward); with
    `want_grad_hyp` (per-voxel hypotheses [B,D,H,W] only) -> (grad_score, grad_hypotheses)."""
    lib = L.load()
    score = _dev_f32(score, "score")
    B, D, H, W = score.shape
    depth = _dev_f32(depth.contiguous(), "depth")
    if interval is not None:
        interval = _dev_f32(interval.contiguous().view(-1), "interval")
    mode = _depth_mode(depth, interval, B, D, H, W)
    grad_depth = _dev_f32(grad_depth.contiguous(), "grad_depth")
    if grad_depth.numel() != B * H * W:
        raise L.Mvsb200Error("depth_regress_backward: grad_depth has shape %s" % (tuple(grad_depth.shape),))
    g = torch.empty_like(score)
    gh = torch.empty_like(score) if want_grad_hyp else None
    L.check(lib.mvsb200_depth_regress_backward(_ptr(score), B, D, H, W, mode, _ptr(depth), _ptr(interval), _ptr(grad_depth),
                                               _ptr(g), _ptr(gh), _stream()), "mvsb200_depth_regress_backward")
    return (g, gh) if want_grad_hyp else g


class _DepthRegressFn(torch.autograd.Function):
    """softmax over D + expectation of the hypotheses as a differentiable op (K3 forward / backward).  The confidence is
    returned without a gradient, as in the reference (torch.no_grad(), models/MVSNet/model.py:211-215)."""

    @staticmethod
    def forward(ctx, score, depth, interval, conf_mode):
        score = score.contiguous()
        out = depth_regress(score, depth, interval, conf_mode=conf_mode)
        ctx.save_for_backward(score, depth, interval)
        conf = out["conf"] if out["conf"] is not None else out["depth"].new_zeros(())
        ctx.mark_non_differentiable(conf)
        return out["depth"], conf

    @staticmethod
    def backward(ctx, g_depth, _g_conf):
        score, depth, interval = ctx.saved_tensors
        if ctx.needs_input_grad[1]:    # per-voxel hypotheses that depend on an earlier depth map (CVP refinement levels)
            if depth.dim() != 4 or interval is not None:
                raise L.Mvsb200Error("regress_depth: only per-voxel hypotheses [B,D,H,W] can receive a gradient")
            g, gh = depth_regress_backward(g_depth, score, depth, None, want_grad_hyp=True)
            return g, gh, None, None
        return depth_regress_backward(g_depth, score, depth, interval), None, None, None


def regress_depth(score, depth, interval=None, conf_mode=L.CONF_NONE):
    """Differentiable depth regression (training, row f2): -> (depth [B,H,W], confidence [B,H,W] or a 0-d zero)."""
    return _DepthRegressFn.apply(score, depth, interval, conf_mode)


@_on_operand_device
def vis_fuse(interms, uncerts):
    """interms list of [B,D,H,W,G]; uncerts list of [B,H,W] -> [B,D,H,W,G]."""
    lib = L.load()
    S = len(interms)
    B, D, H, W, G = interms[0].shape
    ip, up = (ctypes.c_void_p * S)(), (ctypes.c_void_p * S)()
    for i in range(S):
        _dev_f32(interms[i], "interm[%d]" % i)
        _dev_f32(uncerts[i], "uncert[%d]" % i)
        assert interms[i].shape == interms[0].shape and uncerts[i].numel() == B * H * W
        ip[i], up[i] = interms[i].data_ptr(), uncerts[i].data_ptr()
    out = torch.empty_like(interms[0])
    L.check(lib.mvsb200_vis_fuse(ip, up, S, B, D, H, W, G, _ptr(out), _stream()), "mvsb200_vis_fuse")
    return out


def uncert_net_params(conv1, conv2, head):
    """Host parameter block of mvsb200_vis_uncert_net from three PackedConv (2-D, BN folded): a ctypes float array."""
    parts = [conv1.w.reshape(-1), conv1.scale, conv1.bias, conv2.w.reshape(-1), conv2.scale, conv2.bias, head.w.reshape(-1)]
    flat = torch.cat([t.detach().float().reshape(-1).cpu() for t in parts])
    if flat.numel() != 752:
        raise L.Mvsb200Error("uncert_net_params: expected 752 parameters, got %d" % flat.numel())
    return (ctypes.c_float * 752)(*flat.tolist())


@_on_operand_device
def vis_uncert_net(entropy, params):
    """entropy [N,H,W] -> log-uncertainty [N,H,W] (UncertNet, VisMVSNet/model_cas.py:77-98, one fused kernel)."""
    entropy = _dev_f32(entropy.contiguous(), "entropy")
    N, H, W = entropy.shape
    out = torch.empty_like(entropy)
    L.check(L.load().mvsb200_vis_uncert_net(_ptr(entropy), N, H, W, params, _ptr(out), _stream()), "mvsb200_vis_uncert_net")
    return out


# ------------------------------------------------------------------------------------------------
# K7
# ------------------------------------------------------------------------------------------------
class PackedConv2d:
    """A 2-D conv (+ eval-mode BatchNorm) packed for mvsb200_conv2d: weights [k*k][Cin_padded][Cout], folded scale/bias."""

    def __init__(self, weight, bn=None, conv_bias=None, stride=1, relu=False):
        w = weight.detach().float()                      # [Cout,Cin,k,k]
        self.cout, cin, self.k, _ = w.shape
        self.cin = (cin + 3) // 4 * 4                    # a 3-channel image is consumed zero-padded to 4 channels
        if self.cin != cin:
            w = torch.cat([w, w.new_zeros(self.cout, self.cin - cin, self.k, self.k)], 1)
        self.w = w.permute(2, 3, 1, 0).contiguous()
        self.stride, self.relu = stride, int(relu)
        if bn is not None:
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
            if conv_bias is not None:
                bias = bias + conv_bias.detach().float() * scale
            self.scale, self.bias = scale.contiguous(), bias.contiguous()
        else:
            self.scale = None
            self.bias = conv_bias.detach().float().contiguous() if conv_bias is not None else None


@_on_operand_device
def conv2d(x, layer):
    """x [B,H,W,Cin] channels-last -> [B,Ho,Wo,Cout] (K7)."""
    x = _dev_f32(x, "x")
    B, H, W, C = x.shape
    if C != layer.cin:
        raise L.Mvsb200Error("conv2d: input has %d channels, layer expects %d" % (C, layer.cin))
    pad = layer.k // 2
    ho, wo = (H + 2 * pad - layer.k) // layer.stride + 1, (W + 2 * pad - layer.k) // layer.stride + 1
    y = torch.empty(B, ho, wo, layer.cout, device=x.device, dtype=torch.float32)
    L.check(L.load().mvsb200_conv2d(B, H, W, C, layer.cout, layer.k, layer.stride, layer.relu, _ptr(x), _ptr(layer.w),
                                    _ptr(layer.scale), _ptr(layer.bias), _ptr(y), _stream()), "mvsb200_conv2d")
    return y


@_on_operand_device
def bias_act_(y, bias, scale=None, residual=None, slope=0.0, nhwc=False):
    """In place y = act(y * scale[c] + bias[c] (+ residual)) over a channels-last map (K7b, mvsb200_bias_act): the fused
    epilogue behind a cuDNN convolution call that has none.  `y` is [N,C,H,W] in torch.channels_last memory (what cuDNN
    returns for channels-last inputs) or, with nhwc=True, [N,H,W,C] contiguous; C % 4 == 0; slope 0 = ReLU, 0.1 = CVP's
    LeakyReLU, 1 = none.  Returns y."""
    if y.dim() != 4:
        raise L.Mvsb200Error("bias_act_: y must be 4-D")
    if nhwc and y.is_contiguous():
        C, npix = y.shape[3], y.shape[0] * y.shape[1] * y.shape[2]
    elif not nhwc and y.is_contiguous(memory_format=torch.channels_last):
        C, npix = y.shape[1], y.shape[0] * y.shape[2] * y.shape[3]
    else:
        raise L.Mvsb200Error("bias_act_: y must be dense channels-last memory")
    if not (y.is_cuda and y.dtype == torch.float32) or bias.numel() != C:
        raise L.Mvsb200Error("bias_act_: y must be CUDA float32 with %d == len(bias) channels" % C)
    if residual is not None and (residual.shape != y.shape or residual.stride() != y.stride() or residual.dtype != torch.float32):
        raise L.Mvsb200Error("bias_act_: residual must have y's shape and layout")
    bias = _dev_f32(bias.contiguous(), "bias")
    if scale is not None:
        scale = _dev_f32(scale.contiguous(), "scale")
    L.check(L.load().mvsb200_bias_act(_ptr(y), npix, C, _ptr(scale), _ptr(bias), _ptr(residual), ctypes.c_float(slope), None, _stream()),
            "mvsb200_bias_act")
    return y


# ------------------------------------------------------------------------------------------------
# K5
# ------------------------------------------------------------------------------------------------
@_on_operand_device
def cvp_depth_delta(ref_depth, ref_in, src_in, ref_ex, src_ex):
    """ref_depth [B,H,W]; ref_in, src_in [B,3,3]; ref_ex, src_ex [B,4,4] -> |delta| [B,H*W] fp64, +inf where invalid
    (the per-pixel solve of calDepthHypo, CVP_MVSNet/models/modules.py:131-214)."""
    ref_depth = _dev_f32(ref_depth.contiguous(), "ref_depth")
    B, H, W = ref_depth.shape
    mats = [_dev_f32(m.contiguous(), n) for m, n in ((ref_in, "ref_in"), (src_in, "src_in"), (ref_ex, "ref_ex"), (src_ex, "src_ex"))]
    out = torch.empty(B, H * W, device=ref_depth.device, dtype=torch.float64)
    L.check(L.load().mvsb200_cvp_depth_delta(_ptr(ref_depth), *[_ptr(m) for m in mats], B, H, W, _ptr(out), _stream()),
            "mvsb200_cvp_depth_delta")
    return out
