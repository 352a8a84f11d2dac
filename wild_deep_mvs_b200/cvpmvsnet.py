"""Drop-in CVP-MVSNet (`Frontend`) running its coarse-to-fine cost-volume pyramid on libmvsb200.

Mirrors models/CVP_MVSNet/frontend.py:5-38 and models/CVP_MVSNet/models/net.py:86-229 of the reference: same class
layout (`Frontend.model.{featurePyramid, cost_reg_refine}`), state_dict key names, `nscale` attribute / kwarg,
forward signature and returned dict.

Per pyramid level (coarsest first):
  K1  fused warp + variance ((M1/V)^2 rounding order, net.py:152, modules.py:289) over 16-channel features, with
      scalar hypotheses [B,D] at the coarsest level and per-pixel hypotheses [B,8,H,W] at the refine levels
      (homo_warping / proj_cost, modules.py:74-128,229-293)
  K2  the shared CostRegNet (net.py:50-85) on the tcgen05 tensor cores, BN/ReLU/skip fused
  K3  softmax + (per-pixel) depth regression, and the 4-bin confidence on the last level (net.py:161-162,203-219)
  K5  the fp64 per-pixel epipolar solve of `calDepthHypo` (modules.py:131-226; mvsb200_cvp_depth_delta) -- only the median of
      its result (a device sort) is a PyTorch op
FeaturePyramid (2-D CNN, all same-sized views as one batch; row f1) runs its convolutions on cuDNN and their bias +
LeakyReLU epilogue as one in-place library pass (K7b; MVSB200_CVP_PYRAMID=torch selects the plain modules); the bicubic
x2 depth up-sampling (net.py:169-170) stays in PyTorch on the device.
Eval mode runs the kernels above; training mode (row f2) uses the K1 / K3 forward + backward kernels behind autograd
Functions with the regulariser as PyTorch modules (network._forward_train; tests/test_gpu_backward.py).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops


def _conv2d(cin, cout):
    return nn.Sequential(nn.Conv2d(cin, cout, 3, stride=1, padding=1, dilation=1, bias=True), nn.LeakyReLU(0.1))


class FeaturePyramid(nn.Module):
    """Same parameters as models/CVP_MVSNet/models/net.py:21-47."""

    _ORDER = ("conv0aa", "conv0ba", "conv0bb", "conv0bc", "conv0bd", "conv0be", "conv0bf", "conv0bg", "conv0bh")

    def __init__(self):
        super().__init__()
        chans = [(3, 64), (64, 64), (64, 64), (64, 32), (32, 32), (32, 32), (32, 16), (16, 16), (16, 16)]
        for name, (a, b) in zip(self._ORDER, chans):
            setattr(self, name, _conv2d(a, b))

    def _features(self, img):
        f = img.contiguous(memory_format=torch.channels_last)
        fused = f.is_cuda and not torch.is_grad_enabled() and os.environ.get("MVSB200_CVP_PYRAMID", "fused") == "fused"
        for name in self._ORDER:
            m = getattr(self, name)
            if fused:
                # cuDNN convolution without its bias, then ONE in-place pass for bias + LeakyReLU (K7b) instead of the
                # broadcast-add and the activation kernel nn.Conv2d(bias) + nn.LeakyReLU run (each a full read + write
                # of the 64-channel full-resolution maps: 15 of the pyramid's 24.5 ms at cfg4)
                f = F.conv2d(f, m[0].weight, None, 1, 1)
                if not f.is_contiguous(memory_format=torch.channels_last):
                    f = f.contiguous(memory_format=torch.channels_last)
                ops.bias_act_(f, m[0].bias, slope=m[1].negative_slope)
            else:
                f = m(f)
        return f

    def forward(self, img, scales=5):
        fp = [self._features(img)]
        for _ in range(scales - 1):
            img = F.interpolate(img, scale_factor=0.5, mode="bilinear", align_corners=None).detach()
            fp.append(self._features(img))
        return fp


def _cbr3d(cin, cout, stride=1):
    m = nn.Module()
    m.conv = nn.Conv3d(cin, cout, 3, stride=stride, padding=1, bias=False)
    m.bn = nn.BatchNorm3d(cout)
    return m


class CostRegNet(nn.Module):
    """Parameters named as models/CVP_MVSNet/models/net.py:50-74, executed by K2."""

    def __init__(self):
        super().__init__()
        self.conv0, self.conv0a = _cbr3d(16, 16), _cbr3d(16, 16)
        self.conv1 = _cbr3d(16, 32, stride=2)
        self.conv2, self.conv2a = _cbr3d(32, 32), _cbr3d(32, 32)
        self.conv3 = _cbr3d(32, 64)
        self.conv4, self.conv4a = _cbr3d(64, 64), _cbr3d(64, 64)
        self.conv5 = nn.Sequential(nn.ConvTranspose3d(64, 32, 3, padding=1, output_padding=0, stride=1, bias=False),
                                   nn.BatchNorm3d(32), nn.ReLU(inplace=True))
        self.conv6 = nn.Sequential(nn.ConvTranspose3d(32, 16, 3, padding=1, output_padding=1, stride=2, bias=False),
                                   nn.BatchNorm3d(16), nn.ReLU(inplace=True))
        self.prob0 = nn.Conv3d(16, 1, 3, stride=1, padding=1)
        self._packed = None
        self._packed_key = None

    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(
            (b.data_ptr(), b._version) for b in self.buffers())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        pk = {}
        for name in ("conv0", "conv0a", "conv1", "conv2", "conv2a", "conv3", "conv4", "conv4a"):
            m = getattr(self, name)
            pk[name] = ops.PackedConv(m.conv.weight, m.bn, stride=m.conv.stride[0], relu=True)
        # conv2 + relu(bn(deconv(conv4))) and conv0 + relu(bn(deconv(conv5)))  (net.py:81-82)
        pk["conv5"] = ops.PackedConv(self.conv5[0].weight, self.conv5[1], stride=1, transposed=True, relu=True,
                                     skip_mode=L.SKIP_AFTER_RELU)
        pk["conv6"] = ops.PackedConv(self.conv6[0].weight, self.conv6[1], stride=2, transposed=True, relu=True,
                                     skip_mode=L.SKIP_AFTER_RELU)
        pk["prob0"] = ops.PackedConv(self.prob0.weight, None, conv_bias=self.prob0.bias)
        self._packed, self._packed_key = pk, key
        return pk

    def run(self, vol):
        """vol [B,D,H,W,16] channels-last -> score [B,D,H,W]."""
        if self.training:
            raise NotImplementedError("libmvsb200 runs the regulariser in eval mode only (SURVEY.md 8-f2)")
        B, D, H, W, _ = vol.shape
        if D % 2 or H % 2 or W % 2:
            raise L.Mvsb200Error("CVP CostRegNet needs even D,H,W (got %d,%d,%d)" % (D, H, W))
        pk = self._pack()
        c0 = ops.conv3d(ops.conv3d(vol, pk["conv0"]), pk["conv0a"])
        c2 = ops.conv3d(ops.conv3d(ops.conv3d(c0, pk["conv1"]), pk["conv2"]), pk["conv2a"])
        c4 = ops.conv3d(ops.conv3d(ops.conv3d(c2, pk["conv3"]), pk["conv4"]), pk["conv4a"])
        c5 = ops.conv3d(c4, pk["conv5"], skip=c2)
        del c4
        c6 = ops.conv3d(c5, pk["conv6"], skip=c0)
        del c5, c0
        return ops.conv3d(c6, pk["prob0"]).squeeze(-1)

    def forward_modules(self, x):
        """The same network through the PyTorch modules that own the parameters (net.py:76-85): training mode, where
        BatchNorm uses batch statistics and autograd needs the layer graph.  Convolutions on cuDNN by default; with
        MVSB200_TRAIN_K2=lib on the library forward and backward (ops.conv3d_train), BatchNorm / ReLU stay PyTorch ops."""
        lib = os.environ.get("MVSB200_TRAIN_K2") == "lib" and x.is_cuda

        def conv(v, m, stride=1, transposed=False):
            if lib:
                return ops.as_ncdhw(ops.conv3d_train(ops.to_ndhwc(v), m.weight, m.bias, stride, transposed))
            return m(v)

        cbr = lambda m, v: F.relu(m.bn(conv(v, m.conv, m.conv.stride[0])), inplace=True)
        dbr = lambda m, v, stride: m[2](m[1](conv(v, m[0], stride, True)))
        conv0 = cbr(self.conv0a, cbr(self.conv0, x))
        conv2 = cbr(self.conv2a, cbr(self.conv2, cbr(self.conv1, conv0)))
        conv4 = cbr(self.conv4a, cbr(self.conv4, cbr(self.conv3, conv2)))
        conv5 = conv2 + dbr(self.conv5, conv4, 1)
        conv6 = conv0 + dbr(self.conv6, conv5, 2)
        return conv(conv6, self.prob0).squeeze(1)

    def forward(self, x):
        """Reference signature: x [B,16,D,H,W] -> [B,D,H,W] (net.py:76-85)."""
        if self.training or (torch.is_grad_enabled() and x.requires_grad):
            return self.forward_modules(x)
        return self.run(ops.to_ndhwc(x))


def condition_intrinsics(intrinsics, img_shape, fp_shapes):
    """Per-level K[:2] /= (image height / feature height); models/CVP_MVSNet/models/modules.py:31-46 -> [B,L,3,3]."""
    out = []
    for s in fp_shapes:
        k = intrinsics.clone()
        k[:, :2, :] = k[:, :2, :] / (img_shape[2] / s[2])
        out.append(k)
    return torch.stack(out).permute(1, 0, 2, 3)


def sweeping_depth_hypos(depth_min, depth_max, n):
    """min + i * (max - min) / n  (divisor n, not n-1); modules.py:53-71 -> [B,n]."""
    assert n % 2 == 0
    interval = (depth_max - depth_min) / n
    return depth_min.unsqueeze(1) + torch.arange(n, device=depth_max.device) * interval.unsqueeze(1)


def cal_depth_hypo(ref_depths, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, d=4):
    """Per-level hypotheses depth_up + k * interval, k in [-d, d), with the interval = median over pixels of the depth
    change that moves the projection into the FIRST source view by one pixel along the epipolar line (fp64).
    calDepthHypo, models/CVP_MVSNet/models/modules.py:131-226: the per-pixel solve is K5 (mvsb200_cvp_depth_delta), the
    median a device sort -- no host synchronisation.  ref_depths [B,H,W]; ref_in [B,3,3]; src_in [B,S,3,3];
    ref_ex [B,4,4]; src_ex [B,S,4,4] -> [B,2d,H,W] fp32."""
    B, H, W = ref_depths.shape
    delta = ops.cvp_depth_delta(ref_depths, ref_in, src_in[:, 0], ref_ex, src_ex[:, 0])          # [B,HW] fp64, +inf = invalid
    nvalid = torch.isfinite(delta).sum(1)
    srt = torch.sort(delta, dim=1).values
    med = srt.gather(1, ((nvalid - 1).clamp(min=0) // 2).unsqueeze(1)).squeeze(1)                 # torch.median: the lower middle value
    fallback = ((depth_max - depth_min) / 128).double()                                           # degenerate geometry (modules.py:211-213)
    interval = torch.where(nvalid > 0, med, fallback)
    steps = torch.arange(-d, d, device=ref_depths.device, dtype=torch.float64).view(1, -1, 1, 1)
    hyp = ref_depths.unsqueeze(1).repeat(1, 2 * d, 1, 1)
    hyp += steps * interval.view(B, 1, 1, 1)   # in-place add of fp64 into the fp32 map (:219)
    return hyp


class network(nn.Module):
    def __init__(self):
        super().__init__()
        self.featurePyramid = FeaturePyramid()
        self.cost_reg_refine = CostRegNet()
        self.nscale = 2

    @staticmethod
    def _projs(K, E):
        """[K E[:3]; 0 0 0 1]  (modules.py:89-93)."""
        P = torch.zeros(K.shape[:-2] + (4, 4), device=K.device, dtype=K.dtype)
        P[..., :3, :] = K @ E[..., :3, :]
        P[..., 3, 3] = 1
        return P

    def depth_from_pyramids(self, ref_fp, src_fps, ref_in_ms, src_in_ms, ref_ex, src_ex, depth_min, depth_max):
        """The hot path proper.  ref_fp[l] / src_fps[v][l]: channels-last [B,h_l,w_l,16], level 0 = finest;
        ref_in_ms [B,L,3,3], src_in_ms [B,S,L,3,3] level-conditioned intrinsics.  Returns (depth_est_list coarse->fine,
        confidence of the finest level, per-level seams)."""
        nscale = len(ref_fp)
        S = len(src_fps)
        reg = self.cost_reg_refine
        seams = {}

        def cost_volume(level, hypos):
            rp = self._projs(ref_in_ms[:, level], ref_ex)
            sp = self._projs(src_in_ms[:, :, level], src_ex)
            warp = ops.mvs_relative_proj(rp, sp)
            return ops.build_cost_volume(ref_fp[level], [src_fps[v][level] for v in range(S)], warp, hypos,
                                         hypos.shape[1], L.GEOM_MVS, L.AGG_VARIANCE_MEAN)

        hypos = sweeping_depth_hypos(depth_min, depth_max, 48 if self.training else 96).float().contiguous()
        last = nscale == 1
        score = reg.run(cost_volume(nscale - 1, hypos))
        out = ops.depth_regress(score, hypos, conf_mode=L.CONF_SUM4 if last else L.CONF_NONE)
        seams["reg_out_coarse"] = score
        depth = out["depth"]
        ests = [depth]
        nan_flags = []
        for level in range(nscale - 2, -1, -1):
            depth_up = F.interpolate(depth[None, :], size=None, scale_factor=2, mode="bicubic", align_corners=None).squeeze(0)
            hypos = cal_depth_hypo(depth_up, ref_in_ms[:, level], src_in_ms[:, :, level], ref_ex, src_ex, depth_min,
                                   depth_max).contiguous()
            nan_flags.append(torch.isnan(hypos).any())   # checked once, after the last level (below)
            score = reg.run(cost_volume(level, hypos))
            out = ops.depth_regress(score, hypos, conf_mode=L.CONF_SUM4 if level == 0 else L.CONF_NONE)
            seams["hypos_l%d" % level] = hypos
            seams["reg_out_l%d" % level] = score
            depth = out["depth"]
            ests.append(depth)
        # The reference prints "NAN" per level (net.py:188-189,196-197) -- a host read of a device flag, i.e. a pipeline
        # drain per level.  Here the flags stay on the device and are read ONCE at the end; not at all while the forward is
        # being captured into a CUDA graph (graphed_forward), where a host read is illegal.
        if nan_flags and not torch.cuda.is_current_stream_capturing():
            if torch.stack(nan_flags).any():
                print("NAN")
        return ests, out["conf"], seams

    def _forward_train(self, ref_img, src_imgs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, nscale):
        """Training path (net.py:96-229 with self.training): 48 initial hypotheses, refinement hypotheses at fixed
        intervals halved per level (net.py:176-182) instead of calDepthHypo.  The cost volumes and the regression run on
        K1 / K3 forward + backward (autograd Functions), the pyramid and the regulariser as PyTorch modules.  Gradients:
        features and regulariser at every level, and through each level's hypotheses (built from the up-sampled depth
        of the level before) into the coarser levels -- the sampling grids carry none (modules.py:88,242)."""
        pyrs = [self.featurePyramid(im, nscale) for im in [ref_img] + list(src_imgs)]
        ref_pyr, src_pyrs = pyrs[0], pyrs[1:]
        S = len(src_pyrs)
        ref_in_ms = condition_intrinsics(ref_in, ref_img.shape, [f.shape for f in ref_pyr])
        src_in_ms = torch.stack([condition_intrinsics(src_in[:, i], ref_img.shape, [f.shape for f in src_pyrs[i]])
                                 for i in range(S)]).permute(1, 0, 2, 3, 4)

        def level_depth(level, hypos, last):
            warp = ops.mvs_relative_proj(self._projs(ref_in_ms[:, level], ref_ex), self._projs(src_in_ms[:, :, level], src_ex))
            vol = ops.cost_volume(ops.to_nhwc(ref_pyr[level]), [ops.to_nhwc(src_pyrs[v][level]) for v in range(S)], warp,
                                  hypos.detach().contiguous(), hypos.shape[1], L.GEOM_MVS, L.AGG_VARIANCE_MEAN)
            score = self.cost_reg_refine(ops.as_ncdhw(vol))
            return ops.regress_depth(score, hypos.contiguous(), None, L.CONF_SUM4 if last else L.CONF_NONE)

        hypos = sweeping_depth_hypos(depth_min, depth_max, 48).float()
        depth, conf = level_depth(nscale - 1, hypos, nscale == 1)
        ests = [depth]
        for id_level, level in enumerate(range(nscale - 2, -1, -1)):
            depth_up = F.interpolate(depth[None, :], size=None, scale_factor=2, mode="bicubic", align_corners=None).squeeze(0)
            interval = (depth_max - depth_min) / 48 / 2 ** (id_level + 1)
            hypos = torch.stack([depth_up + i * interval.view(-1, 1, 1) for i in range(-4, 4)], dim=1)
            depth, conf = level_depth(level, hypos, level == 0)
            ests.append(depth)
        ests.reverse()
        return {"depth_est_list": ests, "prob_confidence": conf}

    def forward(self, ref_img, src_imgs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, **kwargs):
        nscale = kwargs["nscale"] if "nscale" in kwargs else self.nscale
        if self.training:
            return self._forward_train(ref_img, src_imgs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, nscale)
        with torch.no_grad():
            pyrs = ops.map_views(lambda im: self.featurePyramid(im, nscale), [ref_img] + list(src_imgs))
            ref_pyr, src_pyrs = pyrs[0], pyrs[1:]
            ref_in_ms = condition_intrinsics(ref_in, ref_img.shape, [f.shape for f in ref_pyr])
            src_in_ms = torch.stack([condition_intrinsics(src_in[:, i], ref_img.shape, [f.shape for f in src_pyrs[i]])
                                     for i in range(len(src_imgs))]).permute(1, 0, 2, 3, 4)
            ests, conf, _ = self.depth_from_pyramids([ops.to_nhwc(f) for f in ref_pyr],
                                                     [[ops.to_nhwc(f) for f in p] for p in src_pyrs],
                                                     ref_in_ms, src_in_ms, ref_ex, src_ex, depth_min, depth_max)
        ests.reverse()  # depth_est_list[0] is the largest scale (net.py:223)
        return {"depth_est_list": ests, "prob_confidence": conf}


class Frontend(nn.Module):
    def __init__(self):
        super().__init__()
        self.model = network()

    def graphed_forward(self, imgs, K, R, t, depth_min, depth_max, reference_frame=0, **kwargs):
        """The whole eval-mode `forward` (images -> dict, same kwargs) captured into ONE CUDA graph for inputs of these shapes
        (mvsnet.GraphedForward): a callable taking the same six tensors (None = keep the captured values) and returning the
        captured output dict (overwritten by the next replay).  The eager forward of this model is launch bound."""
        from .mvsnet import GraphedForward
        return GraphedForward(self, imgs, K, R, t, depth_min, depth_max, reference_frame, **kwargs)

    def forward(self, imgs, K, R, t, depth_min, depth_max, reference_frame=0, **kwargs):
        src_idx = list(range(reference_frame)) + list(range(reference_frame + 1, K.shape[1]))
        if isinstance(imgs, torch.Tensor):
            ref_img = imgs[:, reference_frame]
            src_imgs = [imgs[:, i] for i in src_idx]
        else:
            ref_img = imgs[reference_frame]
            src_imgs = imgs[:reference_frame] + imgs[reference_frame + 1:]
        b, n = ref_img.shape[0], len(src_imgs)
        # (0, 0, 0, 1) built on the device (fill kernels): no host -> device copy, so the forward can be captured into a CUDA graph
        last = torch.cat((torch.zeros(3, device=K.device, dtype=K.dtype), torch.ones(1, device=K.device, dtype=K.dtype)))
        # (views picked with Python indices, not `x[:, src_idx]`: a list index becomes a host tensor that is copied to the device)
        pick = lambda x: torch.stack([x[:, i] for i in src_idx], 1)
        ref_ex = torch.cat((torch.cat((R[:, reference_frame], t[:, reference_frame]), dim=2),
                            last.view(1, 1, 4).expand(b, 1, 4)), dim=1)
        src_ex = torch.cat((torch.cat((pick(R), pick(t)), dim=3),
                            last.view(1, 1, 1, 4).expand(b, n, 1, 4)), dim=2)
        output = self.model(ref_img, src_imgs, K[:, reference_frame], pick(K), ref_ex, src_ex,
                            depth_min[:, reference_frame], depth_max[:, reference_frame], **kwargs)
        return {"depth": output["depth_est_list"][0].squeeze(1), "depth_est_list": output["depth_est_list"],
                "depth_pair_list": [], "photometric_confidence": output["prob_confidence"].unsqueeze(1)}
