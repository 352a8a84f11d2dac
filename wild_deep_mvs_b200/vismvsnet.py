"""Drop-in Vis-MVSNet (`Frontend`) running its three cascade stages on libmvsb200.

Mirrors models/VisMVSNet/frontend.py:6-109 of the reference: same class layout (`Frontend.model.feat_ext`,
`.stage1/2/3.{reg, reg_fuse, reg_pair, uncert_net}`), state_dict key names, `depth_nums` / `interval_scales`
attributes and kwargs, forward signature and returned dict.

Per stage (SingleStage.forward, models/VisMVSNet/model_cas.py:303-420, mode='soft'):
  K1  closed-form homographies + warp + 8-group correlation for ALL source views in one launch
  K2  Reg U-Net + RegPair head on the S pair volumes stacked along the batch axis (shared weights)
  K3  pair soft-argmin + entropy;  UncertNet (three tiny 2-D convs, also K2);
  K4  visibility-weighted fusion;  K2 RegFuse U-Net + head;  K3 soft-argmin with the +-2 window confidence.
FeatExt (2-D U-Net) and the bilinear up-sampling of the previous stage's depth stay in PyTorch ("next" rows; FeatExt has
an opt-in library path, MVSB200_VIS_FEATEXT=lib).
Eval mode runs the kernels above; training mode (`net.train()`, row f2) builds the per-pair group-correlation volumes with
the K1 forward / backward kernels (ops.cost_volume) and runs everything downstream as the reference's own PyTorch modules,
so every output carries the reference's gradient (SingleStage._run_train; tests/test_gpu_backward.py).
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops


class _Named(nn.Module):
    """Container whose children carry the reference's (non-identifier) names, e.g. 'reg14_0'."""

    def __init__(self, items):
        super().__init__()
        for name, m in items:
            self.add_module(str(name), m)

    def __getitem__(self, name):
        return self._modules[str(name)]

    def __iter__(self):
        return iter(self._modules.values())


def _conv_train(m, x, transposed=False):
    """A conv module of the PyTorch execution path.  3x3x3 layers of the regularisers go to the library forward AND
    backward (ops.conv3d_train: K2 engines + wgrad kernel) when MVSB200_TRAIN_K2=lib; everything else (2-D layers, the
    1x1x1 shortcuts) is the module itself."""
    if (os.environ.get("MVSB200_TRAIN_K2") == "lib" and x.is_cuda and x.dim() == 5 and tuple(m.kernel_size) == (3, 3, 3)):
        return ops.as_ncdhw(ops.conv3d_train(ops.to_ndhwc(x), m.weight, m.bias, m.stride[0], transposed))
    return m(x)


class _Block(nn.Module):
    """Parameters of BasicBlock (models/VisMVSNet/nn_utils.py:123-171)."""

    def __init__(self, cin, cout, stride, dim):
        super().__init__()
        conv = nn.Conv2d if dim == 2 else nn.Conv3d
        bn = nn.BatchNorm2d if dim == 2 else nn.BatchNorm3d
        self.conv1 = conv(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = bn(cout)
        self.conv2 = conv(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = bn(cout)
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(conv(cin, cout, 1, stride, bias=False), bn(cout))
        self.stride = stride

    def forward(self, x):  # PyTorch execution: the 2-D feature extractor, and the 3-D regularisers in training mode
        y = F.relu(self.bn1(_conv_train(self.conv1, x)), inplace=True)
        y = self.bn2(_conv_train(self.conv2, y))
        r = x if self.downsample is None else self.downsample(x)
        return F.relu(y + r, inplace=True)


def _layer(cin, cout, blocks, stride, dim):
    mods = [_Block(cin, cout, stride, dim)] + [_Block(cout, cout, 1, dim) for _ in range(blocks - 1)]
    return nn.Sequential(*mods)


class _FeatUNet(nn.Module):
    """UNet(16, 2, 1, 2, [], [32, 64, 128], [], '2d', 2) of models/VisMVSNet/model_cas.py:27."""

    def __init__(self):
        super().__init__()
        self.enc_blocks = _Named([("2d2_0", _layer(16, 32, 2, 1, 2)), ("2d4_1", _layer(32, 64, 2, 2, 2)),
                                  ("2d8_2", _layer(64, 128, 2, 2, 2))])
        self.dec_blocks = _Named([
            ("2d16_3", _Named([(0, nn.ConvTranspose2d(128, 64, 3, 2, 1, 1, bias=False)), (1, nn.Conv2d(128, 64, 3, 1, 1, bias=False)),
                               (2, _layer(64, 64, 1, 1, 2))])),
            ("2d8_4", _Named([(0, nn.ConvTranspose2d(64, 32, 3, 2, 1, 1, bias=False)), (1, nn.Conv2d(64, 32, 3, 1, 1, bias=False)),
                              (2, _layer(32, 32, 1, 1, 2))]))])

    def forward(self, x):
        enc = []
        for b in self.enc_blocks:
            x = b(x)
            enc.append(x)
        outs = [x]
        for i, b in enumerate(self.dec_blocks):
            x = b[0](x)
            x = b[1](torch.cat([x, enc[-2 - i]], 1))
            x = b[2](x)
            outs.append(x)
        return outs


def _w3d(w2d):
    """A 3x3 2-D kernel as the middle depth tap of a 3x3x3 kernel: on a one-plane volume ([N,1,H,W,C]) the 3-D conv
    engines (K2) then compute exactly the 2-D convolution -- the planes above and below are padding, and the z-march
    engine does not even stage them."""
    w3 = w2d.new_zeros(w2d.shape[0], w2d.shape[1], 3, w2d.shape[2], w2d.shape[3])
    w3[:, :, 1] = w2d
    return w3


def _pack_block2d(b):
    """BasicBlock (nn_utils.py:123-171) of the 2-D U-Net packed for K2 on one-plane volumes."""
    pk = {"c1": ops.PackedConv(_w3d(b.conv1.weight), b.bn1, stride=b.stride, relu=True),
          "c2": ops.PackedConv(_w3d(b.conv2.weight), b.bn2, relu=True, skip_mode=L.SKIP_BEFORE_RELU)}
    if b.downsample is not None:
        pk["ds"] = ops.PackedConv(b.downsample[0].weight, b.downsample[1], stride=b.stride)   # 1x1: pointwise kernel
    return pk


def _run_block2d(pk, x):
    r = ops.conv3d(x, pk["ds"]) if "ds" in pk else x
    return ops.conv3d(ops.conv3d(x, pk["c1"]), pk["c2"], skip=r)


class FeatExt(nn.Module):
    """models/VisMVSNet/model_cas.py:18-35: 32-channel features at 1/8, 1/4, 1/2 image resolution."""

    def __init__(self):
        super().__init__()
        self.init_conv = nn.Sequential(nn.Conv2d(3, 16, 5, 2, 2, bias=False), nn.BatchNorm2d(16), nn.ReLU())
        self.unet = _FeatUNet()
        self.final_conv_1 = nn.Conv2d(128, 32, 3, 1, 1, bias=False)
        self.final_conv_2 = nn.Conv2d(64, 32, 3, 1, 1, bias=False)
        self.final_conv_3 = nn.Conv2d(32, 32, 3, 1, 1, bias=False)

    def _pack(self):
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, "_pack_key", None) != key:
            u = self.unet
            pk = {"init": ops.PackedConv2d(self.init_conv[0].weight, self.init_conv[1], stride=2, relu=True),
                  "enc": [[_pack_block2d(b) for b in layer] for layer in u.enc_blocks],
                  "dec": [{"up": ops.PackedConv(_w3d(d[0].weight), None, stride=2, transposed=True),
                           "post": ops.PackedConv(_w3d(d[1].weight), None),
                           "blocks": [_pack_block2d(b) for b in d[2]]} for d in u.dec_blocks],
                  "final": [ops.PackedConv(_w3d(c.weight), None) for c in (self.final_conv_1, self.final_conv_2, self.final_conv_3)]}
            self._packed, self._pack_key = pk, key
        return self._packed

    def run(self, x):
        """x [N,3,H,W] -> three channels-last feature maps [N,H/8,W/8,32], [N,H/4,W/4,32], [N,H/2,W/2,32], every layer on
        the library's kernels (row f1): the 5x5 stride-2 stem on K7, the U-Net on the K2 engines over one-plane volumes,
        the 1x1 shortcuts on the pointwise kernel."""
        pk = self._pack()
        N, C, H, W = x.shape
        x4 = torch.zeros(N, H, W, 4, device=x.device, dtype=torch.float32)
        x4[..., :C] = x.permute(0, 2, 3, 1)
        v = ops.conv2d(x4, pk["init"]).unsqueeze(1)                      # [N,1,H/2,W/2,16]
        enc = []
        for layer in pk["enc"]:
            for b in layer:
                v = _run_block2d(b, v)
            enc.append(v)
        outs = [v]
        for i, d in enumerate(pk["dec"]):
            up = ops.conv3d(v, d["up"])                                   # the 3-D transposed conv doubles the plane count:
            amax = getattr(up, "_mvs_amax", None)                        # (tracked by the z-march engine only)
            up = up[:, :1].contiguous()                                   # plane 0 is the 2-D result, plane 1 has no taps
            if amax is not None:
                ops.set_absmax(up, amax)
            v = ops.conv3d(up, d["post"], x2=enc[-2 - i])                 # conv(cat([up, enc], 1))
            for b in d["blocks"]:
                v = _run_block2d(b, v)
            outs.append(v)
        return tuple(ops.conv3d(o, f).squeeze(1) for o, f in zip(outs, pk["final"]))

    def _fused_params(self):
        """Eval-mode BatchNorm folded into the convolution that precedes it (w * scale, bias), channels-last, cached."""
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, "_fused_key", None) != key:
            def fold(conv, bn):
                s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                return ((conv.weight * s.view(-1, 1, 1, 1)).contiguous(memory_format=torch.channels_last),
                        (bn.bias - bn.running_mean * s).contiguous())

            def block(b):
                d = {"c1": fold(b.conv1, b.bn1), "c2": fold(b.conv2, b.bn2), "stride": b.stride, "ds": None}
                if b.downsample is not None:
                    d["ds"] = fold(b.downsample[0], b.downsample[1])
                return d
            with torch.no_grad():
                self._fused = {"init": fold(self.init_conv[0], self.init_conv[1]),
                               "enc": [[block(b) for b in layer] for layer in self.unet.enc_blocks],
                               "dec": [[block(b) for b in d[2]] for d in self.unet.dec_blocks]}
            self._fused_key = key
        return self._fused

    @staticmethod
    def _block_fused(b, x):
        """BasicBlock (nn_utils.py:123-171) as two cuDNN calls with fused epilogues: conv + bias + ReLU, conv + bias + skip + ReLU."""
        st = (b["stride"], b["stride"])
        y = torch.cudnn_convolution_relu(x, b["c1"][0], b["c1"][1], st, (1, 1), (1, 1), 1)
        r = x if b["ds"] is None else F.conv2d(x, b["ds"][0], b["ds"][1], st)
        return torch.cudnn_convolution_add_relu(y, b["c2"][0], r, 1.0, b["c2"][1], (1, 1), (1, 1), (1, 1), 1)

    def _forward_fused(self, x):
        P = self._fused_params()
        x = torch.cudnn_convolution_relu(x, P["init"][0], P["init"][1], (2, 2), (2, 2), (1, 1), 1)
        enc = []
        for layer in P["enc"]:
            for b in layer:
                x = self._block_fused(b, x)
            enc.append(x)
        outs = [x]
        for i, d in enumerate(self.unet.dec_blocks):
            x = d[0](x)
            x = d[1](torch.cat([x, enc[-2 - i]], 1))
            for b in P["dec"][i]:
                x = self._block_fused(b, x)
            outs.append(x)
        return self.final_conv_1(outs[0]), self.final_conv_2(outs[1]), self.final_conv_3(outs[2])

    def forward(self, x):
        # run() is parity-tested against the reference but, at 4.8 ms for five 640x512 views, slower than cuDNN: 2-D layers
        # with 64-128 channels waste two thirds of the tile engine's taps.  It is therefore opt-in (MVSB200_VIS_FEATEXT=lib)
        # until K7 grows a tensor-core path for wide layers.  Default in inference: the cuDNN convolutions with BatchNorm
        # folded and bias / ReLU / skip fused into the call (no separate elementwise passes); MVSB200_VIS_FEATEXT=torch
        # selects the plain modules (the reference's own execution, also the training path).
        mode = os.environ.get("MVSB200_VIS_FEATEXT", "fused")
        inference = not self.training and x.is_cuda and not torch.is_grad_enabled()
        if mode == "lib" and inference:
            return tuple(f.permute(0, 3, 1, 2) for f in self.run(x))      # reference layout as views of channels-last memory
        x = x.contiguous(memory_format=torch.channels_last)
        if mode == "fused" and inference:
            return self._forward_fused(x)
        o1, o2, o3 = self.unet(self.init_conv(x))
        return self.final_conv_1(o1), self.final_conv_2(o2), self.final_conv_3(o3)


class _RegUNet(nn.Module):
    """UNet(8, 1, 0, 4, [], [8, 16], [], tag, dim=3) (model_cas.py:43,63); executed by K2."""

    def __init__(self, tag):
        super().__init__()
        self.tag = tag
        self.enc_blocks = _Named([(tag + "4_0", _layer(8, 8, 1, 1, 3)), (tag + "8_1", _layer(8, 16, 1, 2, 3))])
        self.dec_blocks = _Named([(tag + "16_2", _Named([(0, nn.ConvTranspose3d(16, 8, 3, 2, 1, 1, bias=False)),
                                                         (1, nn.Conv3d(16, 8, 3, 1, 1, bias=False))]))])

    def pack(self):
        b0 = self.enc_blocks[self.tag + "4_0"][0]
        b1 = self.enc_blocks[self.tag + "8_1"][0]
        dec = self.dec_blocks[self.tag + "16_2"]
        return {
            "b0c1": ops.PackedConv(b0.conv1.weight, b0.bn1, relu=True),
            "b0c2": ops.PackedConv(b0.conv2.weight, b0.bn2, relu=True, skip_mode=L.SKIP_BEFORE_RELU),
            "b1ds": ops.PackedConv(b1.downsample[0].weight, b1.downsample[1], stride=2),
            "b1c1": ops.PackedConv(b1.conv1.weight, b1.bn1, stride=2, relu=True),
            "b1c2": ops.PackedConv(b1.conv2.weight, b1.bn2, relu=True, skip_mode=L.SKIP_BEFORE_RELU),
            "up": ops.PackedConv(dec[0].weight, None, stride=2, transposed=True),
            "post": ops.PackedConv(dec[1].weight, None),
        }

    def forward(self, x):
        """PyTorch execution (training: batch-statistics BatchNorm, autograd): UNet.forward, nn_utils.py:256-278."""
        e0 = self.enc_blocks[self.tag + "4_0"](x)
        e1 = self.enc_blocks[self.tag + "8_1"](e0)
        dec = self.dec_blocks[self.tag + "16_2"]
        return _conv_train(dec[1], torch.cat([_conv_train(dec[0], e1, transposed=True), e0], 1))

    @staticmethod
    def run(pk, x, am):
        """am: ops.AmaxPool -- abs-max scalars of the layer outputs (one fill for the whole stage)."""
        conv = lambda t, name, **kw: ops.conv3d(t, pk[name], amax=am.take(), **kw)
        e0 = conv(conv(x, "b0c1"), "b0c2", skip=x)
        r = conv(e0, "b1ds")
        e1 = conv(conv(e0, "b1c1"), "b1c2", skip=r)
        up = conv(e1, "up")
        return conv(up, "post", x2=e0)  # conv(cat([up, e0], 1)), nn_utils.py:268-269


class Reg(nn.Module):
    def __init__(self):
        super().__init__()
        self.unet = _RegUNet("reg1")


class RegPair(nn.Module):
    def __init__(self):
        super().__init__()
        self.final_conv = nn.Conv3d(8, 1, 3, 1, 1, bias=False)


class RegFuse(nn.Module):
    def __init__(self):
        super().__init__()
        self.unet = _RegUNet("reg2")
        self.final_conv = nn.Conv3d(8, 1, 3, 1, 1, bias=False)


class UncertNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Sequential(nn.Conv2d(1, 8, 3, 1, 1, bias=False), nn.BatchNorm2d(8), nn.ReLU())
        self.conv2 = nn.Sequential(nn.Conv2d(8, 8, 3, 1, 1, bias=False), nn.BatchNorm2d(8), nn.ReLU())
        self.head_convs = _Named([(0, nn.Conv2d(8, 1, 3, 1, 1, bias=False))])

    def forward(self, x):
        """PyTorch execution (training), model_cas.py:93-98."""
        out = self.conv2(self.conv1(x))
        out = out + x
        return [conv(out) for conv in self.head_convs]


class SingleStage(nn.Module):
    def __init__(self):
        super().__init__()
        self.reg = Reg()
        self.reg_fuse = RegFuse()
        self.reg_pair = RegPair()
        self.uncert_net = UncertNet()
        self._packed, self._key = None, None

    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(
            (b.data_ptr(), b._version) for b in self.buffers())
        if self._packed is None or key != self._key:
            u = self.uncert_net
            self._packed = {
                "reg": self.reg.unet.pack(), "fuse": self.reg_fuse.unet.pack(),
                "pair_head": ops.PackedConv(self.reg_pair.final_conv.weight),
                "fuse_head": ops.PackedConv(self.reg_fuse.final_conv.weight),
                "u1": ops.PackedConv(u.conv1[0].weight, u.conv1[1], relu=True),
                "u2": ops.PackedConv(u.conv2[0].weight, u.conv2[1], relu=True, skip_mode=L.SKIP_AFTER_RELU),
                "uh": ops.PackedConv(u.head_convs[0].weight),
            }
            pk = self._packed
            pk["uncert"] = ops.uncert_net_params(pk["u1"], pk["u2"], pk["uh"])
            self._key = key
        return self._packed

    def run(self, ref, srcs, ref_cam, src_cams, depth_num, depth_start, depth_interval, s_scale, out_depth=None):
        """ref [B,H,W,32], srcs list of [B,Hs,Ws,32]; cams [B,2,4,4] / [B,S,2,4,4]; depth_start [B] or [B,H,W];
        depth_interval [B].  Returns est_depth [B,H,W], prob_map [B,H,W], pair list [(depth, uncert)].
        `out_depth` [B,H,W]: caller-owned buffer for est_depth (a gather slice, shard.DepthGather)."""
        if self.training:
            return self._run_train(ref, srcs, ref_cam, src_cams, depth_num, depth_start, depth_interval, s_scale)
        pk = self._pack()
        B, H, W, _ = ref.shape
        S = len(srcs)
        D = depth_num
        if D % 2 or H % 2 or W % 2:
            raise L.Mvsb200Error("Vis-MVSNet regulariser needs even D,H,W (got %d,%d,%d)" % (D, H, W))
        am = ops.AmaxPool(ref.device, 24)
        warp = ops.vis_homography_params(ref_cam, src_cams, 1.0 / s_scale)
        cost = ops.build_cost_volume(ref, srcs, warp, depth_start, D, L.GEOM_VIS, L.AGG_GROUPCORR,
                                     interval=depth_interval, groups=8, amax=am.take())   # [S,B,D,H,W,8]
        cost_sb = cost.view(S * B, D, H, W, 8)                                      # pairs stacked on the batch axis
        ops.set_absmax(cost_sb, cost._mvs_amax)                                       # abs-max tracked by K1 (views drop attributes)
        interm = _RegUNet.run(pk["reg"], cost_sb, am)
        del cost, cost_sb
        score = ops.conv3d(interm, pk["pair_head"]).squeeze(-1)                     # [S*B,D,H,W]
        start_sb = depth_start.repeat(S, *([1] * (depth_start.dim() - 1)))
        pair = ops.depth_regress(score, start_sb, interval=depth_interval.reshape(-1).repeat(S), want_entropy=True)
        u = ops.vis_uncert_net(pair["entropy"].view(S * B, H, W), pk["uncert"]).view(S, B, H, W)   # K6, model_cas.py:77-98
        interm_amax = interm._mvs_amax
        interm = interm.view(S, B, D, H, W, 8)
        fused = ops.vis_fuse([interm[s] for s in range(S)], [u[s] for s in range(S)])
        ops.set_absmax(fused, interm_amax)   # a convex combination of the per-pair volumes: their abs-max bounds it
        pair_depth = pair["depth"].view(S, B, H, W)
        pairs = [[pair_depth[s].unsqueeze(1), [u[s].unsqueeze(1)]] for s in range(S)]
        del interm
        fscore = ops.conv3d(_RegUNet.run(pk["fuse"], fused, am), pk["fuse_head"]).squeeze(-1)
        out = ops.depth_regress(fscore, depth_start, interval=depth_interval, conf_mode=L.CONF_WINDOW, out_depth=out_depth)
        return out["depth"], out["conf"], pairs


    def _run_train(self, ref, srcs, ref_cam, src_cams, depth_num, depth_start, depth_interval, s_scale):
        """Training mode (SingleStage.forward with mode='soft', model_cas.py:303-420): the per-pair group-correlation
        volumes come from K1 forward + backward (ops.cost_volume: no warped 32-channel volumes, no homography tensors),
        everything downstream runs as the PyTorch modules / ops of the reference so that every output -- pair depths,
        uncertainties, fused depth, probability map -- carries the reference's gradient."""
        B, H, W, _ = ref.shape
        S, D = len(srcs), depth_num
        warp = ops.vis_homography_params(ref_cam, src_cams, 1.0 / s_scale)
        cost = ops.cost_volume(ref, srcs, warp, depth_start.detach().contiguous(), D, L.GEOM_VIS, L.AGG_GROUPCORR,
                               interval=depth_interval, groups=8)                      # [S,B,D,H,W,8]
        start = depth_start.view(B, 1, 1, 1) if depth_start.dim() == 1 else depth_start.view(B, 1, H, W)
        interval = depth_interval.view(B, 1, 1, 1)
        index = torch.arange(D, dtype=ref.dtype, device=ref.device).view(1, D, 1, 1)

        def soft_argmin(score):   # nn_utils.py:453-466
            prob = torch.softmax(score, dim=1)
            return prob, torch.sum(index * prob, dim=1, keepdim=True)

        weight_sum, fused, pairs = 0, 0, []
        for s in range(S):
            interm = self.reg.unet(cost[s].permute(0, 4, 1, 2, 3))                      # [B,8,D,H,W]
            prob, cls = soft_argmin(_conv_train(self.reg_pair.final_conv, interm).squeeze(1))
            est = cls * interval + start
            ent = torch.sum(-prob * prob.clamp(1e-9, 1.).log(), dim=1, keepdim=True)    # nn_utils.py:469-470
            heads = self.uncert_net(ent)
            pairs.append([est, heads])
            weight = (-heads[0]).exp().unsqueeze(2)
            weight_sum = weight_sum + weight
            fused = fused + interm * weight
        fused = fused / weight_sum
        prob, cls = soft_argmin(_conv_train(self.reg_fuse.final_conv, self.reg_fuse.unet(fused)).squeeze(1))
        prob_map = torch.sum(prob * ((index - cls).abs() <= 2).to(prob.dtype), dim=1)
        return (cls * interval + start).squeeze(1), prob_map, pairs


class Model(nn.Module):
    def __init__(self):
        super().__init__()
        self.feat_ext = FeatExt()
        self.stage1 = SingleStage()
        self.stage2 = SingleStage()
        self.stage3 = SingleStage()


class _GraphedCascade:
    def __init__(self, net, feats, ref_cam, src_cams, depth_min, depth_interval, depth_nums, interval_scales, warmup=2,
                 out_depth=None):
        self.feats = [[f.clone() for f in fv] for fv in feats]
        self.ref_cam, self.src_cams = ref_cam.clone(), src_cams.clone()
        self.depth_min, self.depth_interval = depth_min.clone(), depth_interval.clone()
        args = (self.feats, self.ref_cam, self.src_cams, self.depth_min, self.depth_interval, depth_nums, interval_scales,
                out_depth)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # first calls pack weights and set kernel attributes: keep them out of the capture
            for _ in range(warmup):
                net.depth_from_features(*args)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = net.depth_from_features(*args)

    def __call__(self, feats=None, ref_cam=None, src_cams=None, depth_min=None, depth_interval=None):
        if feats is not None:
            for dv, sv in zip(self.feats, feats):
                for d, s in zip(dv, sv):
                    d.copy_(s, non_blocking=True)
        for dst, src in ((self.ref_cam, ref_cam), (self.src_cams, src_cams), (self.depth_min, depth_min),
                         (self.depth_interval, depth_interval)):
            if src is not None:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out


class Frontend(nn.Module):
    def __init__(self):
        super().__init__()
        self.model = Model()
        self.depth_nums = [32, 16, 8]
        self.interval_scales = [4, 2, 1]

    @staticmethod
    def fill_cam_array(K, R, t, start_depth, depth_interval):
        """[B,2,4,4]: [:,0] = R|t, [:,1] = K with [1,3,0]=depth start, [1,3,1]=interval (frontend.py:14-24)."""
        res = torch.zeros(K.shape[0], 2, 4, 4, device=K.device)
        res[:, 0, :3, :3] = R
        res[:, 0, :3, 3:4] = t
        res[:, 1, :3, :3] = K
        res[:, 1, 3, 0] = start_depth
        res[:, 1, 3, 1] = depth_interval
        return res

    def depth_from_features(self, feats, ref_cam, src_cams, depth_min, depth_interval, depth_nums, interval_scales,
                            out_depth=None):
        """feats[v][k]: channels-last [B,h_k,w_k,32] of view v (0 = reference) at stage k; ref_cam [B,2,4,4];
        src_cams [B,S,2,4,4]; depth_min, depth_interval [B].  The cascade of frontend.py:66-98.  `out_depth`: caller-owned
        [B,h_3,w_3] buffer the last stage's regression kernel writes the final depth map into."""
        stages = (self.model.stage1, self.model.stage2, self.model.stage3)
        ests, probs, pairs = [], [], []
        start = depth_min.contiguous()
        for k, (stage, s_scale) in enumerate(zip(stages, (8, 4, 2))):
            ref = feats[0][k]
            if k > 0:
                up = F.interpolate(ests[-1].detach().unsqueeze(1), size=(ref.shape[1], ref.shape[2]), mode="bilinear",
                                   align_corners=False).squeeze(1)
                # the reference reads self.interval_scales here, not the kwarg override (frontend.py:76-78)
                start = (up - depth_nums[k] * depth_interval.view(-1, 1, 1) * self.interval_scales[k] / 2).contiguous()
            d, p, pr = stage.run(ref, [f[k] for f in feats[1:]], ref_cam, src_cams, depth_nums[k], start,
                                 (depth_interval * interval_scales[k]).contiguous(), s_scale,
                                 out_depth=out_depth if k == 2 else None)
            ests.append(d)
            probs.append(p)
            pairs.append(pr)
        return ests, probs, pairs

    def graphed(self, feats, ref_cam, src_cams, depth_min, depth_interval, depth_nums, interval_scales, out_depth=None):
        """CUDA-graph version of depth_from_features for inputs of these shapes: the ~90 launches of the three stages
        (and the PyTorch glue between them) replay as one graph.  Returns a callable taking the same tensors (copied
        into the captured buffers) and returning (ests, probs, pairs) -- the captured outputs, overwritten by the next
        replay."""
        return _GraphedCascade(self, feats, ref_cam, src_cams, depth_min, depth_interval, depth_nums, interval_scales,
                               out_depth=out_depth)

    def graphed_forward(self, imgs, K, R, t, depth_min, depth_max, reference_frame=0, **kwargs):
        """The whole eval-mode `forward` (images -> dict, same kwargs) captured into ONE CUDA graph for inputs of these shapes
        (mvsnet.GraphedForward): a callable taking the same six tensors (None = keep the captured values) and returning the
        captured output dict (overwritten by the next replay).  The eager forward of this model is launch bound."""
        from .mvsnet import GraphedForward
        return GraphedForward(self, imgs, K, R, t, depth_min, depth_max, reference_frame, **kwargs)

    def forward(self, imgs, K, R, t, depth_min, depth_max, reference_frame=0, **kwargs):
        depth_interval = (depth_max - depth_min) / 128
        interval_scales = kwargs.get("interval_scales", self.interval_scales)
        depth_nums = kwargs.get("depth_nums", self.depth_nums)
        if not isinstance(imgs, list):
            imgs = list(torch.unbind(imgs, dim=1))
        v = len(imgs)
        src_idx = list(range(reference_frame)) + list(range(reference_frame + 1, v))
        # training (row f2): autograd graph over K1 forward / backward + the reference's modules; one feature-extractor call
        # per view (BatchNorm batch statistics are per call, frontend.py:59-62)
        views = (lambda f, ims: [f(im) for im in ims]) if self.training else ops.map_views
        with torch.enable_grad() if self.training else torch.no_grad():
            ref_cam = self.fill_cam_array(K[:, reference_frame], R[:, reference_frame], t[:, reference_frame],
                                          depth_min[:, reference_frame], depth_interval[:, reference_frame])
            src_cams = torch.stack([self.fill_cam_array(K[:, i], R[:, i], t[:, i], depth_min[:, i], depth_interval[:, i])
                                    for i in src_idx], 1)
            feats = [[ops.to_nhwc(f) for f in fv]
                     for fv in views(self.model.feat_ext, [imgs[i] for i in [reference_frame] + src_idx])]
            ests, probs, pairs = self.depth_from_features(feats, ref_cam, src_cams, depth_min[:, reference_frame],
                                                          depth_interval[:, reference_frame].contiguous(), depth_nums,
                                                          interval_scales)
            p1 = F.interpolate(probs[0].unsqueeze(1), scale_factor=4, mode="bilinear", align_corners=False)
            p2 = F.interpolate(probs[1].unsqueeze(1), scale_factor=2, mode="bilinear", align_corners=False)
        return {"depth": ests[2], "depth_est_list": ests[::-1], "depth_pair_list": pairs[::-1],
                "photometric_confidence": torch.cat([p1, p2, probs[2].unsqueeze(1)], dim=1)}
