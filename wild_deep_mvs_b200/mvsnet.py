"""Drop-in MVSNet / MVSNet-s (variance / softmin aggregation) running its hot path on libmvsb200.

Mirrors the reference's interface for this path -- same class name, constructor argument, attribute
names (`num_depth`, `aggregation`, `temp`), `forward(imgs, K, R, t, depth_min, depth_max,
reference_frame=0, **kwargs)` signature, returned dict and state_dict key names -- so a checkpoint
written by the reference's train.py loads unchanged (models/MVSNet/model.py:86-218).

What runs where:
  * FeatureNet (2-D CNN, models/MVSNet/model.py:21-41) -> K7 (mvsb200_conv2d) in eval mode, PyTorch modules in training
  * build_cost_volume  -> K1 (mvsb200_build_cost_volume), fused warp + aggregation
  * cost_regularization -> K2 (mvsb200_conv3d*), BN/ReLU/skip fused
  * softmax / depth regression / confidence -> K3 (mvsb200_depth_regress)
Training mode (row f2 of SURVEY.md 8): the warp + aggregation and the regression head run on the library forward AND
backward (K1 / K3 backward kernels behind torch.autograd.Function, ops.cost_volume / ops.regress_depth); the 3-D
regulariser runs as the PyTorch modules that own the parameters (batch-statistics BatchNorm) -- on cuDNN by default, on
the K2 engines forward + input gradient with mvsb200_conv3d_wgrad for the weight gradient when MVSB200_TRAIN_K2=lib.
Multi-GPU (row e): `graphed(..., gather=shard.DepthGather)` makes K3 write the depth map into the rank's slice of the
preallocated gather buffer and captures the in-place all-gather into the step's CUDA graph.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L
from . import ops


def _cbr2d(cin, cout, k, stride, pad):
    m = nn.Module()
    m.conv = nn.Conv2d(cin, cout, k, stride=stride, padding=pad, bias=False)
    m.bn = nn.BatchNorm2d(cout)
    return m


class FeatureNet(nn.Module):
    """Same parameters as the reference's FeatureNet (models/MVSNet/model.py:21-41)."""

    def __init__(self):
        super().__init__()
        spec = [(3, 8, 3, 1, 1), (8, 8, 3, 1, 1), (8, 16, 5, 2, 2), (16, 16, 3, 1, 1), (16, 16, 3, 1, 1),
                (16, 32, 5, 2, 2), (32, 32, 3, 1, 1)]
        for i, s in enumerate(spec):
            setattr(self, "conv%d" % i, _cbr2d(*s))
        self.feature = nn.Conv2d(32, 32, 3, 1, 1)

    def _pack(self):
        """Layers packed for K7 (mvsb200_conv2d): BN folded, 3-channel image weights padded to 4 input channels."""
        key = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, "_pack_key", None) != key:
            pk = []
            for i in range(7):
                m = getattr(self, "conv%d" % i)
                pk.append(ops.PackedConv2d(m.conv.weight, m.bn, stride=m.conv.stride[0], relu=True))
            pk.append(ops.PackedConv2d(self.feature.weight, None, conv_bias=self.feature.bias))
            self._packed, self._pack_key = pk, key
        return self._packed

    def run(self, x):
        """x [B,3,H,W] (any memory format) -> channels-last features [B,H/4,W/4,32] through K7."""
        B, C, H, W = x.shape
        x4 = torch.zeros(B, H, W, 4, device=x.device, dtype=torch.float32)
        x4[..., :C] = x.permute(0, 2, 3, 1)
        for layer in self._pack():
            x4 = ops.conv2d(x4, layer)
        return x4

    def forward(self, x):
        if not self.training and x.is_cuda and not torch.is_grad_enabled():
            return self.run(x).permute(0, 3, 1, 2)   # reference layout [B,32,h,w] as a view of channels-last memory
        x = x.contiguous(memory_format=torch.channels_last)
        for i in range(7):
            m = getattr(self, "conv%d" % i)
            x = F.relu(m.bn(m.conv(x)), inplace=True)
        return self.feature(x)


def _cbr3d(cin, cout):
    m = nn.Module()
    m.conv = nn.Conv3d(cin, cout, 3, padding=1, bias=False)
    m.bn = nn.BatchNorm3d(cout)
    return m


def _dbr3d(cin, cout):
    return nn.Sequential(nn.ConvTranspose3d(cin, cout, 3, padding=1, output_padding=1, stride=2, bias=False),
                         nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class CostRegNet(nn.Module):
    """3-D U-Net regulariser; parameters named as models/MVSNet/model.py:43-72, executed by K2."""

    _STRIDES = {"conv1": 2, "conv3": 2, "conv5": 2}

    def __init__(self):
        super().__init__()
        chans = [(32, 8), (8, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 64)]
        for i, (a, b) in enumerate(chans):
            setattr(self, "conv%d" % i, _cbr3d(a, b))
        self.conv7, self.conv9, self.conv11 = _dbr3d(64, 32), _dbr3d(32, 16), _dbr3d(16, 8)
        self.prob = nn.Conv3d(8, 1, 3, stride=1, padding=1)
        self._packed = None
        self._packed_key = None

    def _pack(self):
        key = tuple((p.data_ptr(), p._version) for p in self.parameters()) + tuple(
            (b.data_ptr(), b._version) for b in self.buffers())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        pk = {}
        for i in range(7):
            name = "conv%d" % i
            m = getattr(self, name)
            pk[name] = ops.PackedConv(m.conv.weight, m.bn, stride=self._STRIDES.get(name, 1), relu=True)
        for name in ("conv7", "conv9", "conv11"):
            m = getattr(self, name)
            pk[name] = ops.PackedConv(m[0].weight, m[1], stride=2, transposed=True, relu=True,
                                      skip_mode=L.SKIP_AFTER_RELU)  # x = conv4 + relu(bn(deconv)), model.py:79-81
        pk["prob"] = ops.PackedConv(self.prob.weight, None, conv_bias=self.prob.bias)
        self._packed, self._packed_key = pk, key
        return pk

    def run(self, vol):
        """vol [B,D,H,W,32] channels-last -> score [B,D,H,W]."""
        if self.training:
            raise NotImplementedError("libmvsb200 runs the regulariser in eval mode only (SURVEY.md 8-f2)")
        B, D, H, W, _ = vol.shape
        if D % 8 or H % 8 or W % 8:
            raise L.Mvsb200Error("CostRegNet needs D,H,W divisible by 8 (got %d,%d,%d)" % (D, H, W))
        pk = self._pack()
        am = ops.AmaxPool(vol.device, 16)   # abs-max scalars of the layer outputs: one fill for the whole net
        conv = lambda x, name, skip=None: ops.conv3d(x, pk[name], skip=skip, amax=am.take())
        c0 = conv(vol, "conv0")
        c2 = conv(conv(c0, "conv1"), "conv2")
        c4 = conv(conv(c2, "conv3"), "conv4")
        x = conv(conv(c4, "conv5"), "conv6")
        x = conv(x, "conv7", c4)
        x = conv(x, "conv9", c2)
        x = conv(x, "conv11", c0)
        return conv(x, "prob").squeeze(-1)

    def forward_modules(self, x):
        """The same network through the PyTorch modules that own the parameters (models/MVSNet/model.py:74-84):
        training mode, where BatchNorm uses batch statistics and autograd needs the layer graph.  The convolutions run
        on cuDNN by default; with MVSB200_TRAIN_K2=lib they run on the library forward AND backward (ops.conv3d_train:
        K2 engines for the layer and its input gradient, mvsb200_conv3d_wgrad for the weight gradient), BatchNorm / ReLU
        stay PyTorch ops on the channels-last volumes."""
        lib = os.environ.get("MVSB200_TRAIN_K2") == "lib" and x.is_cuda

        def conv(v, weight, bias=None, stride=1, transposed=False):
            if lib:
                return ops.as_ncdhw(ops.conv3d_train(ops.to_ndhwc(v), weight, bias, stride, transposed))
            if transposed:
                return F.conv_transpose3d(v, weight, bias, stride, 1, stride - 1)
            return F.conv3d(v, weight, bias, stride, 1)

        def cbr(name, v):   # the modules hold the parameters; the stride-2 layers are listed in _STRIDES
            m = getattr(self, name)
            return F.relu(m.bn(conv(v, m.conv.weight, None, self._STRIDES.get(name, 1))), inplace=True)

        def dbr(m, v):
            return m[2](m[1](conv(v, m[0].weight, None, 2, True)))

        conv0 = cbr("conv0", x)
        conv2 = cbr("conv2", cbr("conv1", conv0))
        conv4 = cbr("conv4", cbr("conv3", conv2))
        x = cbr("conv6", cbr("conv5", conv4))
        x = conv4 + dbr(self.conv7, x)
        x = conv2 + dbr(self.conv9, x)
        x = conv0 + dbr(self.conv11, x)
        return conv(x, self.prob.weight, self.prob.bias)

    def forward(self, x, down_ft=None):
        """Reference signature: x [B,32,D,H,W] -> [B,1,D,H,W] (models/MVSNet/model.py:74-84)."""
        if self.training or (torch.is_grad_enabled() and x.requires_grad):
            return self.forward_modules(x)
        return self.run(ops.to_ndhwc(x)).unsqueeze(1)


def build_proj_matrices(K, R, t):
    """[[K R, K t],[0 0 0 1]]  (utils/utils_3D.py:50-62)."""
    res = torch.zeros(K.shape[:-2] + (4, 4), device=K.device, dtype=K.dtype)
    res[..., :3, :3] = K @ R
    res[..., :3, 3:] = K @ t
    res[..., 3, 3] = 1
    return res


class GraphedHotPath:
    """The features -> (depth, confidence) path of one MVSNet captured ONCE into a CUDA graph for fixed shapes and
    replayed per sample: the ~14 kernel launches of a step (K1, 11 x K2, K3, geometry prologue) cost one
    cudaGraphLaunch instead of 14 ctypes calls, which matters once a step is down to a couple of milliseconds.
    Inputs are copied into the captured buffers; the returned tensors are the captured outputs (overwritten by the
    next replay -- clone them to keep them)."""

    def __init__(self, net, feats_nhwc, proj_matrices, depth_values, reference_frame=0, warmup=2, gather=None):
        """`gather` (shard.DepthGather, optional): K3 writes the depth map into this rank's slice of the gather buffer and
        the in-place all-gather is part of the captured graph; `self.gathered` is then the [world * B, h, w] result."""
        self.feats = [f.clone() for f in feats_nhwc]
        self.projs = [p.clone() for p in proj_matrices]
        self.depth_values = depth_values.clone()
        self.gather = gather
        out_depth = gather.local if gather is not None else None
        run = lambda: net.depth_from_features(self.feats, self.projs, self.depth_values, reference_frame, out_depth=out_depth)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # first calls pack weights and set kernel attributes: keep them out of the capture
            for _ in range(warmup):
                run()
                if gather is not None:
                    gather.all_gather()   # the communicator's first collective (connection setup) cannot be captured
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.depth, self.conf = run()
            self.gathered = gather.all_gather() if gather is not None else None

    def __call__(self, feats_nhwc=None, proj_matrices=None, depth_values=None):
        if feats_nhwc is not None:
            for dst, src in zip(self.feats, feats_nhwc):
                dst.copy_(src, non_blocking=True)
        if proj_matrices is not None:
            for dst, src in zip(self.projs, proj_matrices):
                dst.copy_(src, non_blocking=True)
        if depth_values is not None:
            self.depth_values.copy_(depth_values, non_blocking=True)
        self.graph.replay()
        return self.depth, self.conf


class GraphedForward:
    """The eval-mode `forward(imgs, K, R, t, depth_min, depth_max, ...)` of a drop-in model (MVSNet, Vis-MVSNet or
    CVP-MVSNet `Frontend`) as one CUDA graph; see the models' graphed_forward."""

    def __init__(self, net, imgs, K, R, t, depth_min, depth_max, reference_frame=0, warmup=2, **kwargs):
        if net.training or not isinstance(imgs, torch.Tensor):
            raise L.Mvsb200Error("graphed_forward: eval mode and same-sized views (a [B,V,3,H,W] tensor) only")
        self.inputs = [x.clone() for x in (imgs, K, R, t, depth_min, depth_max)]
        run = lambda: net(*self.inputs, reference_frame=reference_frame, **kwargs)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                run()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = run()

    def __call__(self, *inputs):
        for dst, src in zip(self.inputs, inputs):
            if src is not None:
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        return self.out


class StreamedHotPath:
    """A stream of samples whose inputs sit in PINNED HOST memory (the reconstruction pipeline's loop over the reference
    views of a scene, evaluation/run_depthmaps.py:53-68): `slots` graph instances are used round-robin, the H2D copies of
    sample i+1 run on a copy stream while sample i's graph runs, and depth + confidence of every sample are copied back
    to pinned host buffers.  submit() returns (host_depth, host_conf, event); the buffers are valid once the event has
    completed and until the slot is reused (`slots` submissions later)."""

    def __init__(self, net, feats_nhwc, proj_matrices, depth_values, reference_frame=0, slots=2, gathers=None):
        """`gathers`: one shard.DepthGather per slot -- every sample's depth map is then all-gathered inside its graph
        and `submit` copies the GATHERED maps of all ranks back to the host."""
        self.slots = [GraphedHotPath(net, feats_nhwc, proj_matrices, depth_values, reference_frame,
                                     gather=gathers[k] if gathers else None) for k in range(slots)]
        self.copy_stream = torch.cuda.Stream()
        self.ready = [torch.cuda.Event() for _ in range(slots)]
        self.done = [torch.cuda.Event() for _ in range(slots)]
        for e in self.done:
            e.record()
        g = self.slots[0]
        dshape = g.gathered.shape if g.gathered is not None else g.depth.shape
        self.out_depth = [torch.empty(dshape, dtype=g.depth.dtype).pin_memory() for _ in range(slots)]
        self.out_conf = [torch.empty(g.conf.shape, dtype=g.conf.dtype).pin_memory() for _ in range(slots)]
        self.count = 0

    def submit(self, host_feats_nhwc, host_projs, host_depth_values):
        k = self.count % len(self.slots)
        self.count += 1
        g = self.slots[k]
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.done[k])   # the slot's previous sample has been computed and read back
            for dst, src in zip(g.feats, host_feats_nhwc):
                dst.copy_(src, non_blocking=True)
            for dst, src in zip(g.projs, host_projs):
                dst.copy_(src, non_blocking=True)
            g.depth_values.copy_(host_depth_values, non_blocking=True)
            self.ready[k].record(self.copy_stream)
        main.wait_event(self.ready[k])
        g.graph.replay()
        self.out_depth[k].copy_(g.gathered if g.gathered is not None else g.depth, non_blocking=True)
        self.out_conf[k].copy_(g.conf, non_blocking=True)
        self.done[k].record(main)
        return self.out_depth[k], self.out_conf[k], self.done[k]


class MVSNet(nn.Module):
    def __init__(self, aggregation="variance"):
        super().__init__()
        self.feature = FeatureNet()
        self.cost_regularization = CostRegNet()
        if aggregation == "softmin":
            self.register_parameter("temp", nn.Parameter(torch.ones(1)))
        self.aggregation = aggregation
        self.num_depth = 192

    def extract_features(self, imgs):
        # training: one call per view like the reference (model.py:100-107) -- BatchNorm's batch statistics are per call
        views = (lambda f, ims: [f(im) for im in ims]) if self.training else ops.map_views
        if self.aggregation.startswith("norm"):
            return [F.normalize(f, dim=1) for f in views(self.feature, imgs)]
        return views(self.feature, imgs)

    # ---- channels-last engine entry (what forward uses) ------------------------------------------
    def cost_volume_cl(self, ref_nhwc, srcs_nhwc, ref_proj, src_projs, depth_values):
        """ref [B,H,W,C]; srcs list of [B,Hs,Ws,C]; projections [B,4,4]; depth_values [B,D] or [B,D,H,W]."""
        if self.aggregation == "variance":
            agg, temp = L.AGG_VARIANCE, None
        elif self.aggregation == "softmin":
            agg, temp = L.AGG_SOFTMIN, self.temp
        else:
            raise NotImplementedError("Aggregation: " + self.aggregation)
        warp = ops.mvs_relative_proj(ref_proj, torch.stack(list(src_projs), 1))
        if torch.is_grad_enabled() and (ref_nhwc.requires_grad or any(f.requires_grad for f in srcs_nhwc)
                                        or (temp is not None and temp.requires_grad)):
            return ops.cost_volume(ref_nhwc, srcs_nhwc, warp, depth_values, self.num_depth, L.GEOM_MVS, agg, temp=temp)
        return ops.build_cost_volume(ref_nhwc, srcs_nhwc, warp, depth_values, self.num_depth, L.GEOM_MVS, agg, temp=temp)

    def build_cost_volume(self, ref_feature, src_features, ref_proj, src_projs, depth_values):
        """Reference signature (models/MVSNet/model.py:109): NCHW features in, [B,C,D,H,W] out (a view)."""
        vol = self.cost_volume_cl(ops.to_nhwc(ref_feature), [ops.to_nhwc(f) for f in src_features], ref_proj,
                                  src_projs, depth_values)
        return ops.as_ncdhw(vol)

    def depth_from_features(self, feats_nhwc, proj_matrices, depth_values, reference_frame=0, out_depth=None, out_conf=None):
        """The hot path proper: channels-last features -> (depth, confidence).  feats_nhwc: list of V maps.
        `out_depth` / `out_conf` [B,h,w]: caller-owned output buffers (e.g. shard.DepthGather.local)."""
        ref = feats_nhwc[reference_frame]
        srcs = feats_nhwc[:reference_frame] + feats_nhwc[reference_frame + 1:]
        ref_proj = proj_matrices[reference_frame]
        src_projs = proj_matrices[:reference_frame] + proj_matrices[reference_frame + 1:]
        vol = self.cost_volume_cl(ref, srcs, ref_proj, src_projs, depth_values)
        score = self.cost_regularization.run(vol)
        del vol
        out = ops.depth_regress(score, depth_values, conf_mode=L.CONF_SUM4, out_depth=out_depth, out_conf=out_conf)
        return out["depth"], out["conf"]

    def _forward_differentiable(self, imgs, proj_matrices, depth_values, reference_frame):
        """Training path (models/trainer.py:96-206 calls forward with grad enabled): features -> K1 (autograd) ->
        regulariser modules -> K3 (autograd).  Gradients reach the feature extractor, the regulariser and `temp`."""
        features = self.extract_features(imgs)
        ref = features[reference_frame]
        srcs = features[:reference_frame] + features[reference_frame + 1:]
        ref_proj = proj_matrices[reference_frame]
        src_projs = proj_matrices[:reference_frame] + proj_matrices[reference_frame + 1:]
        vol = self.cost_volume_cl(ops.to_nhwc(ref), [ops.to_nhwc(f) for f in srcs], ref_proj, src_projs, depth_values)
        score = self.cost_regularization(ops.as_ncdhw(vol)).squeeze(1)
        return ops.regress_depth(score, depth_values, conf_mode=L.CONF_SUM4)

    def graphed(self, feats_nhwc, proj_matrices, depth_values, reference_frame=0, gather=None):
        """CUDA-graph version of depth_from_features for inputs of these shapes (see GraphedHotPath)."""
        return GraphedHotPath(self, feats_nhwc, proj_matrices, depth_values, reference_frame, gather=gather)

    def streamed(self, feats_nhwc, proj_matrices, depth_values, reference_frame=0, slots=2, gathers=None):
        """Double-buffered, copy-overlapped version for a stream of pinned-host samples (see StreamedHotPath)."""
        return StreamedHotPath(self, feats_nhwc, proj_matrices, depth_values, reference_frame, slots, gathers)

    def graphed_forward(self, imgs, K, R, t, depth_min, depth_max, reference_frame=0):
        """The whole eval-mode `forward` (images -> dict) captured into ONE CUDA graph for inputs of these shapes: K7
        feature extractor, geometry, K1, K2, K3 and the PyTorch glue between them replay as a single launch.  Returns a
        callable taking the same tensors (copied into the captured buffers; None = keep the captured values) and
        returning the captured output dict (overwritten by the next replay)."""
        return GraphedForward(self, imgs, K, R, t, depth_min, depth_max, reference_frame)

    def forward(self, imgs, K, R, t, depth_min, depth_max, reference_frame=0, **kwargs):
        try:
            imgs = torch.unbind(imgs, 1)
        except TypeError:  # already a list (views of different sizes)
            pass
        scaled_K = K.clone()
        scaled_K[:, :, :2] /= 4
        proj_matrices = torch.unbind(build_proj_matrices(scaled_K, R, t), 1)
        steps = torch.arange(self.num_depth, device=depth_min.device).view(1, 1, -1)
        depth_range = (depth_max - depth_min) / (self.num_depth - 1)
        depth_values = depth_min.unsqueeze(-1) + depth_range.unsqueeze(-1) * steps
        assert len(imgs) == len(proj_matrices), "Different number of images and projection matrices"

        if self.training:
            depth, conf = self._forward_differentiable(imgs, list(proj_matrices), depth_values[:, reference_frame].contiguous(),
                                                       reference_frame)
        else:
            with torch.no_grad():
                feats = [ops.to_nhwc(f) for f in self.extract_features(imgs)]
                depth, conf = self.depth_from_features(list(feats), list(proj_matrices),
                                                       depth_values[:, reference_frame].contiguous(), reference_frame)
        return {"depth": depth, "depth_est_list": [depth], "depth_pair_list": [], "photometric_confidence": conf}
